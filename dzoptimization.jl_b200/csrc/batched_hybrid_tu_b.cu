// batched_hybrid_tu_b.cu -- instantiates the batched BFGS step kernel for n = 18 ... 32 (see batched_hybrid_tu_a.cu).
#include "batched_hybrid.cuh"

namespace dzo {
cudaError_t hybrid_launch_n18_32(int n, const BatchedArgs& args, cudaStream_t stream, int device) {
    switch (n) {
        case 18: return hybrid_launch<18>(args, stream, device);
        case 20: return hybrid_launch<20>(args, stream, device);
        case 22: return hybrid_launch<22>(args, stream, device);
        case 24: return hybrid_launch<24>(args, stream, device);
        case 26: return hybrid_launch<26>(args, stream, device);
        case 28: return hybrid_launch<28>(args, stream, device);
        case 30: return hybrid_launch<30>(args, stream, device);
        case 32: return hybrid_launch<32>(args, stream, device);
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace dzo
