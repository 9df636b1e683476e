// batched_hybrid_tu_a.cu -- instantiates the batched BFGS step kernel for n = 2 ... 16 (own translation unit: ptxas
// works on the sixteen kernels of batched_hybrid.cuh in two parallel halves).
#include "batched_hybrid.cuh"

namespace dzo {
cudaError_t hybrid_launch_n2_16(int n, const BatchedArgs& args, cudaStream_t stream, int device) {
    switch (n) {
        case 2: return hybrid_launch<2>(args, stream, device);
        case 4: return hybrid_launch<4>(args, stream, device);
        case 6: return hybrid_launch<6>(args, stream, device);
        case 8: return hybrid_launch<8>(args, stream, device);
        case 10: return hybrid_launch<10>(args, stream, device);
        case 12: return hybrid_launch<12>(args, stream, device);
        case 14: return hybrid_launch<14>(args, stream, device);
        case 16: return hybrid_launch<16>(args, stream, device);
        default: return cudaErrorInvalidValue;
    }
}
}  // namespace dzo
