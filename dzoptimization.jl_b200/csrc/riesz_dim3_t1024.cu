// riesz_dim3_t1024.cu -- instantiates riesz_gd_kernel<3, 1024> (1024-thread CTAs: the symmetric-tile gradient variant
// and the A/B baseline of the 512-thread kernel) in its own translation unit.
#include "gd_kernels.cuh"
namespace dzo {
void* riesz_kernel_dim3_t1024() { return (void*)riesz_gd_kernel<3, 1024>; }
}
