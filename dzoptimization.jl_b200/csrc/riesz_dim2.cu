// riesz_dim2.cu -- instantiates riesz_gd_kernel<2, 512> (the default: 512-thread CTAs, 128 registers per thread) in its
// own translation unit: each instance takes ptxas a minute, the eight units (dimension x CTA size) compile in parallel.
// dzopt_gd.cu launches the kernel through the pointer returned here.
#include "gd_kernels.cuh"
namespace dzo {
void* riesz_kernel_dim2_t1024();
void* riesz_kernel_dim2(int nt) { return nt == 512 ? (void*)riesz_gd_kernel<2, 512> : riesz_kernel_dim2_t1024(); }
}
