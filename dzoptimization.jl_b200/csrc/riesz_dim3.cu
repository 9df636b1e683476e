// riesz_dim3.cu -- instantiates riesz_gd_kernel<3, 512 | 1024> in its own translation unit (each instance takes ptxas about a
// minute; four units compile in parallel).  dzopt_gd.cu launches it through the pointer returned here.
#include "gd_kernels.cuh"
namespace dzo {
void* riesz_kernel_dim3(int nt) { return nt == 512 ? (void*)riesz_gd_kernel<3, 512> : (void*)riesz_gd_kernel<3, 1024>; }
}
