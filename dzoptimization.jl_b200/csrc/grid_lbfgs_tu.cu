// grid_lbfgs_tu.cu -- instantiates the grid-wide live L-BFGS and AdGD kernels in their own translation unit
// (dzopt_lbfgs.cu launches them through these pointers; keeps the slowest ptxas jobs in parallel).
#include "grid_lbfgs.cuh"
namespace dzo {
void* grid_lbfgs_kernel_ptr(int own) { return own == 1 ? (void*)grid_lbfgs_kernel<1> : (void*)grid_lbfgs_kernel<kGridOwnMax>; }
void* grid_adgd_kernel_ptr() { return (void*)grid_adgd_kernel<0>; }
}
