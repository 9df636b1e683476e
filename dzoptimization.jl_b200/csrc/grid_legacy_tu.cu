// grid_legacy_tu.cu -- instantiates the grid-wide legacy L-BFGS / GD kernel in its own translation unit.
#include "grid_legacy_lbfgs.cuh"
namespace dzo {
void* grid_legacy_kernel_ptr(int own) {
    return own == 1 ? (void*)grid_legacy_lbfgs_kernel<1> : (void*)grid_legacy_lbfgs_kernel<kGridOwnMax>;
}
}
