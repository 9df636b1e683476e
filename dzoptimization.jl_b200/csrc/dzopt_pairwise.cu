// dzopt_pairwise.cu -- C ABI of the pairwise radial N-body kernels (src/ExampleFunctions.jl:117-468 of the
// live package).  sm_100a only; -fmad=false (explicit fma() where the reference says muladd).
#include "host_common.h"
#include "pairwise.cuh"

using namespace dzo;

static int pairwise_check(int potential, int order, int64_t n) {
    if (potential != DZO_POT_LENNARD_JONES) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown radial potential id");
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    if (n <= 0) return fail(DZO_ERR_INVALID_ARGUMENT, "n must be positive");
    return DZO_OK;
}

template <int WHAT>
static int launch_pairwise(cudaStream_t stream, int order, const PairwiseArgs& a) {
    if (order == DZO_ORDER_SEQUENTIAL) {
        const unsigned grid = (unsigned)((a.n + kPairSeqThreads - 1) / kPairSeqThreads);
        pairwise_seq_kernel<WHAT, LennardJones><<<grid, kPairSeqThreads, 0, stream>>>(a);
    } else {
        const unsigned grid = (unsigned)((a.n + 31) / 32);
        pairwise_tree_kernel<WHAT, LennardJones><<<grid, 32 * PairTreeWarps<WHAT>::value, 0, stream>>>(a);
    }
    DZO_CUDA(cudaGetLastError());
    return DZO_OK;
}

extern "C" {

uint64_t dzo_pairwise_workspace_bytes(int64_t n) { return n > 0 ? (uint64_t)n * 8u + 64u : 64u; }

int dzo_pairwise_energy_device(void* cuda_stream, int potential, int order, int64_t n, const double* x, const double* y,
                               const double* z, double* point_energies, double* energy_out, void* workspace) {
    DZO_TRY(pairwise_check(potential, order, n));
    if (!x || !y || !z || !energy_out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    if (!point_energies && !workspace) return fail(DZO_ERR_INVALID_ARGUMENT, "energy needs point_energies or a workspace");
    PairwiseArgs a{};
    a.n = n; a.x = x; a.y = y; a.z = z;
    a.o0 = point_energies ? point_energies : static_cast<double*>(workspace);
    cudaStream_t s = (cudaStream_t)cuda_stream;
    DZO_TRY(launch_pairwise<0>(s, order, a));
    pairwise_energy_sum_kernel<<<1, 1024, 0, s>>>(a.o0, n, energy_out);     // :172 sum(point_energies)
    DZO_CUDA(cudaGetLastError());
    return DZO_OK;
}

int dzo_pairwise_gradient_device(void* cuda_stream, int potential, int order, int64_t n, const double* x, const double* y,
                                 const double* z, double* gx, double* gy, double* gz, void* workspace) {
    (void)workspace;
    DZO_TRY(pairwise_check(potential, order, n));
    if (!x || !y || !z || !gx || !gy || !gz) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    PairwiseArgs a{};
    a.n = n; a.x = x; a.y = y; a.z = z; a.o0 = gx; a.o1 = gy; a.o2 = gz;
    return launch_pairwise<1>((cudaStream_t)cuda_stream, order, a);
}

int dzo_pairwise_hvp_device(void* cuda_stream, int potential, int order, int64_t n, const double* x, const double* y,
                            const double* z, const double* u, const double* v, const double* w, double* px, double* py,
                            double* pz, void* workspace) {
    (void)workspace;
    DZO_TRY(pairwise_check(potential, order, n));
    if (!x || !y || !z || !u || !v || !w || !px || !py || !pz) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    PairwiseArgs a{};
    a.n = n; a.x = x; a.y = y; a.z = z; a.u = u; a.v = v; a.w = w; a.o0 = px; a.o1 = py; a.o2 = pz;
    return launch_pairwise<2>((cudaStream_t)cuda_stream, order, a);
}

// ---- host-buffer flavours
static int up(DevBuf& b, const double* src, int64_t n) {
    DZO_TRY(b.alloc((size_t)n * 8));
    if (src) DZO_CUDA(cudaMemcpy(b.p, src, (size_t)n * 8, cudaMemcpyHostToDevice));
    return DZO_OK;
}
static int down(double* dst, const DevBuf& b, int64_t n) {
    DZO_CUDA(cudaMemcpy(dst, b.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return DZO_OK;
}

int dzo_dev_pairwise_energy(int potential, int order, int64_t n, const double* x, const double* y, const double* z,
                            double* point_energies, double* energy, int device) {
    DZO_TRY(pairwise_check(potential, order, n));
    if (!x || !y || !z || !energy) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    DZO_TRY(use_device(device));
    DevBuf dx, dy, dz, de, dout;
    DZO_TRY(up(dx, x, n)); DZO_TRY(up(dy, y, n)); DZO_TRY(up(dz, z, n)); DZO_TRY(up(de, nullptr, n)); DZO_TRY(up(dout, nullptr, 1));
    DZO_TRY(dzo_pairwise_energy_device(nullptr, potential, order, n, dx.as<double>(), dy.as<double>(), dz.as<double>(),
                                       de.as<double>(), dout.as<double>(), nullptr));
    DZO_CUDA(cudaDeviceSynchronize());
    if (point_energies) DZO_TRY(down(point_energies, de, n));
    return down(energy, dout, 1);
}

int dzo_dev_pairwise_gradient(int potential, int order, int64_t n, const double* x, const double* y, const double* z,
                              double* gx, double* gy, double* gz, int device) {
    DZO_TRY(pairwise_check(potential, order, n));
    if (!x || !y || !z || !gx || !gy || !gz) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    DZO_TRY(use_device(device));
    DevBuf dx, dy, dz, ox, oy, oz;
    DZO_TRY(up(dx, x, n)); DZO_TRY(up(dy, y, n)); DZO_TRY(up(dz, z, n));
    DZO_TRY(up(ox, nullptr, n)); DZO_TRY(up(oy, nullptr, n)); DZO_TRY(up(oz, nullptr, n));
    DZO_TRY(dzo_pairwise_gradient_device(nullptr, potential, order, n, dx.as<double>(), dy.as<double>(), dz.as<double>(),
                                         ox.as<double>(), oy.as<double>(), oz.as<double>(), nullptr));
    DZO_CUDA(cudaDeviceSynchronize());
    DZO_TRY(down(gx, ox, n)); DZO_TRY(down(gy, oy, n));
    return down(gz, oz, n);
}

int dzo_dev_pairwise_hvp(int potential, int order, int64_t n, const double* x, const double* y, const double* z,
                         const double* u, const double* v, const double* w, double* px, double* py, double* pz, int device) {
    DZO_TRY(pairwise_check(potential, order, n));
    if (!x || !y || !z || !u || !v || !w || !px || !py || !pz) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    DZO_TRY(use_device(device));
    DevBuf dx, dy, dz, du, dv, dw, ox, oy, oz;
    DZO_TRY(up(dx, x, n)); DZO_TRY(up(dy, y, n)); DZO_TRY(up(dz, z, n));
    DZO_TRY(up(du, u, n)); DZO_TRY(up(dv, v, n)); DZO_TRY(up(dw, w, n));
    DZO_TRY(up(ox, nullptr, n)); DZO_TRY(up(oy, nullptr, n)); DZO_TRY(up(oz, nullptr, n));
    DZO_TRY(dzo_pairwise_hvp_device(nullptr, potential, order, n, dx.as<double>(), dy.as<double>(), dz.as<double>(),
                                    du.as<double>(), dv.as<double>(), dw.as<double>(), ox.as<double>(), oy.as<double>(),
                                    oz.as<double>(), nullptr));
    DZO_CUDA(cudaDeviceSynchronize());
    DZO_TRY(down(px, ox, n)); DZO_TRY(down(py, oy, n));
    return down(pz, oz, n);
}

}  // extern "C"
