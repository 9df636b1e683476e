// small_ops.cuh -- kernel-level entry points in SEQUENTIAL order for n <= DZO_SMALL_N_MAX.
//
// These expose, one at a time, the pieces the batched step! kernel (batched_bfgs.cuh) is made
// of, so tests/ can pin every row of SURVEY.md 8a against the sequential oracle bit for bit:
// Kernels.dot (legacy/Kernels.jl:12-20), mul! (legacy/DZOptimization.jl:875,:958),
// update_inverse_hessian! (:864-889), quadratic_line_search (:49-216) and the example
// objectives.  One warp per problem, lane r owns element r / row r.
#pragma once
#include "batched_bfgs.cuh"

namespace dzo {

DZO_DEVINL Group<32> make_warp_group(int n, double* bc) {
    Group<32> G;
    G.r = threadIdx.x & 31;
    G.n = n;
    G.act = G.r < n;
    G.mask = 0xffffffffu;
    G.bc = bc;
    G.flip = 0;
    G.probes = 0;
    return G;
}

template <class Obj>
static __global__ void __launch_bounds__(32) small_objective_kernel(const double* x, int n, double* f) {
    __shared__ double bc[kBcBufs * 32];
    Group<32> G = make_warp_group(n, bc);
    const long long p = blockIdx.x;
    const double xr = G.act ? x[p * n + G.r] : 0.0;
    const double v = Obj::eval(G, xr);
    if (G.r == 0) f[p] = v;
}

template <class Obj>
static __global__ void __launch_bounds__(32) small_gradient_kernel(const double* x, int n, double* g) {
    __shared__ double bc[kBcBufs * 32];
    Group<32> G = make_warp_group(n, bc);
    const long long p = blockIdx.x;
    const double xr = G.act ? x[p * n + G.r] : 0.0;
    const double v = Obj::grad(G, xr);
    if (G.act) g[p * n + G.r] = v;
}

static __global__ void __launch_bounds__(32) small_dot_kernel(const double* v, const double* w, int n, double* out) {
    __shared__ double bc[kBcBufs * 32];
    Group<32> G = make_warp_group(n, bc);
    const double s = G.seq_sum(G.act ? v[G.r] * w[G.r] : 0.0);
    if (G.r == 0) *out = s;
}

static __global__ void __launch_bounds__(32) small_gemv_kernel(const double* H, const double* v, int n, double* out) {
    const int r = threadIdx.x;
    if (r >= n) return;
    double acc = 0.0;
    for (int j = 0; j < n; ++j) acc += H[r + j * n] * v[j];
    out[r] = acc;
}

// update_inverse_hessian!(H, step_length, step_direction, delta_gradient, scratch) + fused mul!
static __global__ void __launch_bounds__(32) small_update_kernel(double* H, double step_length, double* sd_io,
                                                          const double* dg_in, double* scratch, const double* next_g,
                                                          double* next_d, int n) {
    __shared__ double bc[kBcBufs * 32];
    Group<32> G = make_warp_group(n, bc);
    const int r = G.r;
    const double d = G.act ? sd_io[r] : 0.0;
    const double dg = G.act ? dg_in[r] : 0.0;
    const double g = (G.act && next_g) ? next_g[r] : 0.0;
    const double overlap = G.seq_sum(d * dg);                           // :873
    const double sd = d * (1.0 / overlap);                              // :874
    double* v0 = G.vbuf(0);
    double* v1 = G.vbuf(1);
    double* v2 = G.vbuf(2);
    if (G.act) v0[r] = dg;
    G.sync();
    double t = 0.0;                                                     // :875
    if (G.act)
        for (int j = 0; j < n; ++j) t += H[r + j * n] * v0[j];
    const double delta_norm = step_length * overlap + G.seq_sum(dg * t); // :876
    if (G.act) { v0[r] = g; v1[r] = sd; v2[r] = t; }
    G.sync();
    double dnew = 0.0;
    if (G.act)
        for (int j = 0; j < n; ++j) {
            const double sj = v1[j], tj = v2[j];
            double h = H[r + j * n];
            h += (delta_norm * (sd * sj) - (t * sj + sd * tj));         // :882-884
            H[r + j * n] = h;
            dnew += h * v0[j];                                          // :958-960
        }
    if (G.act) {
        sd_io[r] = sd;
        scratch[r] = t;
        if (next_g && next_d) next_d[r] = dnew;
    }
}

template <class Obj>
static __global__ void __launch_bounds__(32) small_line_search_kernel(const double* x, const double* dir, int n, double f0,
                                                               double t1, double* out2) {
    __shared__ double bc[kBcBufs * 32];
    Group<32> G = make_warp_group(n, bc);
    const double xr = G.act ? x[G.r] : 0.0;
    const double dr = G.act ? dir[G.r] : 0.0;
    double tb, fb;
    group_line_search<32, Obj>(G, xr, dr, f0, t1, tb, fb);
    if (G.r == 0) { out2[0] = tb; out2[1] = fb; }
}

}  // namespace dzo
