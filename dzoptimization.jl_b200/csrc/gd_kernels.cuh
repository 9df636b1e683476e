// gd_kernels.cuh -- GradientDescentOptimizer step! (legacy/DZOptimization.jl:330-374, :393-449)
// with device objectives, TREE summation order.
//
// (1) Riesz energy on the sphere (legacy/ExampleFunctions.jl:30-83; config 5: N = 4096 points,
//     n = 12288): ONE cooperative persistent kernel (one 512-thread CTA per SM; NT template parameter) runs k whole
//     step! calls -- the bracketing line search with its data-dependent number of O(N^2) energy
//     evaluations, the point update, the O(N^2) gradient and the O(n) bookkeeping -- with grid-wide
//     barriers between phases and no host round trip.  Pair work is cut into (32 rows x 128
//     sources) warp items; the 128 trial points of a segment are staged once per item in a
//     per-warp shared-memory buffer and broadcast to the 32 row lanes.
//       energy   e_j   = sum over segments s (ascending, 128 sources each, only i < j) of the
//                        sequential segment partial;  f = canonical tree over rows (row j -> virtual
//                        thread j mod 4096)                                     [oracle riesz_energy_tree]
//       gradient G_j   = sum over segments (ascending) of sequential partials skipping i == j,
//                        then the tangent projection g_j -= (p_j . g_j) p_j      [oracle gradient_]
// (2) extended Rosenbrock at n > 32 runs on the legacy L-BFGS kernels with algo = 1 (legacy_lbfgs.cuh, grid_legacy_lbfgs.cuh).
#pragma once
#include <cooperative_groups.h>

#include "ieee_fast.cuh"
#include "large_bfgs.cuh"

namespace dzo {
namespace cg = cooperative_groups;

struct GdCtrl {
    double f;        // current_objective_value   legacy/DZOptimization.jl:312
    double df;       // delta_objective_value     :313
    double L;        // last_step_length          :321
    long long iter;  // iteration_count           :324
    int term;        // has_terminated            :325
    int pad;
    long long evals;
};

struct RieszGdArgs {
    double *x, *g, *d, *dx, *dg;  // dim x N column-major (point j contiguous)
    double *segE;                 // [2][nseg][N]    energy segment partials (probe 0 | probe 1 of a paired evaluation)
    double *rowE;                 // [2][2][N]       row energies [parity of the phase count][probe]
    unsigned* rbcnt;              // items finished per 32-row block (energy probe 0 | energy probe 1 | gradient); zero between uses
    double *segG;                 // [nseg][N][DIM]  gradient segment partials
    const int2* e_items;          // (row block, segment) items of the strict lower triangle
    int n_e_items;
    GdCtrl* ctrl;
    unsigned* counter;            // (unused by the kernel since the row sums moved into the segment phase)
    double* fbox;                 // result slots of the kernel-level modes 2 and 4
    const int2* g_jobs;           // gradient tile jobs (segment a, segment b), a <= b  (gvariant 1)
    int n_g_jobs;
    int gvariant;                 // 0: (32 rows x 128 sources) warp items; 1: symmetric 128 x 128 CTA tiles
    int esplit;                   // 1: energy items of 32 rows x 128 sources, one lane per row; 2: 16 rows, two lanes per row
    int gcnt_off;                 // offset of the gradient counters inside rbcnt
    int ecnt_stride;              // offset of the second probe's energy counters inside rbcnt (paired evaluations)
    unsigned long long* bar;      // flag-word barrier of the k-step mode (null: cooperative-groups grid.sync everywhere):
                                  //   [kRieszBarStride CTAs][kRieszBarStride sources] arrival counters, then the running count
    int espec;                    // 1: the bracketing search evaluates the probe it needs and the one it will most likely
                                  //    need next in ONE phase (paired evaluation; esplit = 1 only)
    double dscale;                // the direction actually used is dscale * dir[e] (1.0, or alpha with dir = g: see mode 0)
    int N, sphere, max_increases, ksteps;
    double initial_step_length;
    int mode;                     // 0 = k GD step! calls, 1 = GD constructor, 2 = one energy evaluation of x,
                                  // 3 = one gradient of x (kernel-level entries), 4 = one line search,
                                  // 5 = the O(n)/O(N^2) stage of a BFGS step! (search, decision, bookkeeping),
                                  // 6 = BFGSOptimizer constructor
    double ls_f0, ls_t1, ls_sign; // mode 4 inputs; result in fbox[1], fbox[2]
    unsigned long long* prof;     // optional phase log of the leader thread: [0] = events, then (id, %globaltimer ns) pairs
    LargeCtrl* bctrl;             // modes 5, 6: the BFGS control block shared with the n^2 sweep kernels
    double* sd;                   // mode 5: step_direction / overlap  (legacy/DZOptimization.jl:874)
};

constexpr int kRieszProfCap = 8192;  // events of the optional phase log
// phase ids: 1 step begin, 2 energy begin, 3 leader's own energy items done, 4 first energy barrier passed,
// 5 energy end (rows + tree + barrier), 6 line search end, 7 point update + barrier, 8 leader's own gradient items
// done, 9 gradient barrier passed, 10 gradient rows + barrier, 11 step end, 12 paired energy begin, 13 both norms done
DZO_DEVINL void riesz_prof_mark(const RieszGdArgs& a, int id) {
    if (a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned long long k = a.prof[0];
        if (k < (unsigned long long)kRieszProfCap) {
            a.prof[1 + 2 * k] = (unsigned long long)id;
            a.prof[2 + 2 * k] = t;
            a.prof[0] = k + 1;
        }
    }
}

#ifndef DZO_RIESZ_BATCH
#define DZO_RIESZ_BATCH 4
#endif
#ifndef DZO_RIESZ_GRAD_BATCH
#define DZO_RIESZ_GRAD_BATCH 2
#endif
constexpr int kRieszBatch = DZO_RIESZ_BATCH;            // energy pair terms whose sqrt / reciprocal chains run interleaved
constexpr int kRieszGradBatch = DZO_RIESZ_GRAD_BATCH;   // same for the gradient (sqrt, reciprocal, division per term)

// Grid-wide barrier of the k-step mode without an atomic counter (round 2; the same idea as the flagged lines of
// grid_lbfgs.cuh; tuning knob "riesz_bar", OFF by default: measured 4 % slower than grid.sync() on config 5 -- where a
// reduction's VALUES have to travel anyway the inboxes win, for a bare barrier one atomic counter is cheaper): every CTA owns an inbox of 64-bit arrival counters, one per CTA of the grid.  Arriving = storing the
// barrier's running number into MY word of every CTA's inbox (relaxed stores behind __syncthreads + one fence per lane of
// warp 0, so every thread's writes of the phase are visible first); waiting = warp 0 polling its own inbox (relaxed
// loads, warp-uniform exit) until every word has reached the number, then a fence (acquire side: it also drops the SM's
// stale L1 lines), then __syncthreads.  No word
// is written by two CTAs and no line is polled by two: nothing serialises in an L2 slice the way one shared counter does.
// The running number lives behind the inboxes and survives across launches (64 bits: no wrap).
constexpr int kRieszBarStride = 192;                     // CTAs the barrier buffer is laid out for (a B200 runs 148)
struct RieszBar {
    cg::grid_group g;
    unsigned long long* base;                            // null: g.sync()
    unsigned long long seq;
    int nctas, cta;
    DZO_DEVINL void sync() {
        if (base == nullptr) { g.sync(); return; }
        seq += 1;
        __syncthreads();                                     // every thread's writes of the phase happen before ...
        if (threadIdx.x < 32) {
            const int lane = threadIdx.x;
            __threadfence();                                 // ... this fence (cumulative), which orders them before the arrival words
            for (int dst = lane; dst < nctas; dst += 32)
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(base + (size_t)dst * kRieszBarStride + cta), "l"(seq) : "memory");
            const unsigned long long* mine = base + (size_t)cta * kRieszBarStride;
            unsigned long long t0 = 0;
            for (unsigned spins = 1;; ++spins) {
                bool ok = true;
                for (int src = lane; src < nctas; src += 32) {
                    unsigned long long v;
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + src) : "memory");
                    ok &= (v >= seq);
                }
                if (__all_sync(0xffffffffu, ok)) break;
                if ((spins & 1023u) == 0u) {                 // a grid that lost a CTA must not hang the GPU
                    const unsigned long long now = global_timer_ns();
                    if (t0 == 0) t0 = now;
                    if (__any_sync(0xffffffffu, now - t0 > 5000000000ull)) break;
                }
            }
            __threadfence();
        }
        __syncthreads();
    }
    DZO_DEVINL unsigned long long* count_cell() const { return base + (size_t)kRieszBarStride * kRieszBarStride; }
};
inline size_t riesz_bar_bytes() { return sizeof(unsigned long long) * ((size_t)kRieszBarStride * kRieszBarStride + 8); }

// NT = threads per CTA.  512 (default): 16 warps of up to 128 registers, so eight energy / four gradient pair terms run
// interleaved per lane -- one energy evaluation is 2112 warp items, i.e. at most ONE per warp on either CTA size, and what
// fills the FP64 pipe is chains in flight per warp, not warps (1024 threads: 64 registers, four / two terms).
template <int DIM, int NT>
struct RieszDev {
    static constexpr int W = NT / 32;                       // warps per CTA, each with a private staging buffer
    static constexpr int VT = 4096 / NT;                    // virtual threads of the canonical tree per thread
    static constexpr int EB = (NT == 512) ? 2 * kRieszBatch : kRieszBatch;           // energy pair terms in flight per lane
    static constexpr int GB = (NT == 512) ? 2 * kRieszGradBatch : kRieszGradBatch;   // gradient pair terms in flight per lane
    // trial point of point j: w = x_j + alpha*d_j, then constraint_function! (normalise) [pmode 0];
    // pmode 2: the stored point itself (already constrained).
    static DZO_DEVINL void trial_point(const RieszGdArgs& a, const double* dir, int j, double alpha, int pmode,
                                       double (&w)[DIM]) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) w[k] = a.x[(long long)j * DIM + k];
        if (pmode == 2) return;
#pragma unroll
        for (int k = 0; k < DIM; ++k) w[k] = w[k] + alpha * (a.dscale * dir[(long long)j * DIM + k]);   // legacy/Kernels.jl:127-135
        if (a.sphere) {                                                                    // [GLUE] SURVEY 8.0
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) s += w[k] * w[k];
            const double inv = 1.0 / sqrt(s);
#pragma unroll
            for (int k = 0; k < DIM; ++k) w[k] *= inv;
        }
    }

    // ---- riesz_energy (legacy/ExampleFunctions.jl:30-45), phase 1: segment partials
    // The warp that finishes the LAST item of a 32-row block (per-block counter, threadfence pattern) adds that
    // block's segment partials in ascending order and stores the row energies: no grid barrier between the two.
    // PAIRED evaluation (npr = 2): the items of two probes (alpha0, alpha1) share one phase.  One probe occupies fewer than
    // half of the grid's warps (2112 items for 4736 warps at N = 4096), so the second one rides in the idle warps; every
    // probe's arithmetic is what it would be alone.
    static DZO_DEVINL void energy_segments(const RieszGdArgs& a, const double* dir, double alpha0, double alpha1, int npr, int pmode,
                                           int par, double* wsm) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double* buf = wsm + warp * (DZO_RIESZ_SEG * DIM);
        const int nseg_all = (a.N + DZO_RIESZ_SEG - 1) / DZO_RIESZ_SEG;
        for (int idx2 = blockIdx.x + gridDim.x * warp; idx2 < npr * a.n_e_items; idx2 += gridDim.x * W) {
            const int pr = (npr == 2) ? (idx2 & 1) : 0;                 // probes interleaved: both get heavy items first
            const int idx = (npr == 2) ? (idx2 >> 1) : idx2;
            const double alpha = pr ? alpha1 : alpha0;
            double* segE = a.segE + (long long)pr * nseg_all * a.N;
            unsigned* rbcnt = a.rbcnt + pr * a.ecnt_stride;
            const int2 it = a.e_items[idx];
            const int i0 = it.y * DZO_RIESZ_SEG;
            const int cnt = min(DZO_RIESZ_SEG, a.N - i0);
            __syncwarp();
            for (int q = lane; q < cnt; q += 32) {
                double w[DIM];
                trial_point(a, dir, i0 + q, alpha, pmode, w);
#pragma unroll
                for (int k = 0; k < DIM; ++k) buf[q * DIM + k] = w[k];
            }
            __syncwarp();
            const int j = it.x * 32 + lane;
            if (j < a.N) {
                const int lim = min(cnt, j - i0);   // sources i < j inside this segment
                if (lim > 0) {
                    double wj[DIM];
                    trial_point(a, dir, j, alpha, pmode, wj);
                    double seg = 0.0;
                    // the sqrt / reciprocal chains of EB sources run interleaved (ieee_fast.cuh); only the
                    // adds are ordered
                    int i = 0;
                    for (; i + EB <= lim; i += EB) {
                        double ds[EB], t[EB];
                        bool safe = true;
#pragma unroll
                        for (int u = 0; u < EB; ++u) {
                            double dist_sq = 0.0;
#pragma unroll
                            for (int k = 0; k < DIM; ++k) {
                                const double dist = buf[(i + u) * DIM + k] - wj[k];   // points[k,i] - points[k,j]  :37
                                dist_sq += dist * dist;
                            }
                            ds[u] = dist_sq;
                            safe &= ieee_fast_safe(dist_sq);
                        }
                        if (safe) {
#pragma unroll
                            for (int u = 0; u < EB; ++u) t[u] = ieee_fast_rcp(ieee_fast_sqrt(ds[u]));
                        } else {
#pragma unroll
                            for (int u = 0; u < EB; ++u) t[u] = ieee_rsqrt_operators(ds[u]);
                        }
#pragma unroll
                        for (int u = 0; u < EB; ++u) seg += t[u];     // rsqrt(dist_sq)  :41
                    }
                    for (; i < lim; ++i) {
                        double dist_sq = 0.0;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            const double dist = buf[i * DIM + k] - wj[k];
                            dist_sq += dist * dist;
                        }
                        seg += ieee_rsqrt_operators(dist_sq);
                    }
                    segE[(long long)it.y * a.N + j] = seg;
                }
            }
            __threadfence();
            __syncwarp();
            unsigned last = 0;
            if (lane == 0) {
                const int jmax = min(a.N, it.x * 32 + 32) - 1;
                const unsigned expected = (unsigned)((jmax + DZO_RIESZ_SEG - 1) / DZO_RIESZ_SEG);
                last = (atomicAdd(&rbcnt[it.x], 1u) == expected - 1u);
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence();
                if (j < a.N) {
                    // segments below row j in ascending order, eight loads in flight at a time (see gradient_row)
                    const int ns = (j + DZO_RIESZ_SEG - 1) / DZO_RIESZ_SEG;
                    double ej = 0.0;
                    constexpr int CB = 8;
                    for (int s0 = 0; s0 < ns; s0 += CB) {
                        double v[CB];
#pragma unroll
                        for (int u = 0; u < CB; ++u) v[u] = (s0 + u < ns) ? __ldcg(&segE[(long long)(s0 + u) * a.N + j]) : 0.0;
#pragma unroll
                        for (int u = 0; u < CB; ++u)
                            if (s0 + u < ns) ej = (s0 + u == 0) ? v[u] : ej + v[u];
                    }
                    a.rowE[(long long)(2 * par + pr) * a.N + j] = ej;
                }
                if (lane == 0) rbcnt[it.x] = 0;        // next use is behind at least one grid barrier
            }
        }
    }

    // Two lanes per row (esplit = 2): an item is 16 rows x 128 sources; lane l < 16 takes the even sources of row l,
    // lane l + 16 the odd ones, EB terms each per round, and lane l adds the 2*EB values in source
    // order (the odd ones arrive by shuffle).  Twice as many items as with one lane per row, so nearly every warp of
    // the grid has one (2112 -> 4224 items for 4736 warps at N = 4096) and the serial chain per lane halves.
    static DZO_DEVINL void energy_segments_split(const RieszGdArgs& a, const double* dir, double alpha, int pmode, int par,
                                                 double* wsm) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int half = lane >> 4;
        double* buf = wsm + warp * (DZO_RIESZ_SEG * DIM);
        for (int idx = blockIdx.x + gridDim.x * warp; idx < a.n_e_items; idx += gridDim.x * W) {
            const int2 it = a.e_items[idx];                  // (16-row block, segment)
            const int i0 = it.y * DZO_RIESZ_SEG;
            const int cnt = min(DZO_RIESZ_SEG, a.N - i0);
            __syncwarp();
            for (int q = lane; q < cnt; q += 32) {
                double w[DIM];
                trial_point(a, dir, i0 + q, alpha, pmode, w);
#pragma unroll
                for (int k = 0; k < DIM; ++k) buf[q * DIM + k] = w[k];
            }
            __syncwarp();
            const int j = it.x * 16 + (lane & 15);
            const int lim = (j < a.N) ? max(0, min(cnt, j - i0)) : 0;   // sources i < j inside this segment
            const int lim_max = __reduce_max_sync(0xffffffffu, lim);
            double wj[DIM];
            if (j < a.N) trial_point(a, dir, j, alpha, pmode, wj);
            else {
#pragma unroll
                for (int k = 0; k < DIM; ++k) wj[k] = 0.0;
            }
            double seg = 0.0;
            for (int base = 0; base < lim_max; base += 2 * EB) {
                double ds[EB], t[EB];
                bool safe = true;
#pragma unroll
                for (int u = 0; u < EB; ++u) {
                    const int i = base + 2 * u + half;
                    double dist_sq = 0.0;
                    if (i < lim) {
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            const double dist = buf[i * DIM + k] - wj[k];         // points[k,i] - points[k,j]  :37
                            dist_sq += dist * dist;
                        }
                    } else {
                        dist_sq = 1.0;                                            // placeholder, never added
                    }
                    ds[u] = dist_sq;
                    safe &= ieee_fast_safe(dist_sq);
                }
                if (safe) {
#pragma unroll
                    for (int u = 0; u < EB; ++u) t[u] = ieee_fast_rcp(ieee_fast_sqrt(ds[u]));
                } else {
#pragma unroll
                    for (int u = 0; u < EB; ++u) t[u] = ieee_rsqrt_operators(ds[u]);
                }
#pragma unroll
                for (int u = 0; u < EB; ++u) {
                    const double odd = __shfl_down_sync(0xffffffffu, t[u], 16);
                    if (base + 2 * u < lim) seg += t[u];                          // rsqrt(dist_sq)  :41 (even source)
                    if (base + 2 * u + 1 < lim) seg += odd;                       //                     (odd source)
                }
            }
            if (half == 0 && lim > 0) a.segE[(long long)it.y * a.N + j] = seg;
            __threadfence();
            __syncwarp();
            unsigned last = 0;
            if (lane == 0) {
                const int jmax = min(a.N, it.x * 16 + 16) - 1;
                const unsigned expected = (unsigned)((jmax + DZO_RIESZ_SEG - 1) / DZO_RIESZ_SEG);
                last = (atomicAdd(&a.rbcnt[it.x], 1u) == expected - 1u);
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence();
                if (half == 0 && j < a.N) {
                    double ej = 0.0;
                    for (int s0 = 0; s0 < j; s0 += DZO_RIESZ_SEG) {
                        const double sg = __ldcg(&a.segE[(long long)(s0 / DZO_RIESZ_SEG) * a.N + j]);
                        ej = (s0 == 0) ? sg : ej + sg;
                    }
                    a.rowE[(long long)(2 * par) * a.N + j] = ej;
                }
                if (lane == 0) a.rbcnt[it.x] = 0;
            }
        }
    }

    // phase 2 (after ONE grid barrier): every CTA runs the canonical tree over the row energies itself (32 KB from L2
    // at N = 4096), so the value is identical on every thread of the grid without a broadcast or a second barrier.
    template <int NPR>
    static DZO_DEVINL void energy_finish(const RieszGdArgs& a, int par, double* sm, double (&f)[2]) {
        // row j -> virtual thread j mod 4096, ascending; all loads of a round of 4096 rows are issued before the first
        // add (one loop per accumulator made every load wait out its own L2 round trip)
        double p[NPR][VT];
#pragma unroll
        for (int pr = 0; pr < NPR; ++pr)
#pragma unroll
            for (int q = 0; q < VT; ++q) p[pr][q] = 0.0;
        for (int base = 0; base < a.N; base += DZO_TREE_WIDTH) {
            double v[NPR][VT];
#pragma unroll
            for (int pr = 0; pr < NPR; ++pr) {
                const double* rowE = a.rowE + (long long)(2 * par + pr) * a.N;
#pragma unroll
                for (int q = 0; q < VT; ++q) {
                    const int j = base + threadIdx.x + NT * q;
                    v[pr][q] = (j < a.N) ? __ldcg(&rowE[j]) : 0.0;
                }
            }
#pragma unroll
            for (int pr = 0; pr < NPR; ++pr)
#pragma unroll
                for (int q = 0; q < VT; ++q)
                    if (base + threadIdx.x + NT * q < a.N) p[pr][q] += v[pr][q];
        }
        double out[NPR];
        cta_tree_reduce<NPR, NT>(p, sm, out);
#pragma unroll
        for (int pr = 0; pr < NPR; ++pr) f[pr] = out[pr];
    }

    // `epar` counts the evaluation PHASES of this launch (identical on every CTA); its parity selects the rowE buffers, so
    // a CTA that is still reducing phase k cannot be overtaken by the row stores of phase k+1.
    static DZO_DEVINL double energy(const RieszGdArgs& a, RieszBar& grid, int& epar, const double* dir, double alpha, int pmode,
                                    double* wsm, double* sm) {
        const int par = epar & 1;
        epar += 1;
        riesz_prof_mark(a, 2);
        if (a.esplit == 2) energy_segments_split(a, dir, alpha, pmode, par, wsm);
        else energy_segments(a, dir, alpha, 0.0, 1, pmode, par, wsm);
        riesz_prof_mark(a, 3);
        grid.sync();
        riesz_prof_mark(a, 4);
        double f[2];
        energy_finish<1>(a, par, sm, f);
        riesz_prof_mark(a, 5);
        return f[0];
    }
    // f(alpha0) and f(alpha1) in one phase: one grid barrier and one (two-value) tree for both
    static DZO_DEVINL void energy_pair(const RieszGdArgs& a, RieszBar& grid, int& epar, const double* dir, double alpha0,
                                       double alpha1, double* wsm, double* sm, double& f0v, double& f1v) {
        const int par = epar & 1;
        epar += 1;
        riesz_prof_mark(a, 12);
        energy_segments(a, dir, alpha0, alpha1, 2, 0, par, wsm);
        riesz_prof_mark(a, 3);
        grid.sync();
        riesz_prof_mark(a, 4);
        double f[2];
        energy_finish<2>(a, par, sm, f);
        riesz_prof_mark(a, 5);
        f0v = f[0];
        f1v = f[1];
    }

    // ---- riesz_gradient! (:47-83) at the stored points + tangent projection (:361-374)
    // As for the energy, the warp finishing the last segment item of a 32-row block combines that block's rows
    // (ascending segment order), projects, and writes g (and dg = g_new - g_old when with_delta).
    // RP row blocks per item.  RP = 2 (an item stages its 128 sources once and runs two rows per lane side by side: 2048
    // items for the 2368 warps of 512-thread CTAs instead of 4096) was measured and is NOT used: a double item took 50 us
    // where two single items took 43 -- the pair loops are pipe-bound, not latency-bound, and the extra predicates cost
    // more than the shared staging saves (profiles/README.md).
    static DZO_DEVINL void gradient_segments(const RieszGdArgs& a, double* wsm, bool with_delta) {
        constexpr int RP = 1;
        constexpr int GBR = GB;                                      // sources per batch and row: RP * GBR chains per lane
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        double* buf = wsm + warp * (DZO_RIESZ_SEG * DIM);
        const int nseg = (a.N + DZO_RIESZ_SEG - 1) / DZO_RIESZ_SEG;
        const int nrb = (a.N + 31) / 32;
        const int nrbp = (nrb + RP - 1) / RP;
        for (int idx = blockIdx.x + gridDim.x * warp; idx < nrbp * nseg; idx += gridDim.x * W) {
            const int rbp = idx / nseg, s = idx - rbp * nseg;
            const int i0 = s * DZO_RIESZ_SEG;
            const int cnt = min(DZO_RIESZ_SEG, a.N - i0);
            __syncwarp();
            for (int q = lane; q < cnt; q += 32)
#pragma unroll
                for (int k = 0; k < DIM; ++k) buf[q * DIM + k] = a.x[(long long)(i0 + q) * DIM + k];
            __syncwarp();
            int j[RP];
            bool row[RP];
            double xj[RP][DIM], part[RP][DIM];
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                j[r] = (rbp * RP + r) * 32 + lane;
                row[r] = (rbp * RP + r) < nrb && j[r] < a.N;
#pragma unroll
                for (int k = 0; k < DIM; ++k) { xj[r][k] = row[r] ? a.x[(long long)j[r] * DIM + k] : 0.0; part[r][k] = 0.0; }
            }
            int i = 0;
            for (; i + GBR <= cnt; i += GBR) {
                double ds[RP][GBR], c[RP][GBR];
                bool safe = true;
#pragma unroll
                for (int r = 0; r < RP; ++r)
#pragma unroll
                    for (int u = 0; u < GBR; ++u) {
                        double dist_sq = 0.0;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            const double dist = buf[(i + u) * DIM + k] - xj[r][k];
                            dist_sq += dist * dist;
                        }
                        // the skipped self term (:55, :69) and the rows past the end ride along as 1.0
                        ds[r][u] = (!row[r] || i0 + i + u == j[r]) ? 1.0 : dist_sq;
                        safe &= ieee_fast_safe(ds[r][u]);
                    }
                if (safe) {
#pragma unroll
                    for (int r = 0; r < RP; ++r)
#pragma unroll
                        for (int u = 0; u < GBR; ++u) {
                            const double inv_dist = ieee_fast_rcp(ieee_fast_sqrt(ds[r][u]));   // :61
                            c[r][u] = ieee_fast_div(inv_dist, ds[r][u]);                          // :62
                        }
                } else {
#pragma unroll
                    for (int r = 0; r < RP; ++r)
#pragma unroll
                        for (int u = 0; u < GBR; ++u) c[r][u] = ieee_inv_cubed_operators(ds[r][u]);
                }
#pragma unroll
                for (int r = 0; r < RP; ++r)
#pragma unroll
                    for (int u = 0; u < GBR; ++u) {
                        if (!row[r] || i0 + i + u == j[r]) continue;           // :55, :69 (i != j)
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            const double dist = buf[(i + u) * DIM + k] - xj[r][k];
                            part[r][k] += dist * c[r][u];                      // :65
                        }
                    }
            }
            for (; i < cnt; ++i) {
#pragma unroll
                for (int r = 0; r < RP; ++r) {
                    if (!row[r] || i0 + i == j[r]) continue;
                    double dist_sq = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const double dist = buf[i * DIM + k] - xj[r][k];
                        dist_sq += dist * dist;
                    }
                    const double inv_dist_cubed = ieee_inv_cubed_operators(dist_sq);
#pragma unroll
                    for (int k = 0; k < DIM; ++k) {
                        const double dist = buf[i * DIM + k] - xj[r][k];
                        part[r][k] += dist * inv_dist_cubed;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RP; ++r)
                if (row[r])
#pragma unroll
                    for (int k = 0; k < DIM; ++k) a.segG[((long long)s * a.N + j[r]) * DIM + k] = part[r][k];
            __threadfence();
            __syncwarp();
#pragma unroll
            for (int r = 0; r < RP; ++r) {
                const int rb = rbp * RP + r;
                if (rb >= nrb) continue;                                       // uniform over the warp
                unsigned last = 0;
                if (lane == 0) last = (atomicAdd(&a.rbcnt[a.gcnt_off + rb], 1u) == (unsigned)nseg - 1u);
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    __threadfence();
                    if (row[r]) gradient_row(a, j[r], nseg, with_delta);
                    if (lane == 0) a.rbcnt[a.gcnt_off + rb] = 0;
                }
            }
        }
    }
    // ---- riesz_gradient!, symmetric CTA tiles (gvariant 1).  The pair weight inv_dist_cubed(i, j) is the same number
    // for row j / source i and for row i / source j (dist_sq is a sum of squares of exactly negated differences), and the
    // products dist * weight of the two uses are exact negations of each other.  A CTA therefore computes the weights of
    // a whole (segment a) x (segment b) tile ONCE into shared memory -- 16 per thread, 4 at a time through the interleaved
    // IEEE fast paths -- and then BOTH families of segment partials that need them: rows of a over the sources of b and
    // rows of b over the sources of a, each lane walking its 128 sources in ascending order (the reference order of
    // :55-81 inside a segment), one warp per (32 rows, component).  38 -> ~24 FP64 instructions per ordered pair term.
    static DZO_DEVINL void gradient_tiles(const RieszGdArgs& a, double* wsm, bool with_delta) {
        constexpr int LD = DZO_RIESZ_SEG + 1;                       // padded leading dimension (conflict-free columns)
        double* Cw = wsm;                                            // [128][129] pair weights
        double* PA = wsm + DZO_RIESZ_SEG * LD;                       // [128][DIM] points of segment a
        double* PB = PA + DZO_RIESZ_SEG * DIM;                       // [128][DIM] points of segment b
        __shared__ int s_done[64];                                  // at most 8 per tile, 528 tiles / 148 CTAs at N = 4096
        __shared__ int s_ndone;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int nseg = (a.N + DZO_RIESZ_SEG - 1) / DZO_RIESZ_SEG;
        if (threadIdx.x == 0) s_ndone = 0;
        __syncthreads();
        for (int job = blockIdx.x; job < a.n_g_jobs; job += gridDim.x) {
            const int2 t = a.g_jobs[job];
            const int ta = t.x, tb = t.y;
            const int a0 = ta * DZO_RIESZ_SEG, b0 = tb * DZO_RIESZ_SEG;
            const int na = min(DZO_RIESZ_SEG, a.N - a0), nb = min(DZO_RIESZ_SEG, a.N - b0);
            __syncthreads();                                         // the previous job's readers are done with the tile
            for (int e = threadIdx.x; e < na * DIM; e += 1024) PA[e] = a.x[(long long)a0 * DIM + e];
            for (int e = threadIdx.x; e < nb * DIM; e += 1024) PB[e] = a.x[(long long)b0 * DIM + e];
            __syncthreads();
            // weights: warp w owns tile rows 4w .. 4w+3, lane the columns lane, lane+32, lane+64, lane+96
#pragma unroll 1
            for (int r = 0; r < 4; ++r) {
                const int jj = 4 * warp + r;
                if (jj >= na) break;                                 // uniform over the warp
                double pj[DIM];
#pragma unroll
                for (int k = 0; k < DIM; ++k) pj[k] = PA[jj * DIM + k];
                double ds[4], c[4];
                bool safe = true;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int ii = lane + 32 * u;
                    double dist_sq = 1.0;                            // placeholder for the self pair and for columns past nb
                    if (ii < nb && !(ta == tb && ii == jj)) {
                        dist_sq = 0.0;
#pragma unroll
                        for (int k = 0; k < DIM; ++k) {
                            const double dist = PB[ii * DIM + k] - pj[k];
                            dist_sq += dist * dist;
                        }
                    }
                    ds[u] = dist_sq;
                    safe &= ieee_fast_safe(dist_sq);
                }
                if (safe) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[u] = ieee_fast_div(ieee_fast_rcp(ieee_fast_sqrt(ds[u])), ds[u]);   // :61-62
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[u] = ieee_inv_cubed_operators(ds[u]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) Cw[jj * LD + lane + 32 * u] = c[u];
            }
            __syncthreads();
            // segment partials: warps [0, 4*DIM) rows of a over the sources of b, warps [4*DIM, 8*DIM) rows of b over a
            if (warp < 8 * DIM) {
                const int side = warp / (4 * DIM), q = (warp % (4 * DIM)) / DIM, k = warp % DIM;
                const int row = 32 * q + lane;
                if (side == 0) {
                    if (row < na) {
                        const double xr = PA[row * DIM + k];
                        const int self = (ta == tb) ? row : -1;
                        double part = 0.0;
                        int ii = 0;
                        for (; ii + 8 <= nb; ii += 8) {              // 16 shared-memory loads in flight, adds in order
                            double term[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) term[u] = (PB[(ii + u) * DIM + k] - xr) * Cw[row * LD + ii + u];   // :65
#pragma unroll
                            for (int u = 0; u < 8; ++u)
                                if (ii + u != self) part += term[u];                   // :55, :69 (i != j)
                        }
                        for (; ii < nb; ++ii)
                            if (ii != self) part += (PB[ii * DIM + k] - xr) * Cw[row * LD + ii];
                        a.segG[((long long)tb * a.N + (a0 + row)) * DIM + k] = part;
                    }
                } else if (ta != tb) {
                    if (row < nb) {
                        const double xr = PB[row * DIM + k];
                        double part = 0.0;
                        int jj = 0;
                        for (; jj + 8 <= na; jj += 8) {
                            double term[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) term[u] = (PA[(jj + u) * DIM + k] - xr) * Cw[(jj + u) * LD + row];
#pragma unroll
                            for (int u = 0; u < 8; ++u) part += term[u];
                        }
                        for (; jj < na; ++jj) part += (PA[jj * DIM + k] - xr) * Cw[jj * LD + row];
                        a.segG[((long long)ta * a.N + (b0 + row)) * DIM + k] = part;
                    }
                }
            }
            if (warp < 8 * DIM) __threadfence();                    // the writers of segG
            __syncthreads();
            if (threadIdx.x < 8) {
                const int side = threadIdx.x >> 2, q = threadIdx.x & 3;
                const bool valid = (side == 0) ? (32 * q < na) : (ta != tb && 32 * q < nb);
                if (valid) {
                    const int rb = (side == 0 ? ta : tb) * (DZO_RIESZ_SEG / 32) + q;
                    if (atomicAdd(&a.rbcnt[a.gcnt_off + rb], 1u) == (unsigned)nseg - 1u) {
                        a.rbcnt[a.gcnt_off + rb] = 0;
                        s_done[atomicAdd(&s_ndone, 1)] = rb;        // this job completed the row block: combine its rows later
                    }
                }
            }
            __syncthreads();
            if (s_ndone > 56 || job + (int)gridDim.x >= a.n_g_jobs) {
                // combine the row blocks this CTA completed, one warp each -- after its last tile, so that no tile waits
                // for a combination (or earlier when the list is about to overflow: 8 entries per tile at most)
                const int ndone = s_ndone;
                for (int e = warp; e < ndone; e += 32) {
                    __threadfence();
                    const int j = s_done[e] * 32 + lane;
                    if (j < a.N) gradient_row(a, j, nseg, with_delta);
                }
                __syncthreads();
                if (threadIdx.x == 0) s_ndone = 0;
            }
        }
    }

    // one row: combine, project, write g (and dg = g_new - g_old when with_delta)
    static DZO_DEVINL void gradient_row(const RieszGdArgs& a, int j, int nseg, bool with_delta) {
        double acc[DIM], xj[DIM];
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            acc[k] = __ldcg(&a.segG[((long long)0 * a.N + j) * DIM + k]);
            xj[k] = a.x[(long long)j * DIM + k];
        }
        // segments in ascending order; eight segments' partials are requested before the first add (one loop iteration
        // per segment made the warp that finishes a row block wait out an L2 round trip per segment -- 32 of them at
        // N = 4096 -- while the grid barrier waited for it)
        constexpr int CB = 8;
        for (int s0 = 1; s0 < nseg; s0 += CB) {
            double v[CB][DIM];
#pragma unroll
            for (int u = 0; u < CB; ++u)
#pragma unroll
                for (int k = 0; k < DIM; ++k)
                    v[u][k] = (s0 + u < nseg) ? __ldcg(&a.segG[((long long)(s0 + u) * a.N + j) * DIM + k]) : 0.0;
#pragma unroll
            for (int u = 0; u < CB; ++u)
                if (s0 + u < nseg)
#pragma unroll
                    for (int k = 0; k < DIM; ++k) acc[k] += v[u][k];
        }
        if (a.sphere) {
            double overlap = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) overlap += xj[k] * acc[k];
#pragma unroll
            for (int k = 0; k < DIM; ++k) acc[k] -= overlap * xj[k];
        }
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
            const long long e = (long long)j * DIM + k;
            if (with_delta) a.dg[e] = acc[k] - a.g[e];                         // :433-435
            a.g[e] = acc[k];
        }
    }

    static DZO_DEVINL void gradient(const RieszGdArgs& a, double* wsm, bool with_delta) {
        if constexpr (NT == 1024) {
            if (a.gvariant == 1) { gradient_tiles(a, wsm, with_delta); return; }
        }
        gradient_segments(a, wsm, with_delta);
    }

    // every CTA scans all points (N/1024 per thread): cheap, avoids a grid barrier
    static DZO_DEVINL bool all_points(const RieszGdArgs& a, const double* dir, double alpha, double alpha_ref, int what) {
        int bad = 0;
        for (int j = threadIdx.x; j < a.N; j += NT) {
            if (what == 0) {            // step_is_zero: all(dir == 0)  :71-85
#pragma unroll
                for (int k = 0; k < DIM; ++k) bad |= !(a.dscale * dir[(long long)j * DIM + k] == 0.0);
            } else if (what == 1) {     // !point_changed: all(x == x + alpha*dir) before the constraint  :73-80
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    const double xx = a.x[(long long)j * DIM + k];
                    bad |= (xx != xx + alpha * (a.dscale * dir[(long long)j * DIM + k]));
                }
            } else if (what == 2) {     // new_point == initial_point after the constraint  :119
                double w[DIM];
                trial_point(a, dir, j, alpha, 0, w);
#pragma unroll
                for (int k = 0; k < DIM; ++k) bad |= !(a.x[(long long)j * DIM + k] == w[k]);
            } else {                    // new_point == reference_point  :150
                double w[DIM], r[DIM];
                trial_point(a, dir, j, alpha, 0, w);
                trial_point(a, dir, j, alpha_ref, 0, r);
#pragma unroll
                for (int k = 0; k < DIM; ++k) bad |= !(w[k] == r[k]);
            }
        }
        return __syncthreads_or(bad) == 0;
    }

    // step_is_zero (all(dir == 0), :71-85) and !point_changed (all(x == x + alpha*dir), :73-80) from one set of loads
    static DZO_DEVINL void first_checks(const RieszGdArgs& a, const double* dir, double alpha, bool& zero_step, bool& unchanged) {
        int nonzero = 0, moved = 0;
        for (int j = threadIdx.x; j < a.N; j += NT) {
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const double dd = a.dscale * dir[(long long)j * DIM + k];
                const double xx = a.x[(long long)j * DIM + k];
                nonzero |= !(dd == 0.0);
                moved |= (xx != xx + alpha * dd);
            }
        }
        zero_step = __syncthreads_or(nonzero) == 0;
        unchanged = __syncthreads_or(moved) == 0;
    }

    // QuadraticLineSearch (:191-216) over find_three_point_bracket (:49-172), first trial step t1
    static DZO_DEVINL void line_search(const RieszGdArgs& a, RieszBar& grid, int& epar, const double* dir, double f0, double t1,
                                       double sign, double* wsm, double* sm, double& t_best, double& f_best,
                                       long long& evals) {
        double x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0;
        do {
            if (!isfinite(f0)) break;                                          // :64-66
            if (!isfinite(t1) || t1 == 0.0) break;                             // [GLUE]
            double step = t1;
            bool small = false;
            bool zero_step, unchanged;                                         // :71-85 and :73-80 in ONE pass over the points
            first_checks(a, dir, sign * step, zero_step, unchanged);
            if (zero_step) break;
            int cap = DZO_LINESEARCH_CAP;
            bool capped = false;
            while (unchanged) {                                                // :91-101
                step += step;
                small = true;
                unchanged = all_points(a, dir, sign * step, 0.0, 1);
                if (--cap == 0) { capped = true; break; }
            }
            if (capped) break;
            if (small && all_points(a, dir, sign * step, 0.0, 2)) break;       // :107-123
            // probe(t, guess): f at t.  With paired evaluation the probe the search will most likely ask for next
            // (`guess`: the doubled step while expanding -- 9 searches of 10 on config 5 --, the halved one while
            // shrinking) is evaluated in the same phase and kept; asking for it later costs nothing.  Which values the
            // search sees, and in which order, is unchanged.
            const bool paired = (a.espec != 0) && (a.esplit != 2);
            double spec_t = 0.0, spec_f = 0.0;
            bool spec_valid = false;
            auto probe = [&](double t, double guess) -> double {
                ++evals;
                if (spec_valid && spec_t == t) { spec_valid = false; return spec_f; }
                if (!paired) return energy(a, grid, epar, dir, sign * t, 0, wsm, sm);
                double ft;
                energy_pair(a, grid, epar, dir, sign * t, sign * guess, wsm, sm, ft, spec_f);
                spec_t = guess;
                spec_valid = true;
                return ft;
            };
            double fa = probe(step, step + step);                              // :126
            if (fa <= f0) {                                                    // :130
                int num_increases = 0;
                cap = DZO_LINESEARCH_CAP;
                for (;;) {                                                     // :143-156
                    const double ds = step + step;
                    num_increases += 1;
                    const double fb = probe(ds, ds + ds);
                    --cap;
                    if (((a.max_increases > 0) && (num_increases >= a.max_increases)) || !isfinite(fb) || fb > fa ||
                        all_points(a, dir, sign * ds, sign * step, 3) || cap == 0) {
                        x1 = step; f1 = fa; x2 = ds; f2 = fb;
                        break;
                    }
                    step = ds;
                    fa = fb;
                }
            } else {                                                           // :157-171
                cap = DZO_LINESEARCH_CAP;
                for (;;) {
                    const double hs = 0.5 * step;
                    const double fb = probe(hs, 0.5 * hs);
                    --cap;
                    if (fb <= f0 || cap == 0) {
                        x1 = hs; f1 = fb; x2 = step; f2 = fa;
                        break;
                    }
                    step = hs;
                    fa = fb;
                }
            }
        } while (0);
        double xb = 0.0, fb = f0;                                              // :196-202
        if (f1 < fb) { xb = x1; fb = f1; }
        if (f2 < fb) { xb = x2; fb = f2; }
        const double delta_1 = f0 - f1;                                        // :203-205
        const double delta_2 = f2 - f1;
        const double sum_deltas = delta_1 + delta_2;
        if (delta_1 >= 0.0 && delta_2 >= 0.0 && sum_deltas > 0.0) {            // :206-214
            const double twice_delta_1 = delta_1 + delta_1;
            const double delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas);
            const double xq = delta_ratio * x1;
            const double fq = energy(a, grid, epar, dir, sign * xq, 0, wsm, sm);
            ++evals;
            if (fq < fb) { xb = xq; fb = fq; }
        }
        t_best = xb;
        f_best = fb;
    }
};

// dx . dx and g . g in one pass over the canonical tree (every CTA computes both itself).  The loop runs over rounds
// of 4096 pairs with the four virtual threads of a thread side by side, so the 16 loads of a round are in flight
// together (two L2 round trips at n = 12288 instead of six); each accumulator still sees its pairs in ascending order.
template <int NT>
DZO_DEVINL void cta_tree_norms2(const double* __restrict__ v, const double* __restrict__ w, long long n, double* sm,
                                double& vv, double& ww) {
    constexpr int VT = 4096 / NT;
    double p[2][VT];
#pragma unroll
    for (int q = 0; q < VT; ++q) { p[0][q] = 0.0; p[1][q] = 0.0; }
    if ((n & 1) == 0) {
        // every pair as ONE 16-byte load and all of a thread's loads of a round in flight together: one L2 round trip per
        // 4096 pairs (config 5, n = 12288: two) -- the phase log had this pass at 14 us of a 166 us step with 8-byte loads
        const double2* v2 = reinterpret_cast<const double2*>(v);
        const double2* w2 = reinterpret_cast<const double2*>(w);
        const long long m = n >> 1;
        for (long long base = 0; base < m; base += DZO_TREE_WIDTH) {
            double2 a[VT], b[VT];
#pragma unroll
            for (int q = 0; q < VT; ++q) {
                const long long k = base + threadIdx.x + NT * q;
                const bool in = k < m;
                a[q] = in ? __ldcg(&v2[k]) : make_double2(0.0, 0.0);
                b[q] = in ? __ldcg(&w2[k]) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < VT; ++q) {
                const long long k = base + threadIdx.x + NT * q;
                if (k < m) {
                    p[0][q] += a[q].x * a[q].x; p[1][q] += b[q].x * b[q].x;
                    p[0][q] += a[q].y * a[q].y; p[1][q] += b[q].y * b[q].y;
                }
            }
        }
    } else {
        for (long long base = 0; 2 * base < n; base += DZO_TREE_WIDTH) {
#pragma unroll
            for (int q = 0; q < VT; ++q) {
                const long long k = base + threadIdx.x + NT * q;
                if (2 * k < n) { p[0][q] += v[2 * k] * v[2 * k]; p[1][q] += w[2 * k] * w[2 * k]; }
                if (2 * k + 1 < n) { p[0][q] += v[2 * k + 1] * v[2 * k + 1]; p[1][q] += w[2 * k + 1] * w[2 * k + 1]; }
            }
        }
    }
    double out[2];
    cta_tree_reduce<2, NT>(p, sm, out);
    vv = out[0];
    ww = out[1];
}
// v . w over the canonical tree on one CTA of NT threads (cta_tree_dot of large_bfgs.cuh is the 1024-thread case)
template <int NT>
DZO_DEVINL double cta_tree_dot_nt(const double* __restrict__ v, const double* __restrict__ w, long long n, double* sm) {
    constexpr int VT = 4096 / NT;
    double p[1][VT];
#pragma unroll
    for (int q = 0; q < VT; ++q) {
        double acc = 0.0;
        for (long long k = threadIdx.x + NT * q; 2 * k < n; k += DZO_TREE_WIDTH) {
            acc += v[2 * k] * w[2 * k];
            if (2 * k + 1 < n) acc += v[2 * k + 1] * w[2 * k + 1];
        }
        p[0][q] = acc;
    }
    double out[1];
    cta_tree_reduce<1, NT>(p, sm, out);
    return out[0];
}

template <int DIM, int NT>
static __global__ void __launch_bounds__(NT, 1) riesz_gd_kernel(RieszGdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);          // 2 x 132 doubles: tree reduction scratch
    double* wsm = sm + 272;                                    // (NT / 32) warps x 128 x DIM staging
    // mode 0 (k step! calls) synchronises through the flag-word barrier when the host provides its buffer; every other
    // mode keeps grid.sync().  Every CTA reads the running count before its first barrier; the leader stores it back at
    // the end (it cannot get there before every CTA has arrived at every barrier, i.e. has read the count).
    RieszBar grid{cg::this_grid(), (a.mode == 0 && (int)gridDim.x <= kRieszBarStride) ? a.bar : nullptr, 0ull, (int)gridDim.x, (int)blockIdx.x};
    if (grid.base != nullptr) grid.seq = __ldcg(grid.count_cell());
    using R = RieszDev<DIM, NT>;
    const long long n = (long long)a.N * DIM;
    const long long gtid = (long long)blockIdx.x * NT + threadIdx.x, gsize = (long long)gridDim.x * NT;
    const bool leader = (blockIdx.x == 0 && threadIdx.x == 0);
    int epar = 0;                                               // energy evaluations of this launch (rowE parity)

    if (a.mode == 2) {                                          // dzo_dev_objective: f(x) as given
        const double f = R::energy(a, grid, epar, a.x, 0.0, 2, wsm, sm);
        if (leader) a.fbox[1] = f;
        return;
    }
    if (a.mode == 3) {                                          // dzo_dev_gradient
        R::gradient(a, wsm, false);
        return;
    }
    if (a.mode == 4) {                                          // dzo_dev_line_search
        double tb, fb;
        long long ev = 0;
        R::line_search(a, grid, epar, a.d, a.ls_f0, a.ls_t1, a.ls_sign, wsm, sm, tb, fb, ev);
        if (leader) { a.fbox[1] = tb; a.fbox[2] = fb; }
        return;
    }
    if (a.mode == 6) {
        // BFGSOptimizer(f, g!, c!, x0, L0)  legacy/DZOptimization.jl:762-810 with the Riesz objective
        if (a.sphere)
            for (int j = (int)gtid; j < a.N; j += (int)gsize) {                // :770 constraint_function!(x)
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) s += a.x[(long long)j * DIM + k] * a.x[(long long)j * DIM + k];
                const double inv = 1.0 / sqrt(s);
#pragma unroll
                for (int k = 0; k < DIM; ++k) a.x[(long long)j * DIM + k] *= inv;
            }
        for (long long e = gtid; e < n; e += gsize) { a.dx[e] = 0.0; a.dg[e] = 0.0; }   // :777-778
        grid.sync();
        const double f0 = R::energy(a, grid, epar, a.x, 0.0, 2, wsm, sm);      // :772
        R::gradient(a, wsm, false);                                   // :775-776
        grid.sync();
        for (long long e = gtid; e < n; e += gsize) a.d[e] = a.g[e];           // :784 next_step_direction = copy(gradient)
        if (leader) {
            LargeCtrl c;
            c.f = f0; c.L = a.initial_step_length; c.iter = 0; c.type = DZO_STEP_NULL; c.term = 0;
            c.kind = DZO_STEP_NULL; c.pad = 0; c.step_length = 0.0; c.overlap = 0.0; c.delta_norm = 0.0;
            c.evals = 1; c.calls = 0;
            for (int i = 0; i < 64; ++i) c.kind_log[i] = 0;
            *a.bctrl = c;
        }
        return;
    }
    if (a.mode == 5) {
        // step!(::BFGSOptimizer) :891-950 and :873-874 with the Riesz objective: the cooperative grid plays the
        // role cluster_bfgs_search_kernel plays for Rosenbrock; the n^2 sweeps that follow are the same kernels.
        const int term = __ldcg(&a.bctrl->term);
        if (term) {                                                            // :893
            if (leader) a.bctrl->kind = DZO_STEP_NULL;
            return;
        }
        const double f0 = __ldcg(&a.bctrl->f);
        const double step_length = __ldcg(&a.bctrl->L);                        // :918
        const long long iter0 = __ldcg(&a.bctrl->iter), calls0 = __ldcg(&a.bctrl->calls), evals0 = __ldcg(&a.bctrl->evals);
        long long evals = 0;
        grid.sync();                       // every CTA holds the control block before the leader may rewrite it
        double gg, dd;
        cta_tree_norms2<NT>(a.g, a.d, n, sm, gg, dd);
        const double grad_norm = sqrt(gg);                                     // :921
        const double bfgs_norm = sqrt(dd);                                     // :928
        double grad_step_length, grad_obj, bfgs_step_length, bfgs_obj;
        R::line_search(a, grid, epar, a.g, f0, step_length / grad_norm, -1.0, wsm, sm, grad_step_length, grad_obj, evals);  // :922-925
        R::line_search(a, grid, epar, a.d, f0, step_length / bfgs_norm, -1.0, wsm, sm, bfgs_step_length, bfgs_obj, evals);  // :929-932
        int kind;
        double alpha, fnew, Lnew;
        if (bfgs_obj < f0 && !(bfgs_obj > grad_obj)) {                         // :934
            kind = DZO_STEP_BFGS; alpha = -bfgs_step_length; fnew = bfgs_obj; Lnew = bfgs_step_length * bfgs_norm;
        } else if (grad_obj < f0) {                                            // :962
            kind = DZO_STEP_GRADIENT_DESCENT; alpha = -grad_step_length; fnew = grad_obj; Lnew = grad_step_length * grad_norm;
        } else {
            if (leader) {                                                      // :989
                a.bctrl->term = 1;
                a.bctrl->kind = DZO_STEP_NULL;
                a.bctrl->evals = evals0 + evals;
                a.bctrl->kind_log[calls0 & 63] = DZO_STEP_NULL;
                a.bctrl->calls = calls0 + 1;
            }
            return;
        }
        const double* dir = (kind == DZO_STEP_BFGS) ? a.d : a.g;
        for (int j = (int)gtid; j < a.N; j += (int)gsize) {                    // :943-947 / :971-975
            double w[DIM];
            R::trial_point(a, dir, j, alpha, 0, w);                            // x += alpha*dir ; constraint!(x)
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const long long e = (long long)j * DIM + k;
                a.dx[e] = w[k] - a.x[e];                                       // (-x_old) + x_new  :943, :949
            }
#pragma unroll
            for (int k = 0; k < DIM; ++k) a.x[(long long)j * DIM + k] = w[k];
        }
        grid.sync();                                   // every CTA is done reading the old gradient as `dir`
        R::gradient(a, wsm, true);                                    // :948, dg = (-g_old) + g_new  :944, :950
        grid.sync();
        double overlap = 0.0;
        if (kind == DZO_STEP_BFGS) {
            overlap = cta_tree_dot_nt<NT>(a.d, a.dg, n, sm);                          // :873
            const double inv_overlap = 1.0 / overlap;                          // :874
            for (long long e = gtid; e < n; e += gsize) a.sd[e] = a.d[e] * inv_overlap;
        } else {
            for (long long e = gtid; e < n; e += gsize) a.d[e] = a.g[e];       // :984-986
        }
        if (leader) {
            LargeCtrl* c = a.bctrl;
            c->f = fnew; c->L = Lnew; c->type = kind; c->iter = iter0 + 1;     // :937-940 / :965-968
            c->kind = kind; c->step_length = alpha; c->overlap = overlap; c->delta_norm = 0.0;
            c->evals = evals0 + evals;
            c->kind_log[calls0 & 63] = (unsigned char)kind;
            c->calls = calls0 + 1;
        }
        return;
    }
    if (a.mode == 1) {
        // GradientDescentOptimizer(...)  legacy/DZOptimization.jl:330-374
        if (a.sphere)
            for (int j = (int)gtid; j < a.N; j += (int)gsize) {                // :340 constraint_function!(x)
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) s += a.x[(long long)j * DIM + k] * a.x[(long long)j * DIM + k];
                const double inv = 1.0 / sqrt(s);
#pragma unroll
                for (int k = 0; k < DIM; ++k) a.x[(long long)j * DIM + k] *= inv;
            }
        for (long long e = gtid; e < n; e += gsize) { a.dx[e] = 0.0; a.dg[e] = 0.0; a.d[e] = 0.0; }
        grid.sync();
        const double f0 = R::energy(a, grid, epar, a.x, 0.0, 2, wsm, sm);      // :343
        R::gradient(a, wsm, false);                                   // :347-348
        grid.sync();
        const double inv_gradient_norm = 1.0 / sqrt(cta_tree_dot_nt<NT>(a.g, a.g, n, sm));   // :352
        if (isfinite(inv_gradient_norm)) {                                     // :354-357
            const double alpha = -a.initial_step_length * inv_gradient_norm;
            for (long long e = gtid; e < n; e += gsize) a.d[e] = a.g[e] * alpha;
        }
        if (leader) {
            GdCtrl c;
            c.f = f0; c.df = 0.0; c.L = 0.0; c.iter = 0; c.pad = 0; c.evals = 1;
            c.term = (!isfinite(f0)) || (!isfinite(inv_gradient_norm));        // :364-366
            *a.ctrl = c;
        }
        return;
    }

    // ---- k step! calls  (:393-449).  The scalars a step hands to the next (f, iteration count, has_terminated and the
    // factor alpha of next_step_direction = alpha * gradient) are identical on every CTA -- they all come out of
    // grid-wide reductions each CTA finishes itself -- so they travel in registers and the steps of one launch need no
    // barrier between them: the next line search reads its direction as alpha * g[e] (the very product stored in d[e]).
    int term = __ldcg(&a.ctrl->term);
    double f0 = __ldcg(&a.ctrl->f);
    long long iter = __ldcg(&a.ctrl->iter);
    long long evals_total = __ldcg(&a.ctrl->evals);
    const double* dir = a.d;
    for (int s = 0; s < a.ksteps; ++s) {
        if (term) break;                                                       // :402
        long long evals = 0;
        double step_size, objective_value;
        riesz_prof_mark(a, 1);
        R::line_search(a, grid, epar, dir, f0, 1.0, 1.0, wsm, sm, step_size, objective_value, evals);   // :405-407
        riesz_prof_mark(a, 6);
        if (step_size == 0.0 || !(objective_value < f0)) {                     // :410-414
            term = 1;
            if (leader) a.ctrl->term = 1;
            break;
        }
        for (int j = (int)gtid; j < a.N; j += (int)gsize) {                    // :418-423
            double w[DIM];
            R::trial_point(a, dir, j, step_size, 0, w);                        // x += step*d ; constraint!(x)
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const long long e = (long long)j * DIM + k;
                a.dx[e] = w[k] - a.x[e];                                       // delta!(dx, x_new, x_old)
            }
#pragma unroll
            for (int k = 0; k < DIM; ++k) a.x[(long long)j * DIM + k] = w[k];
        }
        grid.sync();
        riesz_prof_mark(a, 7);
        R::gradient(a, wsm, true);                                    // :433-435
        riesz_prof_mark(a, 8);
        grid.sync();
        riesz_prof_mark(a, 10);
        double dxdx, gg;
        cta_tree_norms2<NT>(a.dx, a.g, n, sm, dxdx, gg);
        riesz_prof_mark(a, 13);
        const double step_length = sqrt(dxdx);                                 // :424
        const double inv_gradient_norm = 1.0 / sqrt(gg);                       // :438
        const bool ok = isfinite(inv_gradient_norm);
        const double alpha = -step_length * inv_gradient_norm;                 // :445-446
        if (ok)
            for (long long e = gtid; e < n; e += gsize) a.d[e] = alpha * a.g[e];
        if (leader) {
            a.ctrl->iter = iter + 1;                                           // :415
            a.ctrl->L = step_length;                                           // :425
            a.ctrl->df = objective_value - f0;                                 // :428-429
            a.ctrl->f = objective_value;                                       // :430
            a.ctrl->evals = evals_total + evals;
            if (!ok) a.ctrl->term = 1;                                         // :439-442
        }
        evals_total += evals;
        iter += 1;
        f0 = objective_value;
        if (!ok) term = 1;
        dir = a.g;                         // d[e] == alpha * g[e] bit for bit; g is complete behind the barrier above
        a.dscale = alpha;
        riesz_prof_mark(a, 11);
    }
    if (grid.base != nullptr && leader) *grid.count_cell() = grid.seq;
}

inline size_t riesz_gd_smem(int dim, int nt) {
    const size_t staging = (size_t)(nt / 32) * DZO_RIESZ_SEG * dim;                                        // per-warp source staging
    const size_t tile = (nt == 1024) ? (size_t)DZO_RIESZ_SEG * (DZO_RIESZ_SEG + 1) + 2 * (size_t)DZO_RIESZ_SEG * dim : 0;   // gradient tile + 2 point sets (gvariant 1)
    return sizeof(double) * (272 + (staging > tile ? staging : tile));
}

}  // namespace dzo
