// warp_search.cuh -- the O(n) stage of step! for MEDIUM n (32 < n <= 512) on ONE WARP per problem.
//
// cluster_search.cuh spreads the 4096 virtual threads of the canonical tree over an 8-CTA cluster: right for
// n = 16384, wasteful for n = 64, where 32 pairs occupy one virtual warp and the other 127 hold +0.0 -- a batch of
// 4096 such problems paid 4096 clusters x ~30 cluster barriers each (4.3 ms per step! call, profiles/README.md).
// Here one physical warp emulates the VW = n / 64 (rounded up to a power of two) virtual warps that hold data:
//     virtual thread v = 32*vw + lane owns pair v (n <= 8192 means at most one pair per virtual thread)
//     bits 4..0   xor-shuffle butterfly 16,8,4,2,1 per virtual warp
//     bits 5..    the VW warp values combined in ascending bit order, in registers
//     the rest    every remaining level up to bit 11 adds the +0.0 of an empty subtree: one `+ 0.0`
// which is the oracle's DZO_ORDER_TREE bit for bit (tests/test_gpu_bfgs.py::test_batched_medium_n_trace and the
// single-handle traces at n = 34 ... 512).  The three vectors a probe touches live in registers; no shared memory,
// no barrier of any kind.  A CTA is four independent warps = four problems.
//
// legacy/DZOptimization.jl:891-950 (step! up to the Hessian update), :873-874, :876; line search :49-216.
#pragma once
#include "large_bfgs.cuh"

namespace dzo {

constexpr int kWarpSearchWarps = 4;      // problems per CTA
constexpr int kWarpSearchMaxN = 512;     // VW <= 8: 48 doubles of vector state per lane

template <int VW>
DZO_DEVINL double warp_team_sum(const double (&acc)[VW]) {
    double s[VW];
#pragma unroll
    for (int w = 0; w < VW; ++w) s[w] = warp_butterfly_desc(acc[w]);          // bits 4,3,2,1,0
#pragma unroll
    for (int w = 1; w < VW; w <<= 1)                                          // bits 5, 6, ... ascending
#pragma unroll
        for (int i = 0; i < VW; i += 2 * w) s[i] = s[i] + s[i + w];
    return s[0] + 0.0;                                                        // empty subtrees of the remaining levels
}

template <int VW>
struct WarpTeam {
    double2 X[VW], G[VW], D[VW];      // pair 32*w + lane of current_point, current_gradient, next_step_direction
    bool has[VW];
    int lane;

    // f(x + alpha*dir); MODE 1 also compares the trial point with the one of alpha_ref (:136, :150)
    template <int MODE>
    DZO_DEVINL double probe(const double2 (&Dir)[VW], double alpha, double alpha_ref, bool& changed, bool& same_ref) const {
        double acc[VW];
        unsigned fl = 0;
#pragma unroll
        for (int w = 0; w < VW; ++w) {
            acc[w] = 0.0;
            if (has[w]) {
                const double2 xx = X[w], dd = Dir[w];
                const double w0 = xx.x + alpha * dd.x;
                const double w1 = xx.y + alpha * dd.y;
                if ((xx.x != w0) | (xx.y != w1)) fl |= 1u;
                if (MODE == 1) {
                    const double r0 = xx.x + alpha_ref * dd.x;
                    const double r1 = xx.y + alpha_ref * dd.y;
                    if ((!(w0 == r0)) | (!(w1 == r1))) fl |= 2u;
                }
                acc[w] += RosenbrockVec::term(w0, w1);
            }
        }
        fl = __reduce_or_sync(0xffffffffu, fl);
        changed = (fl & 1u) != 0;
        same_ref = (MODE == 1) ? ((fl & 2u) == 0) : false;
        return warp_team_sum<VW>(acc);
    }
    // bit 0: any(x != x + alpha*dir)   bit 1: any(dir != 0)
    DZO_DEVINL unsigned point_flags(const double2 (&Dir)[VW], double alpha) const {
        unsigned fl = 0;
#pragma unroll
        for (int w = 0; w < VW; ++w)
            if (has[w]) {
                const double2 xx = X[w], dd = Dir[w];
                if ((xx.x != xx.x + alpha * dd.x) | (xx.y != xx.y + alpha * dd.y)) fl |= 1u;
                if ((!(dd.x == 0.0)) | (!(dd.y == 0.0))) fl |= 2u;
            }
        return __reduce_or_sync(0xffffffffu, fl);
    }
    // quadratic_line_search [GLUE: legacy/DZOptimization.jl:49-172 with first step t1, then :191-216]
    DZO_DEVINL void line_search(const double2 (&Dir)[VW], double f0, double t1, double sign, double& t_best, double& f_best,
                                long long& evals) const {
        double x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0;
        bool changed, same;
        do {
            if (!isfinite(f0)) break;                                         // :64-66
            if (!isfinite(t1) || t1 == 0.0) break;                            // [GLUE]
            double step = t1;
            unsigned fl = point_flags(Dir, sign * step);
            if (!(fl & 2u)) break;                                            // :71-85 step_is_zero
            int cap = DZO_LINESEARCH_CAP;
            bool capped = false;
            while (!(fl & 1u)) {                                              // :91-101
                step += step;
                fl = point_flags(Dir, sign * step);
                if (--cap == 0) { capped = true; break; }
            }
            if (capped) break;
            double fa = probe<0>(Dir, sign * step, 0.0, changed, same);       // :126
            ++evals;
            if (fa <= f0) {                                                   // :130
                cap = DZO_LINESEARCH_CAP;
                for (;;) {                                                    // :143-156 (max_increases = 0)
                    const double ds = step + step;
                    const double fb = probe<1>(Dir, sign * ds, sign * step, changed, same);
                    ++evals;
                    --cap;
                    if (!isfinite(fb) || fb > fa || same || cap == 0) {
                        x1 = step; f1 = fa; x2 = ds; f2 = fb;
                        break;
                    }
                    step = ds;
                    fa = fb;
                }
            } else {                                                          // :157-171
                cap = DZO_LINESEARCH_CAP;
                for (;;) {
                    const double hs = 0.5 * step;
                    const double fb = probe<0>(Dir, sign * hs, 0.0, changed, same);
                    ++evals;
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
        double xb = 0.0, fb = f0;                                             // :196-202
        if (f1 < fb) { xb = x1; fb = f1; }
        if (f2 < fb) { xb = x2; fb = f2; }
        const double delta_1 = f0 - f1;                                       // :203-205
        const double delta_2 = f2 - f1;
        const double sum_deltas = delta_1 + delta_2;
        if (delta_1 >= 0.0 && delta_2 >= 0.0 && sum_deltas > 0.0) {           // :206-214
            const double twice_delta_1 = delta_1 + delta_1;
            const double delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas);
            const double xq = delta_ratio * x1;
            const double fq = probe<0>(Dir, sign * xq, 0.0, changed, same);
            ++evals;
            if (fq < fb) { xb = xq; fb = fq; }
        }
        t_best = xb;
        f_best = fb;
    }
};

// step! :891-950 and :873-874 for `batch` problems, one warp each.  Same control block, same vectors and the same
// results as cluster_bfgs_search_kernel.
template <int VW>
static __global__ void __launch_bounds__(32 * kWarpSearchWarps) warp_bfgs_search_kernel(LargeVecs a_, long long batch) {
    const long long q = (long long)blockIdx.x * kWarpSearchWarps + (threadIdx.x >> 5);
    if (q >= batch) return;
    const LargeVecs a = for_problem(a_, q);
    const int lane = threadIdx.x & 31;
    const long long m = a.n >> 1;
    const LargeCtrl sc = *a.ctrl;                  // every lane holds the control block; lane 0 rewrites it at the end
    __syncwarp();
    if (sc.term) {                                                            // :893
        if (lane == 0) a.ctrl->kind = DZO_STEP_NULL;
        return;
    }
    WarpTeam<VW> T;
    T.lane = lane;
#pragma unroll
    for (int w = 0; w < VW; ++w) {
        const long long k = 32 * w + lane;
        T.has[w] = k < m;
        const double2 z = make_double2(0.0, 0.0);
        T.X[w] = T.has[w] ? reinterpret_cast<const double2*>(a.x)[k] : z;
        T.G[w] = T.has[w] ? reinterpret_cast<const double2*>(a.g)[k] : z;
        T.D[w] = T.has[w] ? reinterpret_cast<const double2*>(a.d)[k] : z;
    }
    const double f0 = sc.f;
    const double step_length = sc.L;                                          // :918
    long long evals = 0;
    double grad_norm, bfgs_norm;                                              // :921, :928
    {
        double ag[VW], ad[VW];
#pragma unroll
        for (int w = 0; w < VW; ++w) {
            ag[w] = 0.0; ad[w] = 0.0;
            if (T.has[w]) {
                ag[w] += T.G[w].x * T.G[w].x; ag[w] += T.G[w].y * T.G[w].y;
                ad[w] += T.D[w].x * T.D[w].x; ad[w] += T.D[w].y * T.D[w].y;
            }
        }
        grad_norm = sqrt(warp_team_sum<VW>(ag));
        bfgs_norm = sqrt(warp_team_sum<VW>(ad));
    }
    double grad_step_length, grad_obj, bfgs_step_length, bfgs_obj;
    T.line_search(T.G, f0, step_length / grad_norm, -1.0, grad_step_length, grad_obj, evals);   // :922-925
    T.line_search(T.D, f0, step_length / bfgs_norm, -1.0, bfgs_step_length, bfgs_obj, evals);   // :929-932

    int kind;
    double alpha, fnew, Lnew;
    if (bfgs_obj < f0 && !(bfgs_obj > grad_obj)) {                            // :934
        kind = DZO_STEP_BFGS; alpha = -bfgs_step_length; fnew = bfgs_obj; Lnew = bfgs_step_length * bfgs_norm;
    } else if (grad_obj < f0) {                                               // :962
        kind = DZO_STEP_GRADIENT_DESCENT; alpha = -grad_step_length; fnew = grad_obj; Lnew = grad_step_length * grad_norm;
    } else {
        if (lane == 0) {                                                      // :989
            a.ctrl->term = 1;
            a.ctrl->kind = DZO_STEP_NULL;
            a.ctrl->evals = sc.evals + evals;
            a.ctrl->kind_log[sc.calls & 63] = DZO_STEP_NULL;
            a.ctrl->calls = sc.calls + 1;
        }
        return;
    }
    double acc[VW];                                                           // :943-950 / :971-978, overlap :873
#pragma unroll
    for (int w = 0; w < VW; ++w) {
        acc[w] = 0.0;
        if (T.has[w]) {
            const long long k = 32 * w + lane;
            const double2 xx = T.X[w], gg = T.G[w];
            const double2 dd = (kind == DZO_STEP_BFGS) ? T.D[w] : T.G[w];
            double2 xn, dxv, dgv;
            xn.x = xx.x + alpha * dd.x;
            xn.y = xx.y + alpha * dd.y;
            const double2 gn = RosenbrockVec::grad(xn.x, xn.y);
            dxv.x = (-xx.x) + xn.x; dxv.y = (-xx.y) + xn.y;
            dgv.x = (-gg.x) + gn.x; dgv.y = (-gg.y) + gn.y;
            reinterpret_cast<double2*>(a.x)[k] = xn;
            reinterpret_cast<double2*>(a.g)[k] = gn;
            reinterpret_cast<double2*>(a.dx)[k] = dxv;
            reinterpret_cast<double2*>(a.dg)[k] = dgv;
            if (kind == DZO_STEP_BFGS) {
                acc[w] += dd.x * dgv.x;
                acc[w] += dd.y * dgv.y;
            } else {
                reinterpret_cast<double2*>(a.d)[k] = gn;                      // :984-986
            }
        }
    }
    double overlap = 0.0;
    if (kind == DZO_STEP_BFGS) {
        overlap = warp_team_sum<VW>(acc);
        const double inv_overlap = 1.0 / overlap;                             // :874
#pragma unroll
        for (int w = 0; w < VW; ++w)
            if (T.has[w])
                reinterpret_cast<double2*>(a.sd)[32 * w + lane] = make_double2(T.D[w].x * inv_overlap, T.D[w].y * inv_overlap);
    }
    if (lane == 0) {
        LargeCtrl c = sc;
        c.f = fnew; c.L = Lnew; c.type = kind; c.iter = sc.iter + 1;
        c.kind = kind; c.step_length = alpha; c.overlap = overlap; c.delta_norm = 0.0;
        c.evals = sc.evals + evals;
        c.kind_log[sc.calls & 63] = (unsigned char)kind;
        c.calls = sc.calls + 1;
        *a.ctrl = c;
    }
}

// :876  delta = alpha*overlap + dg . t, one warp per problem
template <int VW>
static __global__ void __launch_bounds__(32 * kWarpSearchWarps) warp_delta_kernel(LargeVecs a_, long long batch) {
    const long long q = (long long)blockIdx.x * kWarpSearchWarps + (threadIdx.x >> 5);
    if (q >= batch) return;
    const LargeVecs a = for_problem(a_, q);
    if (a.ctrl->kind != DZO_STEP_BFGS) return;
    const int lane = threadIdx.x & 31;
    const long long m = a.n >> 1;
    double acc[VW];
#pragma unroll
    for (int w = 0; w < VW; ++w) {
        acc[w] = 0.0;
        const long long k = 32 * w + lane;
        if (k < m) {
            const double2 dg = reinterpret_cast<const double2*>(a.dg)[k];
            const double2 t = reinterpret_cast<const double2*>(a.t)[k];
            acc[w] += dg.x * t.x;
            acc[w] += dg.y * t.y;
        }
    }
    const double dot = warp_team_sum<VW>(acc);
    if (lane == 0) a.ctrl->delta_norm = a.ctrl->step_length * a.ctrl->overlap + dot;
}

}  // namespace dzo
