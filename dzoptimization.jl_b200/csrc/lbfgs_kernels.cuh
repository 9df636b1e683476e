// lbfgs_kernels.cuh -- the LIVE package's LBFGSOptimizer (src/DZOptimization.jl:321-509) and
// take_backtracking_step! (:107-154) as ONE cluster kernel per step!.
//
// Everything in an L-BFGS step is O(n*m) vector work separated by dot products, so the whole step --
// two-loop recursion (2m+1 reductions), backtracking line search (one fused "move + compare + objective"
// pass and one reduction per trial), gradient, history push, rho -- runs in a single launch of an
// 8-CTA x 512-thread cluster whose 4096 threads ARE the virtual threads of the canonical tree
// (cluster_search.cuh): thread v owns the element pairs v, v+4096, ... of every vector, so between
// reductions each thread only re-reads what it wrote itself and no other synchronisation is needed.
#pragma once
#include "cluster_search.cuh"

namespace dzo {

struct LbfgsCtrl {
    double f, df;                            // current / delta objective value   :332-333
    long long iter;                          // iteration_count                   :328
    int stuck;                               // is_stuck                          :327
    int count;                               // valid history entries
    int head;                                // physical slot of the newest entry (logical index 0)
    int pad;
    double rho[DZO_LBFGS_MAX_HISTORY];       // rho_history, PHYSICAL slots       :342
    long long evals;
    double yy;                               // grid-wide kernel only: dot(y, y) of the newest history entry (:443), computed
                                             // in the accept pass that wrote y -- the same sum the next step! would redo
};

struct LbfgsArgs {
    double *x, *dx, *g, *dg, *d;
    double *S, *Y;                           // m x n, physical slot p at S + p*n
    LbfgsCtrl* ctrl;
    long long n;
    int m, ksteps;
    double initial_step_length;
    int mode;                                // 0 = steps, 1 = constructor
};

DZO_DEVINL bool julia_isequal(double a, double b) {   // isequal: NaN == NaN, -0.0 != 0.0
    if (a == b) return signbit(a) == signbit(b);
    return (a != a) && (b != b);
}

DZO_DEVINL double cluster_dot(cg::cluster_group& cluster, ClusterRed& R, const double* __restrict__ a,
                              const double* __restrict__ b, long long n) {
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    double acc = 0.0;
    for (long long k = v; 2 * k < n; k += DZO_TREE_WIDTH) {
        acc += a[2 * k] * b[2 * k];
        if (2 * k + 1 < n) acc += a[2 * k + 1] * b[2 * k + 1];
    }
    double p[1] = {acc};
    unsigned fl = 0;
    cluster_tree_reduce<1>(cluster, R, p, fl);
    return p[0];
}

// for (own elements) body
#define DZO_OWN_ELEMENTS(e, n, v) for (long long k__ = (v); 2 * k__ < (n); k__ += DZO_TREE_WIDTH) \
                                      for (long long e = 2 * k__; e < 2 * k__ + 2 && e < (n); ++e)

static __global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
    cluster_lbfgs_kernel(LbfgsArgs a) {
    __shared__ ClusterRed R;
    __shared__ LbfgsCtrl sc;
    __shared__ double alpha[DZO_LBFGS_MAX_HISTORY];
    cg::cluster_group cluster = cg::this_cluster();
    const long long n = a.n, m2 = n >> 1;
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    const bool leader = (cluster.block_rank() == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0) { if (a.mode == 0) sc = *a.ctrl; R.parity = 0; }
    __syncthreads();
    cluster.sync();

    if (a.mode == 1) {
        // LBFGSOptimizer(c!, f, g!, x0, L0, m)  :347-427 (x already holds the initial point)
        bool ch, sm_;
        const double f0 = cluster_probe<2>(cluster, R, a.x, a.x, m2, 0.0, 0.0, ch, sm_);        // :417
        for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            reinterpret_cast<double2*>(a.g)[k] = RosenbrockVec::grad(xx.x, xx.y);               // :422
            reinterpret_cast<double2*>(a.dx)[k] = make_double2(0.0, 0.0);                       // :367
            reinterpret_cast<double2*>(a.dg)[k] = make_double2(0.0, 0.0);                       // :372
        }
        const double gnorm = sqrt(cluster_dot(cluster, R, a.g, a.g, n));                        // :376
        const bool stuck = (gnorm == 0.0);                                                      // :377
        const double c = -a.initial_step_length / gnorm;
        DZO_OWN_ELEMENTS(e, n, v) a.d[e] = stuck ? 0.0 : a.g[e] * c;                            // :378-383
        if (leader) {
            LbfgsCtrl t;
            t.f = f0; t.df = 0.0; t.iter = 0; t.stuck = stuck; t.count = 0; t.head = 0; t.pad = 0; t.evals = 1; t.yy = 0.0;
            for (int i = 0; i < DZO_LBFGS_MAX_HISTORY; ++i) t.rho[i] = 0.0;
            *a.ctrl = t;
        }
        return;
    }

    // step! :454-509, k times.  sc is the CTA-local copy of the control block; every CTA updates its copy
    // identically (all values come out of cluster-wide reductions) and the leader publishes it at the end.
    for (int step_i = 0; step_i < a.ksteps; ++step_i) {
        if (sc.stuck) break;                                                                    // :456-458
        const int cnt = sc.count, head = sc.head, m = a.m;
        if (sc.iter > 0) {
            // compute_lbfgs_step_direction!  :430-451 (logical slot i -> physical (head + i) mod m)
            DZO_OWN_ELEMENTS(e, n, v) a.d[e] = a.g[e];
            for (int i = 0; i < cnt; ++i) {
                const int p = (head + i) % m;
                const double al = cluster_dot(cluster, R, a.S + (long long)p * n, a.d, n) / sc.rho[p];   // :439
                if (threadIdx.x == 0) alpha[i] = al;
                const double na = -al;
                const double* y = a.Y + (long long)p * n;
                DZO_OWN_ELEMENTS(e, n, v) a.d[e] += na * y[e];                                  // :440
            }
            if (cnt > 0) {
                const double* y0 = a.Y + (long long)head * n;
                const double c = -sc.rho[head] / cluster_dot(cluster, R, y0, y0, n);            // :443
                DZO_OWN_ELEMENTS(e, n, v) a.d[e] *= c;
            }
            __syncthreads();   // alpha[] visible
            for (int i = cnt - 1; i >= 0; --i) {
                const int p = (head + i) % m;
                const double beta = cluster_dot(cluster, R, a.Y + (long long)p * n, a.d, n) / sc.rho[p];  // :446
                const double na = -(alpha[i] + beta);
                const double* s = a.S + (long long)p * n;
                DZO_OWN_ELEMENTS(e, n, v) a.d[e] += na * s[e];                                  // :447
            }
        }
        // take_backtracking_step!(opt, 1, step_direction)  :107-154
        double step = 1.0, next = 0.0;
        bool accepted = false;
        long long evals = 0;
        for (;;) {
            double acc = 0.0;
            unsigned fl = 0;
            for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
                const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
                const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
                const double w0 = xx.x + step * dd.x, w1 = xx.y + step * dd.y;                 // :124 axpy!
                if (!julia_isequal(w0, xx.x) || !julia_isequal(w1, xx.y)) fl |= 1u;             // :128
                acc += RosenbrockVec::term(w0, w1);                                             // :138
            }
            double pr[1] = {acc};
            cluster_tree_reduce<1>(cluster, R, pr, fl);
            if (!(fl & 1u)) break;                                                              // :128-131 stuck
            ++evals;
            next = pr[0];
            if (next < sc.f) { accepted = true; break; }                                        // :139
            step *= 0.5;                                                                        // :152
        }
        if (!accepted) {
            // the trial point equals the current point: delta_point keeps the COPY of the point (:118)
            DZO_OWN_ELEMENTS(e, n, v) a.dx[e] = a.x[e];
            if (threadIdx.x == 0) { sc.stuck = 1; sc.evals += evals; }
            __syncthreads();
            break;
        }
        // accept: x, delta_point (:145), gradient and delta_gradient (:478-480), history push (:482-505)
        const int slot = (head - 1 + m) % m;
        double* Snew = a.S + (long long)slot * n;
        double* Ynew = a.Y + (long long)slot * n;
        for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
            const double2 go = reinterpret_cast<const double2*>(a.g)[k];
            double2 xn, dxv, dgv;
            xn.x = xx.x + step * dd.x; xn.y = xx.y + step * dd.y;
            dxv.x = 1.0 * xn.x + (-1.0) * xx.x; dxv.y = 1.0 * xn.y + (-1.0) * xx.y;            // axpby!(1, x, -1, dp)
            const double2 gn = RosenbrockVec::grad(xn.x, xn.y);
            dgv.x = 1.0 * gn.x + (-1.0) * go.x; dgv.y = 1.0 * gn.y + (-1.0) * go.y;
            reinterpret_cast<double2*>(a.x)[k] = xn;
            reinterpret_cast<double2*>(a.dx)[k] = dxv;
            reinterpret_cast<double2*>(a.g)[k] = gn;
            reinterpret_cast<double2*>(a.dg)[k] = dgv;
            reinterpret_cast<double2*>(Snew)[k] = dxv;
            reinterpret_cast<double2*>(Ynew)[k] = dgv;
        }
        const double rho_new = cluster_dot(cluster, R, a.dx, a.dg, n);                          // :505
        if (threadIdx.x == 0) {
            sc.df = next - sc.f;                                                                // :142
            sc.f = next;                                                                        // :143
            sc.rho[slot] = rho_new;
            sc.head = slot;
            sc.count = (cnt < m) ? cnt + 1 : m;
            sc.iter += 1;                                                                       // :507
            sc.evals += evals;
        }
        __syncthreads();
    }
    if (leader) *a.ctrl = sc;
}

// ============================================================================= live AdGDOptimizer (:179-312)
struct AdgdCtrl {
    double f, df, cur, prev;
    long long iter;
    int stuck, pad;
};
struct AdgdArgs {
    double *x, *dx, *g, *dg;
    AdgdCtrl* ctrl;
    long long n;
    int ksteps, mode;              // mode 0 = steps, 1 = constructor
    double initial_step_length;
};
DZO_DEVINL double julia_min(double a, double b) {   // Base.min(::Float64, ::Float64)
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}

static __global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
    cluster_adgd_kernel(AdgdArgs a) {
    __shared__ ClusterRed R;
    __shared__ AdgdCtrl sc;
    cg::cluster_group cluster = cg::this_cluster();
    const long long n = a.n, m2 = n >> 1;
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    const bool leader = (cluster.block_rank() == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0) { if (a.mode == 0) sc = *a.ctrl; R.parity = 0; }
    __syncthreads();
    cluster.sync();
    if (a.mode == 1) {                                                                          // :201-271
        bool ch, sm_;
        const double f0 = cluster_probe<2>(cluster, R, a.x, a.x, m2, 0.0, 0.0, ch, sm_);        // :260
        for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            reinterpret_cast<double2*>(a.g)[k] = RosenbrockVec::grad(xx.x, xx.y);               // :265
            reinterpret_cast<double2*>(a.dx)[k] = make_double2(0.0, 0.0);
            reinterpret_cast<double2*>(a.dg)[k] = make_double2(0.0, 0.0);
        }
        const double gnorm = sqrt(cluster_dot(cluster, R, a.g, a.g, n));                        // :233
        if (leader) {
            AdgdCtrl t;
            t.f = f0; t.df = 0.0; t.iter = 0; t.pad = 0;
            t.stuck = (gnorm == 0.0);                                                           // :234
            t.cur = t.prev = t.stuck ? 0.0 : a.initial_step_length / gnorm;                     // :235-236
            *a.ctrl = t;
        }
        return;
    }
    const double inv_sqrt_two = sqrt(0.5);
    for (int step_i = 0; step_i < a.ksteps; ++step_i) {
        if (sc.stuck) break;                                                                    // :276-278
        const double previous = sc.prev, current = sc.cur;
        double next = current;
        if (sc.iter > 0) {                                                                      // :288-297
            const double theta = current / previous;
            next *= sqrt(1.0 + theta);
            const double dgn = sqrt(cluster_dot(cluster, R, a.dg, a.dg, n));
            if (dgn != 0.0) {
                const double inv_L = sqrt(cluster_dot(cluster, R, a.dx, a.dx, n)) / dgn;
                next = julia_min(next, inv_sqrt_two * inv_L);
            }
        }
        // take_backtracking_step!(opt, -next, current_gradient)  :301, :107-154
        double step = -next, nxt = 0.0;
        bool accepted = false;
        for (;;) {
            double acc = 0.0;
            unsigned fl = 0;
            for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
                const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                const double w0 = xx.x + step * gg.x, w1 = xx.y + step * gg.y;
                if (!julia_isequal(w0, xx.x) || !julia_isequal(w1, xx.y)) fl |= 1u;
                acc += RosenbrockVec::term(w0, w1);
            }
            double pr[1] = {acc};
            cluster_tree_reduce<1>(cluster, R, pr, fl);
            if (!(fl & 1u)) break;
            nxt = pr[0];
            if (nxt < sc.f) { accepted = true; break; }
            step *= 0.5;
        }
        if (!accepted) {
            DZO_OWN_ELEMENTS(e, n, v) a.dx[e] = a.x[e];
            if (threadIdx.x == 0) { sc.stuck = 1; sc.prev = current; sc.cur = next; }
            __syncthreads();
            break;
        }
        for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            const double2 go = reinterpret_cast<const double2*>(a.g)[k];
            double2 xn, dxv, dgv;
            xn.x = xx.x + step * go.x; xn.y = xx.y + step * go.y;
            dxv.x = 1.0 * xn.x + (-1.0) * xx.x; dxv.y = 1.0 * xn.y + (-1.0) * xx.y;
            const double2 gn = RosenbrockVec::grad(xn.x, xn.y);
            dgv.x = 1.0 * gn.x + (-1.0) * go.x; dgv.y = 1.0 * gn.y + (-1.0) * go.y;
            reinterpret_cast<double2*>(a.x)[k] = xn;
            reinterpret_cast<double2*>(a.dx)[k] = dxv;
            reinterpret_cast<double2*>(a.g)[k] = gn;
            reinterpret_cast<double2*>(a.dg)[k] = dgv;
        }
        if (threadIdx.x == 0) {
            sc.df = nxt - sc.f; sc.f = nxt; sc.prev = current; sc.cur = next; sc.iter += 1;     // :298-299, :310
        }
        __syncthreads();
    }
    if (leader) *a.ctrl = sc;
}

}  // namespace dzo

// ============================================================================= live LineSearchEvaluator (:12-92)
// (lse::LineSearchEvaluator)(step_size, compute_gradient)  src/DZOptimization.jl:66-92 as one cluster launch:
// trial_point = x + step*dir (:70-71), f_new (:81), improvement_ratio (:85), and with compute_gradient the trial
// gradient (:88) and slope_ratio = dot(trial_gradient, dir) / overlap (:89-90).  out3 = { f_new, improvement_ratio,
// slope_ratio (NaN-free only when compute_gradient) }.
namespace dzo {
struct LseArgs {
    const double *x, *dir;
    double *trial, *trial_g, *out3;
    long long n;
    double f_old, overlap, step;
    int compute_gradient;
};
static __global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
    cluster_lse_kernel(LseArgs a) {
    __shared__ ClusterRed R;
    cg::cluster_group cluster = cg::this_cluster();
    const long long m2 = a.n >> 1;
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    if (threadIdx.x == 0) R.parity = 0;
    __syncthreads();
    cluster.sync();
    double acc = 0.0, ov = 0.0;
    for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
        const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
        const double2 dd = reinterpret_cast<const double2*>(a.dir)[k];
        const double w0 = xx.x + a.step * dd.x, w1 = xx.y + a.step * dd.y;                      // :70-71 copy! + axpy!
        reinterpret_cast<double2*>(a.trial)[k] = make_double2(w0, w1);
        acc += RosenbrockVec::term(w0, w1);                                                     // :81
        if (a.compute_gradient) {
            const double2 tg = RosenbrockVec::grad(w0, w1);                                     // :88
            reinterpret_cast<double2*>(a.trial_g)[k] = tg;
            ov += tg.x * dd.x; ov += tg.y * dd.y;                                               // :89
        }
    }
    double p[2] = {acc, ov};
    unsigned fl = 0;
    cluster_tree_reduce<2>(cluster, R, p, fl);
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
        a.out3[0] = p[0];
        a.out3[1] = (p[0] - a.f_old) / (a.step * a.overlap);                                    // :85
        a.out3[2] = a.compute_gradient ? p[1] / a.overlap : 0.0;                                // :90
    }
}
}  // namespace dzo
