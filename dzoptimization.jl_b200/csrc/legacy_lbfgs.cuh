// legacy_lbfgs.cuh -- the LEGACY LBFGSOptimizer (legacy/DZOptimization.jl:458-695) with the optional
// L2 / uniform-box decorators (:222-296) as ONE cluster kernel per step!.
//
// Same execution model as lbfgs_kernels.cuh: an 8-CTA x 512-thread cluster whose 4096 threads are the
// virtual threads of the canonical tree; thread v owns the element pairs v, v+4096, ... of every vector,
// so between reductions a thread only re-reads what it wrote itself.  A step! is
//   QuadraticLineSearch (:181-216 over find_three_point_bracket :49-172) along next_step_direction
//   [retry along the gradient rescaled to last_step_length, history reset  :589-610]
//   one fused pass: x, delta_point, gradient, delta_gradient, history columns + 4 reductions
//   two-loop correction as written (:656-680), descent check (:683-692)
// Decorators act inside the probe / gradient passes (clamp, lambda*norm2, gradient masking).
#pragma once
#include "lbfgs_kernels.cuh"

namespace dzo {

struct LegacyCtrl {
    double f, df, L;                         // current / delta objective value, last_step_length  :465-466, :474
    long long iter;                          // iteration_count      :477
    int term;                                // has_terminated       :478
    int hist_count;                          // _history_count       :484
    double rho[DZO_LBFGS_MAX_HISTORY];       // _rho   :481 (physical columns)
    double alpha[DZO_LBFGS_MAX_HISTORY];     // _alpha :480
    long long evals;
};

struct LegacyArgs {
    double *x, *dx, *g, *dg, *d;
    double *S, *Y;                           // _delta_point_history / _delta_gradient_history, column c at + c*n
    LegacyCtrl* ctrl;
    long long n;
    int m, ksteps, max_increases, mode;      // mode 0 = steps, 1 = constructor
    int decor;
    int algo;                                // 0 = legacy L-BFGS; 1 = GradientDescentOptimizer (legacy/DZOptimization.jl:393-449):
                                             //     same constructor and line search, no gradient retry, no history
    double initial_step_length, l2, lo, hi;
};

struct LegacyDecor {
    bool l2, box;
    double lam, lo, hi;
    DZO_DEVINL double clamp(double v) const {            // Base.clamp  (UniformBoxConstraint :263-272)
        return box ? ((v > hi) ? hi : ((v < lo) ? lo : v)) : v;
    }
    // gradient! = UniformBoxGradientWrapper(L2GradientWrapper(rosenbrock_gradient!))  :243-251, :281-296
    DZO_DEVINL double2 grad(double x0, double x1) const {
        double2 gg = RosenbrockVec::grad(x0, x1);
        if (l2) {
            const double a = lam + lam;
            gg.x += a * x0;
            gg.y += a * x1;
        }
        if (box) {
            if (((x0 <= lo) && (gg.x >= 0.0)) || ((x0 >= hi) && (gg.x <= 0.0))) gg.x = 0.0;
            if (((x1 <= lo) && (gg.y >= 0.0)) || ((x1 >= hi) && (gg.y <= 0.0))) gg.y = 0.0;
        }
        return gg;
    }
};

// lse(alpha): trial point w = constraint!(x + alpha*dir), objective f(w) (+ lambda*norm2(w)).
// flag bit 0: any(x != x + alpha*dir) BEFORE the constraint (:73-80);  bit 1: any(w != w_ref) (:150, MODE 1);
// bit 2: any(x != w) AFTER the constraint (:119);  MODE 2: evaluate at x itself.
template <int MODE>
DZO_DEVINL double legacy_probe(cg::cluster_group& cluster, ClusterRed& R, const LegacyDecor& D, const double* __restrict__ x,
                               const double* __restrict__ dir, long long m2, double alpha, double alpha_ref, unsigned& flags) {
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    double acc = 0.0, acc2 = 0.0;
    unsigned fl = 0;
    for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
        const double2 xx = reinterpret_cast<const double2*>(x)[k];
        double w0 = xx.x, w1 = xx.y;
        if (MODE != 2) {
            const double2 dd = reinterpret_cast<const double2*>(dir)[k];
            w0 = xx.x + alpha * dd.x;
            w1 = xx.y + alpha * dd.y;
            if ((xx.x != w0) | (xx.y != w1)) fl |= 1u;
            w0 = D.clamp(w0);
            w1 = D.clamp(w1);
            if ((!(xx.x == w0)) | (!(xx.y == w1))) fl |= 4u;
            if (MODE == 1) {
                const double r0 = D.clamp(xx.x + alpha_ref * dd.x);
                const double r1 = D.clamp(xx.y + alpha_ref * dd.y);
                if ((!(w0 == r0)) | (!(w1 == r1))) fl |= 2u;
            }
        }
        acc += RosenbrockVec::term(w0, w1);
        if (D.l2) { acc2 += w0 * w0; acc2 += w1 * w1; }
    }
    double p[2] = {acc, acc2};
    cluster_tree_reduce<2>(cluster, R, p, fl);
    flags = fl;
    return D.l2 ? p[0] + D.lam * p[1] : p[0];                              // :233-234
}

// QuadraticLineSearch(max_increases)(lse, f0, _)  :191-216 with find_three_point_bracket :49-172 (first step 1)
DZO_DEVINL void legacy_line_search(cg::cluster_group& cluster, ClusterRed& R, const LegacyDecor& D, const double* __restrict__ x,
                                   const double* __restrict__ dir, long long n, double f0, int max_increases,
                                   double& t_best, double& f_best, long long& evals) {
    const long long m2 = n >> 1;
    double x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0;
    unsigned pf;
    do {
        if (!isfinite(f0)) break;                                         // :64-66
        double step = 1.0;
        unsigned fl = cluster_point_flags(cluster, R, x, dir, n, step);
        if (!(fl & 2u)) break;                                            // :71-85 step_is_zero
        int cap = DZO_LINESEARCH_CAP;
        bool capped = false, small = false;
        while (!(fl & 1u)) {                                              // :91-101
            step += step;
            small = true;
            fl = cluster_point_flags(cluster, R, x, dir, n, step);
            if (--cap == 0) { capped = true; break; }
        }
        if (capped) break;
        double fa = legacy_probe<0>(cluster, R, D, x, dir, m2, step, 0.0, pf);   // :104, :126
        if (small && !(pf & 4u)) break;                                   // :107-123 (the objective value is unused)
        ++evals;
        if (fa <= f0) {                                                   // :130
            int num_increases = 0;
            cap = DZO_LINESEARCH_CAP;
            for (;;) {                                                    // :143-156
                const double ds = step + step;
                num_increases += 1;
                const double fb = legacy_probe<1>(cluster, R, D, x, dir, m2, ds, step, pf);
                ++evals;
                --cap;
                if (((max_increases > 0) && (num_increases >= max_increases)) || !isfinite(fb) || fb > fa || !(pf & 2u) ||
                    cap == 0) {
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
                const double fb = legacy_probe<0>(cluster, R, D, x, dir, m2, hs, 0.0, pf);
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
        const double fq = legacy_probe<0>(cluster, R, D, x, dir, m2, xq, 0.0, pf);
        ++evals;
        if (fq < fb) { xb = xq; fb = fq; }
    }
    t_best = xb;
    f_best = fb;
}

static __global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
    cluster_legacy_lbfgs_kernel(LegacyArgs a) {
    __shared__ ClusterRed R;
    __shared__ LegacyCtrl sc;
    cg::cluster_group cluster = cg::this_cluster();
    const long long n = a.n, m2 = n >> 1;
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    const bool leader = (cluster.block_rank() == 0 && threadIdx.x == 0);
    LegacyDecor D;
    D.l2 = (a.decor & DZO_DECOR_L2) != 0; D.box = (a.decor & DZO_DECOR_BOX) != 0;
    D.lam = a.l2; D.lo = a.lo; D.hi = a.hi;
    if (threadIdx.x == 0) { if (a.mode == 0) sc = *a.ctrl; R.parity = 0; }
    __syncthreads();
    cluster.sync();

    if (a.mode == 1) {
        // LBFGSOptimizer(c!, f, g!, linesearch, x0, L0, m)  :489-548 (x already holds collect(x0))
        double acc = 0.0;
        for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
            double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            xx.x = D.clamp(xx.x); xx.y = D.clamp(xx.y);                                         // :500
            const double2 gg = D.grad(xx.x, xx.y);                                              // :507-508
            reinterpret_cast<double2*>(a.x)[k] = xx;
            reinterpret_cast<double2*>(a.g)[k] = gg;
            reinterpret_cast<double2*>(a.dx)[k] = make_double2(0.0, 0.0);                       // :501
            reinterpret_cast<double2*>(a.dg)[k] = make_double2(0.0, 0.0);                       // :509
            acc += gg.x * gg.x; acc += gg.y * gg.y;
        }
        unsigned pf;
        const double f0 = legacy_probe<2>(cluster, R, D, a.x, a.x, m2, 0.0, 0.0, pf);           // :503
        double p[1] = {acc};
        unsigned fl = 0;
        cluster_tree_reduce<1>(cluster, R, p, fl);
        const double inv_gradient_norm = 1.0 / sqrt(p[0]);                                      // :512
        const bool ok = isfinite(inv_gradient_norm);
        const double c = -a.initial_step_length * inv_gradient_norm;
        DZO_OWN_ELEMENTS(e, n, v) a.d[e] = ok ? a.g[e] * c : 0.0;                               // :513-517
        if (leader) {
            LegacyCtrl t;
            t.f = f0; t.df = 0.0; t.L = 0.0; t.iter = 0; t.hist_count = 0; t.evals = 1;
            t.term = (!isfinite(f0)) || (!ok);                                                  // :524-526
            for (int i = 0; i < DZO_LBFGS_MAX_HISTORY; ++i) { t.rho[i] = 0.0; t.alpha[i] = 0.0; }
            *a.ctrl = t;
        }
        return;
    }

    // step!  :565-695, k times.  sc is the CTA-local copy of the control block; every CTA updates its copy
    // identically (all values come out of cluster-wide reductions); the leader publishes it at the end.
    const int m = a.m;
    for (int step_i = 0; step_i < a.ksteps; ++step_i) {
        if (sc.term) break;                                                                     // :578
        const double f0 = sc.f;
        long long evals = 0;
        double step_size, objective_value;
        legacy_line_search(cluster, R, D, a.x, a.d, n, f0, a.max_increases, step_size, objective_value, evals);   // :584-586
        bool reset_history = false;
        if (a.algo == 1 && (step_size == 0.0 || !(objective_value < f0))) {                     // GradientDescentOptimizer :410-414
            if (threadIdx.x == 0) { sc.term = 1; sc.evals += evals; }
            __syncthreads();
            break;
        }
        if (step_size == 0.0 || !(objective_value < f0)) {                                      // :589-590
            const double gn2 = cluster_dot(cluster, R, a.g, a.g, n);
            const double c = -sc.L * (1.0 / sqrt(gn2));                                         // :594-595
            DZO_OWN_ELEMENTS(e, n, v) a.d[e] = a.g[e] * c;                                      // :593
            legacy_line_search(cluster, R, D, a.x, a.d, n, f0, a.max_increases, step_size, objective_value, evals);   // :596-598
            if (step_size == 0.0 || !(objective_value < f0)) {                                  // :601-605
                if (threadIdx.x == 0) { sc.term = 1; sc.evals += evals; }
                __syncthreads();
                break;
            }
            reset_history = true;                                                               // :609
        }
        const long long iter = sc.iter + 1;                                                     // :611
        const int cnew = (int)((iter - 1) % m);                                                 // :641
        double* Snew = a.S + (long long)cnew * n;
        double* Ynew = a.Y + (long long)cnew * n;
        double q[4] = {0.0, 0.0, 0.0, 0.0};   // norm2(dx), norm2(g), dot(dx, dg), norm2(dg)
        for (long long k = v; k < m2; k += DZO_TREE_WIDTH) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
            const double2 go = reinterpret_cast<const double2*>(a.g)[k];
            double2 xn, dxv, dgv;
            xn.x = D.clamp(xx.x + step_size * dd.x);                                            // :615-616
            xn.y = D.clamp(xx.y + step_size * dd.y);
            dxv.x = xn.x - xx.x; dxv.y = xn.y - xx.y;                                           // :614, :619 delta!
            const double2 gn = D.grad(xn.x, xn.y);                                              // :630
            dgv.x = gn.x - go.x; dgv.y = gn.y - go.y;                                           // :629, :631
            reinterpret_cast<double2*>(a.x)[k] = xn;
            reinterpret_cast<double2*>(a.dx)[k] = dxv;
            reinterpret_cast<double2*>(a.g)[k] = gn;
            reinterpret_cast<double2*>(a.dg)[k] = dgv;
            if (a.algo == 0) {
                reinterpret_cast<double2*>(Snew)[k] = dxv;                                      // :642-643 (unused once terminated)
                reinterpret_cast<double2*>(Ynew)[k] = dgv;
            }
            q[0] += dxv.x * dxv.x; q[0] += dxv.y * dxv.y;
            q[1] += gn.x * gn.x;   q[1] += gn.y * gn.y;
            q[2] += dxv.x * dgv.x; q[2] += dxv.y * dgv.y;
            q[3] += dgv.x * dgv.x; q[3] += dgv.y * dgv.y;
        }
        unsigned fl = 0;
        cluster_tree_reduce<4>(cluster, R, q, fl);
        const double step_length = sqrt(q[0]);                                                  // :620-621
        const double inv_gradient_norm = 1.0 / sqrt(q[1]);                                      // :634
        if (threadIdx.x == 0) {
            sc.iter = iter;
            sc.L = step_length;
            sc.df = objective_value - f0;                                                       // :624-626
            sc.f = objective_value;
            sc.evals += evals;
            if (reset_history) sc.hist_count = 0;
        }
        if (!isfinite(inv_gradient_norm)) {                                                     // :635-638
            if (threadIdx.x == 0) sc.term = 1;
            __syncthreads();
            break;
        }
        if (a.algo == 1) {                                                                      // GradientDescentOptimizer :445-446
            const double c = -step_length * inv_gradient_norm;
            DZO_OWN_ELEMENTS(e, n, v) a.d[e] = c * a.g[e];
            __syncthreads();
            continue;
        }
        const double delta_overlap = q[2];                                                      // :646
        const double rho_new = 1.0 / delta_overlap;                                             // :647
        const int hist_prev = reset_history ? 0 : sc.hist_count;
        const int hist_count = (hist_prev + 1 < m) ? hist_prev + 1 : m;                         // :650
        const long long hist_end = iter, hist_begin = hist_end - hist_count + 1;                // :651-652
        __syncthreads();                       // everyone has read sc.hist_count / sc.rho before thread 0 rewrites them
        if (threadIdx.x == 0) { sc.rho[cnew] = rho_new; sc.hist_count = hist_count; }
        __syncthreads();
        DZO_OWN_ELEMENTS(e, n, v) a.d[e] = a.g[e];                                              // :656
        for (long long it = hist_end; it >= hist_begin; --it) {                                 // :659-666
            const int c = (int)((it - 1) % m);
            const double alpha = sc.rho[c] * cluster_dot(cluster, R, a.d, a.S + (long long)c * n, n);
            if (threadIdx.x == 0) sc.alpha[c] = alpha;
            const double* y = a.Y + (long long)c * n;
            DZO_OWN_ELEMENTS(e, n, v) a.d[e] += alpha * y[e];
        }
        const double gamma = delta_overlap / q[3];                                              // :669-670
        DZO_OWN_ELEMENTS(e, n, v) a.d[e] *= gamma;
        __syncthreads();                       // sc.alpha[] visible
        for (long long it = hist_begin; it <= hist_end; ++it) {                                 // :673-680
            const int c = (int)((it - 1) % m);
            const double beta = sc.alpha[c] - sc.rho[c] * cluster_dot(cluster, R, a.d, a.Y + (long long)c * n, n);
            const double* s = a.S + (long long)c * n;
            DZO_OWN_ELEMENTS(e, n, v) a.d[e] += beta * s[e];
        }
        DZO_OWN_ELEMENTS(e, n, v) a.d[e] = -a.d[e];                                             // :683-684 negate!
        const double gradient_overlap = cluster_dot(cluster, R, a.d, a.g, n);
        if (!isfinite(gradient_overlap)) {                                                      // :687-688
            if (threadIdx.x == 0) sc.term = 1;
        } else if (gradient_overlap >= 0.0) {                                                   // :689-692
            const double c = -step_length * inv_gradient_norm;
            DZO_OWN_ELEMENTS(e, n, v) a.d[e] = c * a.g[e];
        }
        __syncthreads();
    }
    if (leader) *a.ctrl = sc;
}

}  // namespace dzo
