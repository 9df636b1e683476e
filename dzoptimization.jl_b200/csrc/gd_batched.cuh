// gd_batched.cuh -- batched small-n GradientDescentOptimizer (n <= 32): ONE THREAD PER PROBLEM.
//
// A gradient-descent step! (legacy/DZOptimization.jl:393-449) touches only O(n) state, so a whole problem --
// bracketing line search (:49-172, :191-216 with QuadraticLineSearch(max_increases)), point / gradient
// bookkeeping, norms -- lives in one thread's registers / local memory and every sum is the strictly
// sequential sum of the reference (legacy/Kernels.jl:12-20, :49-55): results are bitwise those of `batch`
// separate CPU optimizers (oracle DZO_ORDER_SEQUENTIAL).  Objectives: extended Rosenbrock and Riesz energy
// (legacy/ExampleFunctions.jl:10-24, :30-83) with the optional sphere constraint.
#pragma once
#include "common.cuh"

namespace dzo {

struct GdBatchedArgs {
    double *x, *dx, *g, *dg, *d;       // n x batch
    double *f, *df, *L;                // batch
    long long* iter;
    unsigned char* term;
    int n, dim, objective, sphere, max_increases, ksteps;
    long long batch;
    double initial_step_length;
    int mode;                          // 0 = k step! calls, 1 = constructor
};

constexpr int kGdSmallMax = DZO_SMALL_N_MAX;

struct SmallProblem {
    int n, dim, objective, sphere;

    // constraint_function!(x): NONE -> true; SPHERE -> normalise every column [GLUE SURVEY 8.0]
    DZO_DEVINL void constrain(double* x) const {
        if (!sphere) return;
        const int np = n / dim;
        for (int j = 0; j < np; ++j) {
            double s = 0.0;
            for (int k = 0; k < dim; ++k) s += x[k + j * dim] * x[k + j * dim];
            const double inv = 1.0 / sqrt(s);
            for (int k = 0; k < dim; ++k) x[k + j * dim] *= inv;
        }
    }
    DZO_DEVINL double f(const double* x) const {
        double result = 0.0;
        if (objective == DZO_OBJ_ROSENBROCK) {                     // legacy/ExampleFunctions.jl:10-15 per pair
            for (int k = 0; k < n / 2; ++k) {
                const double a = x[2 * k], b = x[2 * k + 1];
                const double t1 = 1 - a;
                const double t2 = b - a * a;
                result += t1 * t1 + 100 * (t2 * t2);
            }
            return result;
        }
        const int np = n / dim;                                    // riesz_energy :30-45
        for (int j = 1; j < np; ++j)
            for (int i = 0; i < j; ++i) {
                double dist_sq = 0.0;
                for (int k = 0; k < dim; ++k) {
                    const double dist = x[k + i * dim] - x[k + j * dim];
                    dist_sq += dist * dist;
                }
                result += 1.0 / sqrt(dist_sq);
            }
        return result;
    }
    DZO_DEVINL void grad(double* g, const double* x) const {
        if (objective == DZO_OBJ_ROSENBROCK) {                     // :17-24
            for (int k = 0; k < n / 2; ++k) {
                const double a = x[2 * k], b = x[2 * k + 1];
                const double t1 = 1 - a;
                const double t2 = b - a * a;
                g[2 * k] = -2 * t1 - 400 * a * t2;
                g[2 * k + 1] = 200 * t2;
            }
            return;
        }
        const int np = n / dim;                                    // riesz_gradient! :47-83 (+ :361-374)
        for (int j = 0; j < np; ++j) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            for (int i = 0; i < np; ++i) {
                if (i == j) continue;
                double dist_sq = 0.0;
                for (int k = 0; k < dim; ++k) {
                    const double dist = x[k + i * dim] - x[k + j * dim];
                    dist_sq += dist * dist;
                }
                const double inv_dist = 1.0 / sqrt(dist_sq);
                const double inv_dist_cubed = inv_dist / dist_sq;
                for (int k = 0; k < dim; ++k) {
                    const double dist = x[k + i * dim] - x[k + j * dim];
                    acc[k] += dist * inv_dist_cubed;
                }
            }
            for (int k = 0; k < dim; ++k) g[k + j * dim] = acc[k];
            if (sphere) {
                double overlap = 0.0;
                for (int k = 0; k < dim; ++k) overlap += x[k + j * dim] * g[k + j * dim];
                for (int k = 0; k < dim; ++k) g[k + j * dim] -= overlap * x[k + j * dim];
            }
        }
    }
    DZO_DEVINL double norm2(const double* v) const {               // Kernels.norm2 :49-55
        double r = 0.0;
        for (int i = 0; i < n; ++i) r += v[i] * v[i];
        return r;
    }
};

// lse(t) :25-46: w = t*d + x; constraint!(w); f(w).  Returns "point changed" through `changed`.
DZO_DEVINL double small_probe(const SmallProblem& P, const double* x, const double* d, double t, double* w, bool& changed) {
    changed = false;   // t carries the sign: +t for LineSearchEvaluator (:33), -t for the BFGS functor (:945,:973)
    for (int i = 0; i < P.n; ++i) {
        const double nw = x[i] + t * d[i];
        changed |= (x[i] != nw);
        w[i] = nw;
    }
    P.constrain(w);
    return P.f(w);
}
DZO_DEVINL bool small_equal(const double* a, const double* b, int n) {
    bool eq = true;
    for (int i = 0; i < n; ++i) eq &= (a[i] == b[i]);
    return eq;
}

// QuadraticLineSearch(max_increases)(lse, f0, _)  :191-216 over find_three_point_bracket :49-172 (first step 1)
DZO_DEVINL void small_line_search(const SmallProblem& P, const double* x, const double* d, double f0, int max_increases,
                                  double* w, double* ref, double& t_best, double& f_best, double t1 = 1.0,
                                  double sign = 1.0) {
    const int n = P.n;
    double x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0;
    do {
        if (!isfinite(f0)) break;                                  // :64-66
        if (!isfinite(t1) || t1 == 0.0) break;                     // [GLUE]
        bool zero = true;
        for (int i = 0; i < n; ++i) zero &= (d[i] == 0.0);
        if (zero) break;                                           // :71-85
        double step = t1;
        bool changed;
        bool small = false;
        int cap = DZO_LINESEARCH_CAP;
        double fa = 0.0;
        for (;;) {                                                 // :73-101
            changed = false;
            for (int i = 0; i < n; ++i) {
                const double nw = x[i] + (sign * step) * d[i];
                changed |= (x[i] != nw);
                w[i] = nw;
            }
            if (changed) break;
            step += step;
            small = true;
            if (--cap == 0) break;
        }
        if (!changed) break;
        P.constrain(w);                                            // :104
        if (small && small_equal(x, w, n)) break;                  // :107-123
        fa = P.f(w);                                               // :126
        if (fa <= f0) {                                            // :130
            for (int i = 0; i < n; ++i) ref[i] = w[i];             // :136
            int num_increases = 0;
            cap = DZO_LINESEARCH_CAP;
            for (;;) {                                             // :143-156
                const double ds = step + step;
                num_increases += 1;
                bool ch;
                const double fb = small_probe(P, x, d, sign * ds, w, ch);
                --cap;
                if (((max_increases > 0) && (num_increases >= max_increases)) || !isfinite(fb) || fb > fa ||
                    small_equal(w, ref, n) || cap == 0) {
                    x1 = step; f1 = fa; x2 = ds; f2 = fb;
                    break;
                }
                step = ds;
                fa = fb;
                for (int i = 0; i < n; ++i) ref[i] = w[i];
            }
        } else {                                                   // :157-171
            cap = DZO_LINESEARCH_CAP;
            for (;;) {
                const double hs = 0.5 * step;
                bool ch;
                const double fb = small_probe(P, x, d, sign * hs, w, ch);
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
    double xb = 0.0, fb = f0;                                      // :196-202
    if (f1 < fb) { xb = x1; fb = f1; }
    if (f2 < fb) { xb = x2; fb = f2; }
    const double delta_1 = f0 - f1;                                // :203-205
    const double delta_2 = f2 - f1;
    const double sum_deltas = delta_1 + delta_2;
    if (delta_1 >= 0.0 && delta_2 >= 0.0 && sum_deltas > 0.0) {    // :206-214
        const double twice_delta_1 = delta_1 + delta_1;
        const double delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas);
        const double xq = delta_ratio * x1;
        bool ch;
        const double fq = small_probe(P, x, d, sign * xq, w, ch);
        if (fq < fb) { xb = xq; fb = fq; }
    }
    t_best = xb;
    f_best = fb;
}

static __global__ void __launch_bounds__(128) gd_batched_kernel(GdBatchedArgs a) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.batch) return;
    SmallProblem P{a.n, a.dim > 0 ? a.dim : 1, a.objective, a.sphere};
    const int n = a.n;
    double x[kGdSmallMax], g[kGdSmallMax], d[kGdSmallMax], w[kGdSmallMax], ref[kGdSmallMax];
    for (int i = 0; i < n; ++i) x[i] = a.x[p * n + i];
    if (a.mode == 1) {
        // GradientDescentOptimizer(...)  :330-374
        P.constrain(x);                                            // :340
        const double f0 = P.f(x);                                  // :343
        P.grad(g, x);                                              // :347-348
        const double inv_gradient_norm = 1.0 / sqrt(P.norm2(g));   // :352
        const bool ok = isfinite(inv_gradient_norm);
        const double alpha = -a.initial_step_length * inv_gradient_norm;
        for (int i = 0; i < n; ++i) {
            a.x[p * n + i] = x[i];
            a.g[p * n + i] = g[i];
            a.dx[p * n + i] = 0.0;
            a.dg[p * n + i] = 0.0;
            a.d[p * n + i] = ok ? g[i] * alpha : 0.0;              // :354-357
        }
        a.f[p] = f0; a.df[p] = 0.0; a.L[p] = 0.0; a.iter[p] = 0;
        a.term[p] = (!isfinite(f0)) || !ok;                        // :364-366
        return;
    }
    if (a.term[p]) return;                                         // :402
    for (int i = 0; i < n; ++i) { g[i] = a.g[p * n + i]; d[i] = a.d[p * n + i]; }
    double f0 = a.f[p], df = a.df[p], L = a.L[p];
    long long iter = a.iter[p];
    bool term = false, moved = false;
    double dxv[kGdSmallMax], dgv[kGdSmallMax];
    for (int s = 0; s < a.ksteps && !term; ++s) {
        double step_size, objective_value;
        small_line_search(P, x, d, f0, a.max_increases, w, ref, step_size, objective_value);   // :405-407
        if (step_size == 0.0 || !(objective_value < f0)) { term = true; break; }              // :410-414
        iter += 1;                                                 // :415
        moved = true;
        for (int i = 0; i < n; ++i) { dxv[i] = x[i]; x[i] += step_size * d[i]; }               // :418-419
        P.constrain(x);                                            // :420
        for (int i = 0; i < n; ++i) dxv[i] = x[i] - dxv[i];        // :423
        L = sqrt(P.norm2(dxv));                                    // :424-425
        df = objective_value - f0;                                 // :428-429
        f0 = objective_value;                                      // :430
        for (int i = 0; i < n; ++i) dgv[i] = g[i];                 // :433
        P.grad(g, x);                                              // :434
        for (int i = 0; i < n; ++i) dgv[i] = g[i] - dgv[i];        // :435
        const double inv_gradient_norm = 1.0 / sqrt(P.norm2(g));   // :438
        if (!isfinite(inv_gradient_norm)) { term = true; break; }  // :439-442
        const double alpha = -L * inv_gradient_norm;
        for (int i = 0; i < n; ++i) d[i] = alpha * g[i];           // :445-446
    }
    if (moved) {
        for (int i = 0; i < n; ++i) {
            a.x[p * n + i] = x[i]; a.g[p * n + i] = g[i]; a.d[p * n + i] = d[i];
            a.dx[p * n + i] = dxv[i]; a.dg[p * n + i] = dgv[i];
        }
        a.f[p] = f0; a.df[p] = df; a.L[p] = L; a.iter[p] = iter;
    }
    if (term) a.term[p] = 1;
}

// ----------------------------------------------------------------------------- generic batched BFGS, one thread per problem
// step!(::BFGSOptimizer) :891-994 for objectives the warp-resident kernels do not cover (Riesz energy with or
// without the sphere constraint, n <= 32).  A literal transcription, bitwise the sequential reference order.  The
// inverse Hessians are stored BATCH-INNERMOST (element (i, j) of problem p at H[(i + j*n) * batch + p]): the threads of a
// warp walk their matrices in lockstep, so every access is 32 consecutive doubles (round 2: with the problem-major layout
// every thread streamed its own 4.6 KB matrix; 200 k problems of 8 points x 3: 9.05 -> 6.58 ms per step! call).  The
// rank-2 update and d = H*g share one sweep (each d[i] still accumulates its columns in ascending order).  What is left
// is the thread-per-problem line search itself (divergent probe counts, vectors in local memory): still "correct, not
// tuned" next to the warp-resident Rosenbrock kernel.
struct BfgsGenericArgs {
    double *x, *g, *d, *dx, *dg, *H, *f, *L;
    long long* iter;
    int* type;
    unsigned char* term;
    int n, dim, objective, sphere, ksteps;
    long long batch;
    double initial_step_length;
    int mode;   // 0 = steps, 1 = constructor, 2 = resume (recompute f, g, d = H*g)
    long long hs;   // stride between consecutive elements of one problem's matrix: batch (batch-innermost) or 1 (problem-major)
};

static __global__ void __launch_bounds__(64) bfgs_generic_kernel(BfgsGenericArgs a) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.batch) return;
    SmallProblem P{a.n, a.dim > 0 ? a.dim : 1, a.objective, a.sphere};
    const int n = a.n;
    const long long hs = a.hs;
    double* H = a.H + ((hs == 1) ? p * n * n : p);          // element (i, j) at H[(i + j * n) * hs]
    double x[kGdSmallMax], g[kGdSmallMax], d[kGdSmallMax], w[kGdSmallMax], ref[kGdSmallMax];
    for (int i = 0; i < n; ++i) x[i] = a.x[p * n + i];
    if (a.mode == 1 || a.mode == 2) {
        P.constrain(x);                                            // :770 / :826
        const double f0 = P.f(x);                                  // :772 / :828
        P.grad(g, x);                                              // :775 / :830
        for (int i = 0; i < n; ++i) {
            a.x[p * n + i] = x[i];
            a.g[p * n + i] = g[i];
        }
        if (a.mode == 1) {
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) H[(i + j * n) * hs] = (i == j) ? 1.0 : 0.0;    // :781-783
            for (int i = 0; i < n; ++i) { a.d[p * n + i] = g[i]; a.dx[p * n + i] = 0.0; a.dg[p * n + i] = 0.0; }   // :784
            a.L[p] = a.initial_step_length; a.iter[p] = 0; a.type[p] = DZO_STEP_NULL;
        } else {
            for (int i = 0; i < n; ++i) {                          // :833-836 d = H*g
                double acc = 0.0;
                for (int j = 0; j < n; ++j) acc += H[(i + j * n) * hs] * g[j];
                a.d[p * n + i] = acc;
            }
        }
        a.f[p] = f0;
        a.term[p] = 0;
        return;
    }
    if (a.term[p]) return;                                         // :893
    for (int i = 0; i < n; ++i) { g[i] = a.g[p * n + i]; d[i] = a.d[p * n + i]; }
    double f0 = a.f[p], L = a.L[p];
    long long iter = a.iter[p];
    int type = DZO_STEP_NULL;
    bool term = false, moved = false;
    double dxv[kGdSmallMax], dgv[kGdSmallMax], tv[kGdSmallMax];
    for (int s = 0; s < a.ksteps && !term; ++s) {
        const double grad_norm = sqrt(P.norm2(g));                 // :921
        double tg, fg, tb, fb;
        small_line_search(P, x, g, f0, 0, w, ref, tg, fg, L / grad_norm, -1.0);    // :922-925
        const double bfgs_norm = sqrt(P.norm2(d));                 // :928
        small_line_search(P, x, d, f0, 0, w, ref, tb, fb, L / bfgs_norm, -1.0);    // :929-932
        int kind;
        double alpha;
        if (fb < f0 && !(fb > fg)) { kind = DZO_STEP_BFGS; alpha = -tb; f0 = fb; L = tb * bfgs_norm; }          // :934-938
        else if (fg < f0) { kind = DZO_STEP_GRADIENT_DESCENT; alpha = -tg; f0 = fg; L = tg * grad_norm; }       // :962-966
        else { term = true; break; }                               // :989
        type = kind; iter += 1; moved = true;
        const double* dir = (kind == DZO_STEP_BFGS) ? d : g;
        for (int i = 0; i < n; ++i) { dxv[i] = -x[i]; dgv[i] = -g[i]; }            // :943-944
        for (int i = 0; i < n; ++i) x[i] += alpha * dir[i];        // :945 (dir may alias g: g is rewritten below)
        P.constrain(x);                                            // :946
        P.grad(g, x);                                              // :948
        for (int i = 0; i < n; ++i) { dxv[i] += x[i]; dgv[i] += g[i]; }            // :949-950
        if (kind == DZO_STEP_BFGS) {
            double overlap = 0.0;                                  // :873
            for (int i = 0; i < n; ++i) overlap += d[i] * dgv[i];
            const double inv_overlap = 1.0 / overlap;
            for (int i = 0; i < n; ++i) d[i] *= inv_overlap;       // :874
            for (int i = 0; i < n; ++i) tv[i] = 0.0;               // :875, column by column: tv[i] still sums j ascending
            for (int j = 0; j < n; ++j) {
                const double vj = dgv[j];
                for (int i = 0; i < n; ++i) tv[i] += H[(i + j * n) * hs] * vj;
            }
            double dot = 0.0;
            for (int i = 0; i < n; ++i) dot += dgv[i] * tv[i];
            const double delta_norm = alpha * overlap + dot;       // :876
            for (int i = 0; i < n; ++i) w[i] = 0.0;
            for (int j = 0; j < n; ++j) {                          // :878-886 fused with :958-960 (d = H*g, into w)
                const double sj = d[j], tj = tv[j], gj = g[j];
                for (int i = 0; i < n; ++i) {
                    const double h = H[(i + j * n) * hs] + (delta_norm * (d[i] * sj) - (tv[i] * sj + d[i] * tj));
                    H[(i + j * n) * hs] = h;
                    w[i] += h * gj;
                }
            }
            for (int i = 0; i < n; ++i) d[i] = w[i];
        } else {
            for (int j = 0; j < n; ++j)
                for (int i = 0; i < n; ++i) H[(i + j * n) * hs] = (i == j) ? 1.0 : 0.0;    // :981
            for (int i = 0; i < n; ++i) d[i] = g[i];               // :984-986
        }
    }
    if (moved) {
        for (int i = 0; i < n; ++i) {
            a.x[p * n + i] = x[i]; a.g[p * n + i] = g[i]; a.d[p * n + i] = d[i];
            a.dx[p * n + i] = dxv[i]; a.dg[p * n + i] = dgv[i];
        }
        a.f[p] = f0; a.L[p] = L; a.iter[p] = iter; a.type[p] = type;
    }
    if (term) a.term[p] = 1;
}

}  // namespace dzo
