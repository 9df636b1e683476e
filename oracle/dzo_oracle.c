/* dzo_oracle.c -- CPU restatement of the DZOptimization.jl optimizer step! hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE.  It is the checker for the CUDA library, never the
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  libdzopt_b200.so does not link or call it.
 *
 * PARITY UNPINNED.  The reference (dzhang314/DZOptimization.jl v0.6.0) ships no tests,
 * golden vectors or known-answer files for this path, the path itself is commented-out
 * Julia that calls undefined helpers (legacy/DZOptimization.jl:1, :698, :1144, :1147),
 * and there is no Julia in this environment.  The restatement below follows the
 * reference line by line where the reference has lines, and SURVEY.md section 8.0 where
 * it has gaps (marked [GLUE]).  It is cross-checked against an independent pure-Python
 * restatement (oracle/dzo_oracle_py.py), the run_and_test! invariants
 * (legacy/DZOptimization.jl:998-1049), analytic minima and SciPy; see tests/.
 * Third-party arithmetic the reference reaches through un-pinned dependencies and its
 * substitutes here: LinearAlgebra.mul! (OpenBLAS dgemv; :834,:875,:958) -> row-wise
 * sequential dot; LinearAlgebra.norm (:921,:928) -> sqrt(sequential sum of squares);
 * MultiFloats.rsqrt (legacy/Kernels.jl:141, legacy/ExampleFunctions.jl:41,61,74) ->
 * 1.0/sqrt(x) with two IEEE roundings.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp  (no FMA contraction: the
 * legacy Julia contains no muladd and no @simd reductions on this path).
 *
 * Summation orders
 *   DZO_ORDER_SEQUENTIAL  strict left-to-right, the reference order
 *                         (legacy/Kernels.jl:12-20 dot, :49-55 norm2).
 *   DZO_ORDER_TREE        the canonical tree of the large-n CUDA kernels:
 *       reductions over elements: DZO_TREE_WIDTH (4096) virtual threads; element e is
 *       owned by virtual thread (e div G) mod 4096 (G = 2 for vectors and Rosenbrock
 *       pairs); each virtual thread accumulates its elements in ascending order
 *       starting from +0.0; the 4096 partials are combined by a butterfly over the
 *       index bits in the order 4,3,2,1,0,5,6,7,8,9,10,11 (warp xor-shuffle 16..1, then
 *       warps, then CTAs/replicas, ascending).
 *       GEMV rows: sequential partial per chunk of DZO_GEMV_CHUNK columns, chunk
 *       partials added in ascending chunk order starting from partial 0.
 */
#include "../include/dzopt.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char g_err[256];
const char* dzo_cpu_last_error(void) { return g_err; }
static int fail(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

/* ======================================================================= PCG
 * legacy/PCG.jl:7-22 */
static inline uint64_t pcg_advance(uint64_t s) { /* :7-8 */
    return 0x5851F42D4C957F2DULL * s + 0x14057B7EF767814FULL;
}
static inline uint32_t pcg_extract(uint64_t s) { /* :11-12: rotate RIGHT by the top 5 bits */
    uint32_t v = (uint32_t)(((s >> 18) ^ s) >> 27);
    unsigned r = (unsigned)(s >> 59);
    return (v >> r) | (v << ((32u - r) & 31u));
}
int dzo_cpu_pcg_fill(double* x, int64_t count, uint64_t seed) { /* :15-22 */
    if (!x || count < 0) return fail(DZO_ERR_INVALID_ARGUMENT, "pcg_fill: bad arguments");
    uint64_t state = pcg_advance(0x14057B7EF767814FULL + seed);
    for (int64_t i = 0; i < count; ++i) {
        x[i] = 2.3283064365386962890625E-10 * (double)pcg_extract(state);
        state = pcg_advance(state);
    }
    return DZO_OK;
}

/* ======================================================================= reductions */
static double tree_combine(double* p) {
    static const int bits[12] = {4, 3, 2, 1, 0, 5, 6, 7, 8, 9, 10, 11};
    for (int b = 0; b < 12; ++b) {
        const int m = 1 << bits[b];
        for (int v = 0; v < DZO_TREE_WIDTH; ++v)
            if (!(v & m)) p[v] = p[v] + p[v | m];
    }
    return p[0];
}

/* Kernels.dot  legacy/Kernels.jl:12-20 */
static double dot_(int order, const double* v, const double* w, int64_t n) {
    if (order == DZO_ORDER_SEQUENTIAL) {
        double result = 0.0;
        for (int64_t i = 0; i < n; ++i) result += v[i] * w[i];
        return result;
    }
    double p[DZO_TREE_WIDTH];
    if (order == DZO_ORDER_TREE_BLOCKED) { /* blocks of DZO_TREE_BLOCK elements, block results added in ascending order */
        double total = 0.0;
        for (int64_t b0 = 0; b0 < n || b0 == 0; b0 += DZO_TREE_BLOCK) {
            const int64_t b1 = (b0 + DZO_TREE_BLOCK < n) ? b0 + DZO_TREE_BLOCK : n;
            for (int i = 0; i < DZO_TREE_WIDTH; ++i) p[i] = 0.0;
            for (int64_t e = b0; e < b1; ++e) p[((e - b0) >> 1) & (DZO_TREE_WIDTH - 1)] += v[e] * w[e];
            const double t = tree_combine(p);
            total = (b0 == 0) ? t : total + t;
            if (n == 0) break;
        }
        return total;
    }
    for (int i = 0; i < DZO_TREE_WIDTH; ++i) p[i] = 0.0;
    for (int64_t e = 0; e < n; ++e) p[(e >> 1) & (DZO_TREE_WIDTH - 1)] += v[e] * w[e];
    return tree_combine(p);
}
/* Kernels.norm2  legacy/Kernels.jl:49-55 (abs2(x) = x*x) */
static double norm2_(int order, const double* x, int64_t n) { return dot_(order, x, x, n); }

/* mul!(out, H, v): row i = sum_j H[i,j] v[j], j ascending (substitute for OpenBLAS dgemv,
 * legacy/DZOptimization.jl:875, :958).  rows [r0, r1) only (row-sharded mode), H has
 * leading dimension ld and holds rows r0.. at local row 0. */
static void gemv_rows_(int order, int64_t n, int64_t r0, int64_t r1, const double* H, int64_t ld,
                       const double* v, double* out, int nthreads) {
    const int64_t chunk = (order == DZO_ORDER_SEQUENTIAL) ? (n > 0 ? n : 1) : DZO_GEMV_CHUNK;
    const int64_t rows = r1 - r0;
    const int64_t RB = 512; /* row blocking only changes the traversal, not any sum */
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
    for (int64_t b = 0; b < (rows + RB - 1) / RB; ++b) {
        const int64_t i0 = b * RB, i1 = (i0 + RB < rows) ? i0 + RB : rows;
        double part[512];
        for (int64_t c0 = 0; c0 < n; c0 += chunk) {
            const int64_t c1 = (c0 + chunk < n) ? c0 + chunk : n;
            for (int64_t i = i0; i < i1; ++i) part[i - i0] = 0.0;
            for (int64_t j = c0; j < c1; ++j) {
                const double vj = v[j];
                const double* col = H + j * ld;
                for (int64_t i = i0; i < i1; ++i) part[i - i0] += col[i] * vj;
            }
            if (c0 == 0) for (int64_t i = i0; i < i1; ++i) out[r0 + i] = part[i - i0];
            else         for (int64_t i = i0; i < i1; ++i) out[r0 + i] += part[i - i0];
        }
    }
}

/* ======================================================================= objectives */
typedef struct {
    int objective, constraint, order;
    int64_t n, dim;
    /* decorators (legacy/DZOptimization.jl:222-296); zero = undecorated */
    int decor;
    double l2_lambda, box_lo, box_hi;
} problem_t;

/* rosenbrock_function  legacy/ExampleFunctions.jl:10-15, one pair term */
static inline double rosen_term(double x, double y) {
    const double t1 = 1 - x;
    const double t2 = y - x * x;
    return t1 * t1 + 100 * (t2 * t2);
}

/* riesz_energy  legacy/ExampleFunctions.jl:30-45 */
static double riesz_energy_seq(const double* p, int64_t dim, int64_t np) {
    double result = 0.0;
    for (int64_t j = 1; j < np; ++j)
        for (int64_t i = 0; i < j; ++i) {
            double dist_sq = 0.0;
            for (int64_t k = 0; k < dim; ++k) {
                const double dist = p[k + i * dim] - p[k + j * dim];
                dist_sq += dist * dist;
            }
            result += 1.0 / sqrt(dist_sq); /* rsqrt */
        }
    return result;
}
/* TREE order for the pair sum (DESIGN.md "Riesz tree"): row sum e_j = sum_{i<j} in
 * ascending i, accumulated per segment of DZO_RIESZ_SEG source points (segment partials
 * added in ascending order starting from partial 0); rows then go through the canonical
 * tree with row j owned by virtual thread j mod 4096. */

static double riesz_energy_tree(const double* p, int64_t dim, int64_t np) {
    double part[DZO_TREE_WIDTH];
    for (int i = 0; i < DZO_TREE_WIDTH; ++i) part[i] = 0.0;
    for (int64_t j = 0; j < np; ++j) {
        double ej = 0.0;
        for (int64_t s0 = 0; s0 < j; s0 += DZO_RIESZ_SEG) {
            const int64_t s1 = (s0 + DZO_RIESZ_SEG < j) ? s0 + DZO_RIESZ_SEG : j;
            double seg = 0.0;
            for (int64_t i = s0; i < s1; ++i) {
                double dist_sq = 0.0;
                for (int64_t k = 0; k < dim; ++k) {
                    const double dist = p[k + i * dim] - p[k + j * dim];
                    dist_sq += dist * dist;
                }
                seg += 1.0 / sqrt(dist_sq);
            }
            if (s0 == 0) ej = seg; else ej += seg;
        }
        part[j & (DZO_TREE_WIDTH - 1)] += ej;
    }
    return tree_combine(part);
}

static double base_objective_(const problem_t* P, const double* x) {
    if (P->objective == DZO_OBJ_ROSENBROCK) {
        /* [GLUE] extended Rosenbrock: sum over pairs k ascending */
        const int64_t m = P->n / 2;
        if (P->order == DZO_ORDER_SEQUENTIAL) {
            double result = 0.0;
            for (int64_t k = 0; k < m; ++k) result += rosen_term(x[2 * k], x[2 * k + 1]);
            return result;
        }
        double p[DZO_TREE_WIDTH];
        if (P->order == DZO_ORDER_TREE_BLOCKED) { /* blocks of DZO_TREE_BLOCK / 2 pairs */
            const int64_t pb = DZO_TREE_BLOCK / 2;
            double total = 0.0;
            for (int64_t k0 = 0; k0 < m || k0 == 0; k0 += pb) {
                const int64_t k1 = (k0 + pb < m) ? k0 + pb : m;
                for (int i = 0; i < DZO_TREE_WIDTH; ++i) p[i] = 0.0;
                for (int64_t k = k0; k < k1; ++k) p[(k - k0) & (DZO_TREE_WIDTH - 1)] += rosen_term(x[2 * k], x[2 * k + 1]);
                const double t = tree_combine(p);
                total = (k0 == 0) ? t : total + t;
                if (m == 0) break;
            }
            return total;
        }
        for (int i = 0; i < DZO_TREE_WIDTH; ++i) p[i] = 0.0;
        for (int64_t k = 0; k < m; ++k) p[k & (DZO_TREE_WIDTH - 1)] += rosen_term(x[2 * k], x[2 * k + 1]);
        return tree_combine(p);
    }
    const int64_t np = P->n / P->dim;
    return (P->order == DZO_ORDER_SEQUENTIAL) ? riesz_energy_seq(x, P->dim, np)
                                              : riesz_energy_tree(x, P->dim, np);
}

/* rosenbrock_gradient!  legacy/ExampleFunctions.jl:17-24;  riesz_gradient!  :47-83;
 * with DZO_CONSTRAINT_SPHERE followed by constrain_riesz_gradient_sphere! :361-374. */
static void base_gradient_(const problem_t* P, double* g, const double* x) {
    if (P->objective == DZO_OBJ_ROSENBROCK) {
        for (int64_t k = 0; k < P->n / 2; ++k) {
            const double xx = x[2 * k], y = x[2 * k + 1];
            const double t1 = 1 - xx;
            const double t2 = y - xx * xx;
            g[2 * k] = -2 * t1 - 400 * xx * t2;
            g[2 * k + 1] = 200 * t2;
        }
        return;
    }
    const int64_t dim = P->dim, np = P->n / P->dim;
    /* SEQUENTIAL: one segment covering all sources = the reference loop.  TREE: the
     * sum over sources i is accumulated per segment of DZO_RIESZ_SEG sources, segment
     * partials added in ascending order starting from partial 0. */
    const int64_t seg = (P->order == DZO_ORDER_SEQUENTIAL) ? np : DZO_RIESZ_SEG;
    double acc[16], part[16];
    for (int64_t j = 0; j < np; ++j) {
        for (int64_t k = 0; k < dim; ++k) acc[k] = 0.0;          /* :52-54 */
        for (int64_t s0 = 0; s0 < np; s0 += seg) {
            const int64_t s1 = (s0 + seg < np) ? s0 + seg : np;
            for (int64_t k = 0; k < dim; ++k) part[k] = 0.0;
            for (int64_t i = s0; i < s1; ++i) { /* :55-68 then :69-81: i ascending, skipping j */
                if (i == j) continue;
                double dist_sq = 0.0;
                for (int64_t k = 0; k < dim; ++k) {
                    const double dist = x[k + i * dim] - x[k + j * dim];
                    dist_sq += dist * dist;
                }
                const double inv_dist = 1.0 / sqrt(dist_sq);     /* rsqrt */
                const double inv_dist_cubed = inv_dist / dist_sq;
                for (int64_t k = 0; k < dim; ++k) {
                    const double dist = x[k + i * dim] - x[k + j * dim];
                    part[k] += dist * inv_dist_cubed;
                }
            }
            if (s0 == 0) for (int64_t k = 0; k < dim; ++k) acc[k] = part[k];
            else         for (int64_t k = 0; k < dim; ++k) acc[k] += part[k];
        }
        for (int64_t k = 0; k < dim; ++k) g[k + j * dim] = acc[k];
        if (P->constraint == DZO_CONSTRAINT_SPHERE) { /* :361-374 */
            double overlap = 0.0;
            for (int64_t k = 0; k < dim; ++k) overlap += x[k + j * dim] * g[k + j * dim];
            for (int64_t k = 0; k < dim; ++k) g[k + j * dim] -= overlap * x[k + j * dim];
        }
    }
}

/* constraint_function!(x)::Bool.  NONE: x -> true.  SPHERE [GLUE, SURVEY.md 8.0]:
 * each column p <- p * (1/sqrt(sum p^2)); returns true. */
static int base_constraint_(const problem_t* P, double* x) {
    if (P->constraint == DZO_CONSTRAINT_NONE) return 1;
    const int64_t dim = P->dim, np = P->n / P->dim;
    for (int64_t j = 0; j < np; ++j) {
        double s = 0.0;
        for (int64_t k = 0; k < dim; ++k) s += x[k + j * dim] * x[k + j * dim];
        const double inv = 1.0 / sqrt(s);
        for (int64_t k = 0; k < dim; ++k) x[k + j * dim] *= inv;
    }
    return 1;
}

/* Decorators  legacy/DZOptimization.jl:222-296.  [GLUE] composition order:
 * objective = L2RegularizationWrapper(f, lambda); gradient! = UniformBoxGradientWrapper(
 * L2GradientWrapper(g!, lambda), lo, hi); constraint! = UniformBoxConstraint(lo, hi) after the
 * objective's own constraint. */
static double objective_(const problem_t* P, const double* x) {
    const double f = base_objective_(P, x);
    if (P->decor & DZO_DECOR_L2) return f + P->l2_lambda * norm2_(P->order, x, P->n);  /* :233-234 */
    return f;
}
static void gradient_(const problem_t* P, double* g, const double* x) {
    base_gradient_(P, g, x);
    if (P->decor & DZO_DECOR_L2) {                                   /* :243-251 axpy!(g, 2 lambda, x) */
        const double a = P->l2_lambda + P->l2_lambda;
        for (int64_t i = 0; i < P->n; ++i) g[i] += a * x[i];
    }
    if (P->decor & DZO_DECOR_BOX)                                    /* :281-296 */
        for (int64_t i = 0; i < P->n; ++i)
            if (((x[i] <= P->box_lo) && (g[i] >= 0.0)) || ((x[i] >= P->box_hi) && (g[i] <= 0.0))) g[i] = 0.0;
}
static inline double julia_clamp(double x, double lo, double hi) { /* Base.clamp: ifelse(x > hi, hi, ifelse(x < lo, lo, x)) */
    return (x > hi) ? hi : ((x < lo) ? lo : x);
}
static int constraint_(const problem_t* P, double* x) {
    const int ok = base_constraint_(P, x);
    if (P->decor & DZO_DECOR_BOX)                                    /* :263-272, returns true */
        for (int64_t i = 0; i < P->n; ++i) x[i] = julia_clamp(x[i], P->box_lo, P->box_hi);
    return ok;
}

static int arrays_equal(const double* a, const double* b, int64_t n) { /* Julia == on arrays */
    for (int64_t i = 0; i < n; ++i)
        if (!(a[i] == b[i])) return 0;
    return 1;
}

/* ======================================================================= line search
 * One routine serves both optimizers.  The trial point along the ray is
 *     w[i] = x[i] + alpha * dir[i],   alpha = sign * t,
 * sign = +1: LineSearchEvaluator, legacy/DZOptimization.jl:25-46 (axpy!/5,
 *            legacy/Kernels.jl:127-135: alpha*x[i] + y[i]);
 * sign = -1: the (undefined in the reference) LineSearchFunctor of the BFGS code, whose
 *            sign is fixed by add!(point, -step, direction) at :945,:973 [GLUE].
 * x + (-t)*d and x - t*d are the same IEEE operation, so one formula is exact for both. */
typedef struct {
    const problem_t* P;
    const double* x;   /* initial_point   */
    const double* dir; /* step_direction  */
    double* w;         /* new_point       */
    double* ref;       /* reference_point */
    double sign;
    int64_t evals; /* objective evaluations (statistics only) */
} ray_t;

static int ray_move(ray_t* r, double t) { /* returns point_changed */
    const int64_t n = r->P->n;
    const double alpha = r->sign * t;
    int changed = 0;
    for (int64_t i = 0; i < n; ++i) {
        const double initial = r->x[i];
        const double nw = initial + alpha * r->dir[i];
        changed |= (initial != nw);
        r->w[i] = nw;
    }
    return changed;
}
/* lse(step_size)  :25-46 */
static double ray_eval(ray_t* r, double t) {
    ray_move(r, t);
    const int feasible = constraint_(r->P, r->w);
    if (!feasible) return INFINITY;
    r->evals++;
    return objective_(r->P, r->w);
}

/* find_three_point_bracket  legacy/DZOptimization.jl:49-172, first trial step t1
 * (the reference hard-codes 1 at :77,:89; the BFGS call sites pass L/norm, :922-932). */
static void three_point_bracket(ray_t* r, double f0, double t1, int max_increases,
                                double* x1, double* f1o, double* x2, double* f2o) {
    const int64_t n = r->P->n;
    *x1 = 0.0; *f1o = f0; *x2 = 0.0; *f2o = f0;
    if (!isfinite(f0)) return;                 /* :64-66 */
    if (!isfinite(t1) || t1 == 0.0) return;    /* [GLUE] ||dir|| = 0 or L = 0 => t1 = Inf/NaN/0 */
    int step_is_zero = 1;                      /* :71-85 */
    for (int64_t i = 0; i < n; ++i) step_is_zero &= (r->dir[i] == 0.0);
    if (step_is_zero) return;
    double step_size = t1;
    int point_changed = ray_move(r, step_size); /* :73-80 */
    int step_is_small = 0;
    int cap = DZO_LINESEARCH_CAP;
    while (!point_changed) {                   /* :91-101 */
        step_size += step_size;
        step_is_small = 1;
        point_changed = ray_move(r, step_size);
        if (--cap == 0) return;                /* [GLUE] */
    }
    const int is_feasible = constraint_(r->P, r->w); /* :104 */
    if (step_is_small) {                       /* :107-123 */
        if (!is_feasible) return;
        if (arrays_equal(r->x, r->w, n)) return;
    }
    double f1;
    if (is_feasible) { r->evals++; f1 = objective_(r->P, r->w); } else f1 = INFINITY; /* :126 */
    if (f1 <= f0) {                            /* :130 */
        memcpy(r->ref, r->w, (size_t)n * sizeof(double)); /* :136 */
        int num_increases = 0;
        cap = DZO_LINESEARCH_CAP;
        for (;;) {                             /* :143-156 */
            const double double_step_size = step_size + step_size;
            num_increases += 1;
            const double f2 = ray_eval(r, double_step_size);
            if (((max_increases > 0) && (num_increases >= max_increases)) || (!isfinite(f2)) ||
                (f2 > f1) || arrays_equal(r->w, r->ref, n) || --cap == 0) {
                *x1 = step_size; *f1o = f1; *x2 = double_step_size; *f2o = f2;
                return;
            }
            step_size = double_step_size;
            f1 = f2;
            memcpy(r->ref, r->w, (size_t)n * sizeof(double));
        }
    } else {                                   /* :157-171 */
        cap = DZO_LINESEARCH_CAP;
        for (;;) {
            const double half_step_size = 0.5 * step_size;
            const double f2 = ray_eval(r, half_step_size);
            if (f2 <= f0 || --cap == 0) {
                *x1 = half_step_size; *f1o = f2; *x2 = step_size; *f2o = f1;
                return;
            }
            step_size = half_step_size;
            f1 = f2;
        }
    }
}

/* (qls::QuadraticLineSearch)(lse, f0, _)  legacy/DZOptimization.jl:191-216 */
static void quadratic_line_search(ray_t* r, double f0, double t1, int max_increases,
                                  double* xb_out, double* fb_out) {
    double x1, f1, x2, f2;
    three_point_bracket(r, f0, t1, max_increases, &x1, &f1, &x2, &f2);
    double xb = 0.0, fb = f0;
    if (f1 < fb) { xb = x1; fb = f1; }
    if (f2 < fb) { xb = x2; fb = f2; }
    const double delta_1 = f0 - f1;
    const double delta_2 = f2 - f1;
    const double sum_deltas = delta_1 + delta_2;
    if ((delta_1 >= 0.0) && (delta_2 >= 0.0) && (sum_deltas > 0.0)) {
        const double twice_delta_1 = delta_1 + delta_1;
        const double delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas);
        const double xq = delta_ratio * x1;
        const double fq = ray_eval(r, xq);
        if (fq < fb) { xb = xq; fb = fq; }
    }
    *xb_out = xb;
    *fb_out = fb;
}

int dzo_cpu_line_search(int objective, int constraint, int64_t obj_param, int order, int64_t n,
                        const double* x, const double* dir, double f0, double t1,
                        double* t_best, double* f_best) {
    problem_t P = {objective, constraint, order, n, obj_param > 0 ? obj_param : 1, 0, 0.0, 0.0, 0.0};
    double* w = (double*)malloc((size_t)(2 * n + 1) * sizeof(double));
    if (!w) return fail(DZO_ERR_ALLOC, "line_search: out of memory");
    ray_t r = {&P, x, dir, w, w + n, -1.0, 0};
    quadratic_line_search(&r, f0, t1, 0, t_best, f_best);
    free(w);
    return DZO_OK;
}

/* ======================================================================= BFGS */
struct dzo_cpu_bfgs {
    problem_t P;
    int64_t batch;
    int nthreads;
    double *x, *g, *dx, *dg, *d, *H, *f, *L;
    int64_t* iter;
    int32_t* type;
    uint8_t* term;
    double* scratch; /* per-thread: w, ref, t  (3n each) */
    int nscratch;
};

static int check_problem(int objective, int constraint, int64_t obj_param, int64_t n, int64_t batch) {
    if (n <= 0 || batch <= 0) return fail(DZO_ERR_INVALID_ARGUMENT, "n and batch must be positive");
    if (objective == DZO_OBJ_ROSENBROCK) {
        if (n % 2) return fail(DZO_ERR_INVALID_ARGUMENT, "extended Rosenbrock needs even n");
        if (constraint != DZO_CONSTRAINT_NONE)
            return fail(DZO_ERR_INVALID_ARGUMENT, "Rosenbrock takes DZO_CONSTRAINT_NONE");
    } else if (objective == DZO_OBJ_RIESZ) {
        if (obj_param <= 0 || obj_param > 16 || n % obj_param) return fail(DZO_ERR_INVALID_ARGUMENT, "Riesz needs n = dim * N, 1 <= dim <= 16");
        if (constraint != DZO_CONSTRAINT_NONE && constraint != DZO_CONSTRAINT_SPHERE)
            return fail(DZO_ERR_INVALID_ARGUMENT, "unknown constraint id");
    } else
        return fail(DZO_ERR_INVALID_ARGUMENT, "unknown objective id");
    return DZO_OK;
}

/* identity_matrix!  legacy/DZOptimization.jl:712-720 */
static void identity_(double* H, int64_t n) {
    for (int64_t j = 0; j < n; ++j)
        for (int64_t i = 0; i < n; ++i) H[i + j * n] = (i == j) ? 1.0 : 0.0;
}

int dzo_cpu_bfgs_create(dzo_cpu_bfgs** out, int objective, int constraint, int64_t obj_param,
                        int64_t n, int64_t batch, const double* x0, double initial_step_length,
                        int order, int nthreads) {
    if (!out || !x0) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    *out = NULL;
    int rc = check_problem(objective, constraint, obj_param, n, batch);
    if (rc) return rc;
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE)
        return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    if (nthreads < 1) nthreads = 1;
    dzo_cpu_bfgs* o = (dzo_cpu_bfgs*)calloc(1, sizeof *o);
    if (!o) return fail(DZO_ERR_ALLOC, "out of memory");
    o->P.objective = objective; o->P.constraint = constraint; o->P.order = order;
    o->P.n = n; o->P.dim = obj_param > 0 ? obj_param : 1;
    o->batch = batch; o->nthreads = nthreads;
    const size_t nb = (size_t)n * (size_t)batch;
    o->x = (double*)malloc(nb * 8); o->g = (double*)malloc(nb * 8);
    o->dx = (double*)calloc(nb, 8); o->dg = (double*)calloc(nb, 8); /* :777-778 */
    o->d = (double*)malloc(nb * 8);
    o->H = (double*)malloc(nb * (size_t)n * 8);
    o->f = (double*)malloc((size_t)batch * 8); o->L = (double*)malloc((size_t)batch * 8);
    o->iter = (int64_t*)calloc((size_t)batch, 8);            /* :767 */
    o->type = (int32_t*)calloc((size_t)batch, 4);            /* :780 NullStep */
    o->term = (uint8_t*)calloc((size_t)batch, 1);            /* :768 */
    o->nscratch = nthreads;
    o->scratch = (double*)malloc((size_t)nthreads * 3 * (size_t)n * 8);
    if (!o->x || !o->g || !o->dx || !o->dg || !o->d || !o->H || !o->f || !o->L || !o->iter ||
        !o->type || !o->term || !o->scratch) {
        dzo_cpu_bfgs_destroy(o);
        return fail(DZO_ERR_ALLOC, "out of memory");
    }
    memcpy(o->x, x0, nb * 8); /* :769 copy(initial_point) */
    int bad_constraint = 0, bad_nan = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1 && batch > 1) \
    reduction(| : bad_constraint, bad_nan)
    for (int64_t p = 0; p < batch; ++p) {
        double* x = o->x + p * n;
        if (!constraint_(&o->P, x)) bad_constraint |= 1;     /* :770-771 */
        const double f0 = objective_(&o->P, x);              /* :772 */
        if (isnan(f0)) bad_nan |= 1;                         /* :773 */
        o->f[p] = f0;
        gradient_(&o->P, o->g + p * n, x);                   /* :775-776 */
        o->L[p] = initial_step_length;                       /* :779 */
        identity_(o->H + (size_t)p * n * n, n);              /* :781-783 */
        memcpy(o->d + p * n, o->g + p * n, (size_t)n * 8);   /* :784 */
    }
    if (bad_constraint) { dzo_cpu_bfgs_destroy(o); return fail(DZO_ERR_CONSTRAINT_FAILED, "constraint_function! failed on the initial point"); }
    if (bad_nan) { dzo_cpu_bfgs_destroy(o); return fail(DZO_ERR_NAN_OBJECTIVE, "objective is NaN at the initial point"); }
    *out = o;
    return DZO_OK;
}

/* update_inverse_hessian!  legacy/DZOptimization.jl:864-889 (+ fused mul! :958-960) */
static void update_inverse_hessian_(int order, int64_t n, double* H, double step_length, double* sd,
                                    const double* dg, double* scratch, const double* next_g,
                                    double* next_d, int nthreads) {
    const double overlap = dot_(order, sd, dg, n);            /* :873 */
    const double inv_overlap = 1.0 / overlap;                 /* :874 inv(overlap) */
    for (int64_t i = 0; i < n; ++i) sd[i] *= inv_overlap;     /* :874 scalar_mul! */
    gemv_rows_(order, n, 0, n, H, n, dg, scratch, nthreads);  /* :875 */
    const double delta_norm = step_length * overlap + dot_(order, dg, scratch, n); /* :876 */
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
    for (int64_t j = 0; j < n; ++j) {                         /* :878-886 */
        const double sj = sd[j];
        const double tj = scratch[j];
        double* col = H + j * n;
        for (int64_t i = 0; i < n; ++i)
            col[i] += (delta_norm * (sd[i] * sj) - (scratch[i] * sj + sd[i] * tj));
    }
    if (next_g && next_d) gemv_rows_(order, n, 0, n, H, n, next_g, next_d, nthreads); /* :958-960 */
}

/* step!(opt::BFGSOptimizer)  legacy/DZOptimization.jl:891-994, one problem */
static void bfgs_step_one(dzo_cpu_bfgs* o, int64_t p, double* scr, int inner_threads) {
    if (o->term[p]) return;                                   /* :893 */
    const problem_t* P = &o->P;
    const int64_t n = P->n;
    double* point = o->x + p * n;
    double* delta_point = o->dx + p * n;
    double* gradient = o->g + p * n;
    double* delta_gradient = o->dg + p * n;
    double* bfgs_direction = o->d + p * n;
    double* H = o->H + (size_t)p * n * n;
    double* w = scr; double* ref = scr + n; double* tvec = scr + 2 * n;
    const double f0 = o->f[p];
    const double step_length = o->L[p];                       /* :918 */

    const double grad_norm = sqrt(norm2_(P->order, gradient, n));       /* :921 */
    ray_t rg = {P, point, gradient, w, ref, -1.0, 0};
    double grad_step_length, grad_obj;
    quadratic_line_search(&rg, f0, step_length / grad_norm, 0, &grad_step_length, &grad_obj); /* :922-925 */

    const double bfgs_norm = sqrt(norm2_(P->order, bfgs_direction, n)); /* :928 */
    ray_t rb = {P, point, bfgs_direction, w, ref, -1.0, 0};
    double bfgs_step_length, bfgs_obj;
    quadratic_line_search(&rb, f0, step_length / bfgs_norm, 0, &bfgs_step_length, &bfgs_obj); /* :929-932 */

    if (bfgs_obj < f0 && !(bfgs_obj > grad_obj)) {            /* :934 */
        o->f[p] = bfgs_obj;                                   /* :937 */
        o->L[p] = bfgs_step_length * bfgs_norm;               /* :938 */
        o->type[p] = DZO_STEP_BFGS;                           /* :939 */
        o->iter[p] += 1;                                      /* :940 */
        for (int64_t i = 0; i < n; ++i) delta_point[i] = -point[i];        /* :943 */
        for (int64_t i = 0; i < n; ++i) delta_gradient[i] = -gradient[i];  /* :944 */
        const double a = -bfgs_step_length;
        for (int64_t i = 0; i < n; ++i) point[i] += a * bfgs_direction[i]; /* :945 */
        constraint_(P, point);                                /* :946-947 */
        gradient_(P, gradient, point);                        /* :948 */
        for (int64_t i = 0; i < n; ++i) delta_point[i] += point[i];        /* :949 */
        for (int64_t i = 0; i < n; ++i) delta_gradient[i] += gradient[i];  /* :950 */
        update_inverse_hessian_(P->order, n, H, -bfgs_step_length, bfgs_direction, delta_gradient,
                                tvec, gradient, bfgs_direction, inner_threads); /* :953-960 */
    } else if (grad_obj < f0) {                               /* :962 */
        o->f[p] = grad_obj;                                   /* :965 */
        o->L[p] = grad_step_length * grad_norm;               /* :966 */
        o->type[p] = DZO_STEP_GRADIENT_DESCENT;               /* :967 */
        o->iter[p] += 1;                                      /* :968 */
        for (int64_t i = 0; i < n; ++i) delta_point[i] = -point[i];        /* :971 */
        for (int64_t i = 0; i < n; ++i) delta_gradient[i] = -gradient[i];  /* :972 */
        const double a = -grad_step_length;
        /* NB: direction is the OLD gradient; it is overwritten only at :976 */
        for (int64_t i = 0; i < n; ++i) point[i] += a * gradient[i];       /* :973 */
        constraint_(P, point);                                /* :974-975 */
        gradient_(P, gradient, point);                        /* :976 */
        for (int64_t i = 0; i < n; ++i) delta_point[i] += point[i];        /* :977 */
        for (int64_t i = 0; i < n; ++i) delta_gradient[i] += gradient[i];  /* :978 */
        identity_(H, n);                                      /* :981 */
        memcpy(bfgs_direction, gradient, (size_t)n * 8);      /* :984-986 */
    } else {
        o->term[p] = 1;                                       /* :989 */
    }
}

int dzo_cpu_bfgs_step(dzo_cpu_bfgs* o, int k) {
    if (!o || k < 0) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    const int64_t n = o->P.n;
    if (o->batch == 1) {
        for (int s = 0; s < k; ++s) bfgs_step_one(o, 0, o->scratch, o->nthreads);
        return DZO_OK;
    }
#pragma omp parallel num_threads(o->nthreads) if (o->nthreads > 1)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double* scr = o->scratch + (size_t)tid * 3 * n;
#pragma omp for schedule(dynamic, 64)
        for (int64_t p = 0; p < o->batch; ++p)
            for (int s = 0; s < k; ++s) bfgs_step_one(o, p, scr, 1);
    }
    return DZO_OK;
}

#define GETTER(name, field, type, count)                                                     \
    int name(dzo_cpu_bfgs* o, type* out) {                                                   \
        if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");               \
        memcpy(out, o->field, (size_t)(count) * sizeof(type));                               \
        return DZO_OK;                                                                       \
    }
GETTER(dzo_cpu_bfgs_get_point, x, double, o->P.n* o->batch)
GETTER(dzo_cpu_bfgs_get_gradient, g, double, o->P.n* o->batch)
GETTER(dzo_cpu_bfgs_get_delta_point, dx, double, o->P.n* o->batch)
GETTER(dzo_cpu_bfgs_get_delta_gradient, dg, double, o->P.n* o->batch)
GETTER(dzo_cpu_bfgs_get_direction, d, double, o->P.n* o->batch)
GETTER(dzo_cpu_bfgs_get_objective, f, double, o->batch)
GETTER(dzo_cpu_bfgs_get_step_length, L, double, o->batch)
GETTER(dzo_cpu_bfgs_get_step_type, type, int32_t, o->batch)
GETTER(dzo_cpu_bfgs_get_iteration_count, iter, int64_t, o->batch)
GETTER(dzo_cpu_bfgs_get_terminated, term, uint8_t, o->batch)
#undef GETTER

int dzo_cpu_bfgs_get_inverse_hessian(dzo_cpu_bfgs* o, int64_t problem, double* out) {
    if (!o || !out || problem < 0 || problem >= o->batch) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    const size_t nn = (size_t)o->P.n * (size_t)o->P.n;
    memcpy(out, o->H + (size_t)problem * nn, nn * 8);
    return DZO_OK;
}
int dzo_cpu_bfgs_count_active(dzo_cpu_bfgs* o, int64_t* out) {
    if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    int64_t c = 0;
    for (int64_t p = 0; p < o->batch; ++p) c += !o->term[p];
    *out = c;
    return DZO_OK;
}

/* BFGSOptimizer(T, f, g!, c!, opt)  legacy/DZOptimization.jl:819-862 with T == Float64 */
int dzo_cpu_bfgs_set_state(dzo_cpu_bfgs* o, const double* point, const double* inverse_hessian,
                           const double* delta_point, const double* delta_gradient,
                           const double* last_step_length, const int32_t* last_step_type,
                           const int64_t* iteration_count) {
    if (!o || !point || !inverse_hessian || !delta_point || !delta_gradient || !last_step_length ||
        !last_step_type || !iteration_count)
        return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    const int64_t n = o->P.n;
    const size_t nb = (size_t)n * (size_t)o->batch;
    memcpy(o->x, point, nb * 8);                               /* :825 */
    memcpy(o->H, inverse_hessian, nb * (size_t)n * 8);         /* :832 */
    memcpy(o->dx, delta_point, nb * 8);                        /* :853 */
    memcpy(o->dg, delta_gradient, nb * 8);                     /* :854 */
    int bad_nan = 0;
    for (int64_t p = 0; p < o->batch; ++p) {
        double* x = o->x + p * n;
        constraint_(&o->P, x);                                 /* :826-827 */
        o->f[p] = objective_(&o->P, x);                        /* :828 */
        if (isnan(o->f[p])) bad_nan = 1;                       /* :829 */
        gradient_(&o->P, o->g + p * n, x);                     /* :830-831 */
        gemv_rows_(o->P.order, n, 0, n, o->H + (size_t)p * n * n, n, o->g + p * n, o->d + p * n,
                   o->batch == 1 ? o->nthreads : 1);           /* :833-836 */
        o->iter[p] = iteration_count[p];                       /* :848 */
        o->term[p] = 0;                                        /* :849 */
        o->L[p] = last_step_length[p];                         /* :855 */
        o->type[p] = last_step_type[p];                        /* :856 */
    }
    if (bad_nan) return fail(DZO_ERR_NAN_OBJECTIVE, "objective is NaN at the restored point");
    return DZO_OK;
}

void dzo_cpu_bfgs_destroy(dzo_cpu_bfgs* o) {
    if (!o) return;
    free(o->x); free(o->g); free(o->dx); free(o->dg); free(o->d); free(o->H); free(o->f);
    free(o->L); free(o->iter); free(o->type); free(o->term); free(o->scratch);
    free(o);
}

/* ======================================================================= GradientDescent
 * legacy/DZOptimization.jl:305-449 */
struct dzo_cpu_gd {
    problem_t P;
    int64_t batch;
    int nthreads, max_increases;
    double *x, *dx, *g, *dg, *d, *f, *df, *L;
    int64_t* iter;
    uint8_t* term;
    double* scratch; /* per-thread new_point, reference_point (2n) */
};

int dzo_cpu_gd_create(dzo_cpu_gd** out, int objective, int constraint, int64_t obj_param, int64_t n,
                      int64_t batch, const double* x0, double initial_step_length, int max_increases,
                      int order, int nthreads) {
    if (!out || !x0) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    *out = NULL;
    int rc = check_problem(objective, constraint, obj_param, n, batch);
    if (rc) return rc;
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE && order != DZO_ORDER_TREE_BLOCKED)
        return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    if (nthreads < 1) nthreads = 1;
    dzo_cpu_gd* o = (dzo_cpu_gd*)calloc(1, sizeof *o);
    if (!o) return fail(DZO_ERR_ALLOC, "out of memory");
    o->P.objective = objective; o->P.constraint = constraint; o->P.order = order;
    o->P.n = n; o->P.dim = obj_param > 0 ? obj_param : 1;
    o->batch = batch; o->nthreads = nthreads; o->max_increases = max_increases;
    const size_t nb = (size_t)n * (size_t)batch;
    o->x = (double*)malloc(nb * 8); o->dx = (double*)calloc(nb, 8);
    o->g = (double*)malloc(nb * 8); o->dg = (double*)calloc(nb, 8);
    o->d = (double*)calloc(nb, 8);
    o->f = (double*)malloc((size_t)batch * 8); o->df = (double*)calloc((size_t)batch, 8);
    o->L = (double*)calloc((size_t)batch, 8);
    o->iter = (int64_t*)calloc((size_t)batch, 8);
    o->term = (uint8_t*)calloc((size_t)batch, 1);
    o->scratch = (double*)malloc((size_t)nthreads * 2 * (size_t)n * 8);
    if (!o->x || !o->dx || !o->g || !o->dg || !o->d || !o->f || !o->df || !o->L || !o->iter ||
        !o->term || !o->scratch) {
        dzo_cpu_gd_destroy(o);
        return fail(DZO_ERR_ALLOC, "out of memory");
    }
    memcpy(o->x, x0, nb * 8);                                  /* :339 collect */
    int bad_constraint = 0;
    for (int64_t p = 0; p < batch; ++p) {
        double* x = o->x + p * n;
        if (!constraint_(&o->P, x)) bad_constraint = 1;        /* :340 */
        const double f0 = objective_(&o->P, x);                /* :343 */
        o->f[p] = f0;
        gradient_(&o->P, o->g + p * n, x);                     /* :347-348 */
        const double inv_gradient_norm = 1.0 / sqrt(norm2_(order, o->g + p * n, n)); /* :352 */
        if (isfinite(inv_gradient_norm)) {                     /* :354-357 */
            const double alpha = -initial_step_length * inv_gradient_norm;
            for (int64_t i = 0; i < n; ++i) o->d[p * n + i] = o->g[p * n + i];
            for (int64_t i = 0; i < n; ++i) o->d[p * n + i] *= alpha;
        }
        o->term[p] = (!isfinite(f0)) || (!isfinite(inv_gradient_norm)); /* :364-366 */
    }
    if (bad_constraint) { dzo_cpu_gd_destroy(o); return fail(DZO_ERR_CONSTRAINT_FAILED, "constraint_function! failed on the initial point"); }
    *out = o;
    return DZO_OK;
}

/* step!(opt::GradientDescentOptimizer)  legacy/DZOptimization.jl:393-449 */
static void gd_step_one(dzo_cpu_gd* o, int64_t p, double* scr) {
    if (o->term[p]) return;                                    /* :402 */
    const problem_t* P = &o->P;
    const int64_t n = P->n;
    double* x = o->x + p * n; double* dx = o->dx + p * n;
    double* g = o->g + p * n; double* dg = o->dg + p * n; double* d = o->d + p * n;
    ray_t r = {P, x, d, scr, scr + n, +1.0, 0};
    double step_size, objective_value;
    quadratic_line_search(&r, o->f[p], 1.0, o->max_increases, &step_size, &objective_value); /* :405-407 */
    if (step_size == 0.0 || !(objective_value < o->f[p])) {   /* :410-414 */
        o->term[p] = 1;
        return;
    }
    o->iter[p] += 1;                                           /* :415 */
    memcpy(dx, x, (size_t)n * 8);                              /* :418 */
    for (int64_t i = 0; i < n; ++i) x[i] += step_size * d[i];  /* :419 */
    constraint_(P, x);                                         /* :420 */
    for (int64_t i = 0; i < n; ++i) dx[i] = x[i] - dx[i];      /* :423 delta! */
    const double step_length = sqrt(norm2_(P->order, dx, n));  /* :424 */
    o->L[p] = step_length;                                     /* :425 */
    o->df[p] = objective_value - o->f[p];                      /* :428-429 */
    o->f[p] = objective_value;                                 /* :430 */
    memcpy(dg, g, (size_t)n * 8);                              /* :433 */
    gradient_(P, g, x);                                        /* :434 */
    for (int64_t i = 0; i < n; ++i) dg[i] = g[i] - dg[i];      /* :435 */
    const double inv_gradient_norm = 1.0 / sqrt(norm2_(P->order, g, n)); /* :438 */
    if (!isfinite(inv_gradient_norm)) {                        /* :439-442 */
        o->term[p] = 1;
        return;
    }
    const double alpha = -step_length * inv_gradient_norm;
    for (int64_t i = 0; i < n; ++i) d[i] = alpha * g[i];       /* :445-446 scale!/4 */
}

int dzo_cpu_gd_step(dzo_cpu_gd* o, int k) {
    if (!o || k < 0) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    const int64_t n = o->P.n;
#pragma omp parallel num_threads(o->nthreads) if (o->nthreads > 1 && o->batch > 1)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double* scr = o->scratch + (size_t)tid * 2 * n;
#pragma omp for schedule(dynamic, 64)
        for (int64_t p = 0; p < o->batch; ++p)
            for (int s = 0; s < k; ++s) gd_step_one(o, p, scr);
    }
    return DZO_OK;
}

#define GETTER(name, field, type, count)                                                     \
    int name(dzo_cpu_gd* o, type* out) {                                                     \
        if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");               \
        memcpy(out, o->field, (size_t)(count) * sizeof(type));                               \
        return DZO_OK;                                                                       \
    }
GETTER(dzo_cpu_gd_get_point, x, double, o->P.n* o->batch)
GETTER(dzo_cpu_gd_get_delta_point, dx, double, o->P.n* o->batch)
GETTER(dzo_cpu_gd_get_gradient, g, double, o->P.n* o->batch)
GETTER(dzo_cpu_gd_get_delta_gradient, dg, double, o->P.n* o->batch)
GETTER(dzo_cpu_gd_get_direction, d, double, o->P.n* o->batch)
GETTER(dzo_cpu_gd_get_objective, f, double, o->batch)
GETTER(dzo_cpu_gd_get_delta_objective, df, double, o->batch)
GETTER(dzo_cpu_gd_get_step_length, L, double, o->batch)
GETTER(dzo_cpu_gd_get_iteration_count, iter, int64_t, o->batch)
GETTER(dzo_cpu_gd_get_terminated, term, uint8_t, o->batch)
#undef GETTER

void dzo_cpu_gd_destroy(dzo_cpu_gd* o) {
    if (!o) return;
    free(o->x); free(o->dx); free(o->g); free(o->dg); free(o->d); free(o->f); free(o->df);
    free(o->L); free(o->iter); free(o->term); free(o->scratch);
    free(o);
}

/* ======================================================================= kernel-level twins */
int dzo_cpu_objective(int objective, int constraint, int64_t obj_param, int order, int64_t n,
                      int64_t batch, const double* x, double* f) {
    int rc = check_problem(objective, constraint, obj_param, n, batch);
    if (rc) return rc;
    if (!x || !f) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    problem_t P = {objective, constraint, order, n, obj_param > 0 ? obj_param : 1, 0, 0.0, 0.0, 0.0};
    for (int64_t p = 0; p < batch; ++p) f[p] = objective_(&P, x + p * n);
    return DZO_OK;
}
int dzo_cpu_gradient(int objective, int constraint, int64_t obj_param, int order, int64_t n,
                     int64_t batch, const double* x, double* g) {
    int rc = check_problem(objective, constraint, obj_param, n, batch);
    if (rc) return rc;
    if (!x || !g) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    problem_t P = {objective, constraint, order, n, obj_param > 0 ? obj_param : 1, 0, 0.0, 0.0, 0.0};
    for (int64_t p = 0; p < batch; ++p) gradient_(&P, g + p * n, x + p * n);
    return DZO_OK;
}
int dzo_cpu_dot(int order, int64_t n, const double* v, const double* w, double* out) {
    if (!v || !w || !out || n < 0) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    *out = dot_(order, v, w, n);
    return DZO_OK;
}
int dzo_cpu_gemv(int order, int64_t n, const double* H, const double* v, double* out, int nthreads) {
    if (!H || !v || !out || n <= 0) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    gemv_rows_(order, n, 0, n, H, n, v, out, nthreads < 1 ? 1 : nthreads);
    return DZO_OK;
}
int dzo_cpu_update_inverse_hessian(int order, int64_t n, double* H, double step_length,
                                   double* step_direction, const double* delta_gradient,
                                   double* scratch, const double* next_gradient,
                                   double* next_direction, int nthreads) {
    if (!H || !step_direction || !delta_gradient || !scratch || n <= 0)
        return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    update_inverse_hessian_(order, n, H, step_length, step_direction, delta_gradient, scratch,
                            next_gradient, next_direction, nthreads < 1 ? 1 : nthreads);
    return DZO_OK;
}

/* ======================================================================= row-slab twins (sharded-mode model) */
int dzo_cpu_gemv_rows(int order, int64_t n, int64_t row_begin, int64_t row_end, const double* Hslab,
                      const double* v, double* out) {
    if (!Hslab || !v || !out || n <= 0 || row_begin < 0 || row_end > n || row_begin >= row_end)
        return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    const int64_t rows = row_end - row_begin;
    /* gemv_rows_ indexes the slab by local row: pass r0 = 0 and a leading dimension of `rows` */
    gemv_rows_(order, n, 0, rows, Hslab, rows, v, out, 1);
    return DZO_OK;
}
int dzo_cpu_update_rows(int64_t n, int64_t row_begin, int64_t row_end, double* Hslab, double delta_norm,
                        const double* sd, const double* scratch) {
    if (!Hslab || !sd || !scratch || n <= 0 || row_begin < 0 || row_end > n || row_begin >= row_end)
        return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    const int64_t rows = row_end - row_begin;
    for (int64_t j = 0; j < n; ++j) {                          /* legacy/DZOptimization.jl:878-886 */
        const double sj = sd[j], tj = scratch[j];
        double* col = Hslab + j * rows;
        for (int64_t i = 0; i < rows; ++i) {
            const int64_t gi = row_begin + i;
            col[i] += (delta_norm * (sd[gi] * sj) - (scratch[gi] * sj + sd[gi] * tj));
        }
    }
    return DZO_OK;
}

/* ======================================================================= pairwise radial kernels
 * src/ExampleFunctions.jl (the LIVE package).  muladd == fma (explicit); nothing else contracts. */
static inline double lj_energy(double r2) {            /* :16-27 */
    const double inv_r2 = 1.0 / r2;
    const double inv_r4 = inv_r2 * inv_r2;
    const double inv_r6 = inv_r4 * inv_r2;
    return 4.0 * fma(inv_r6, inv_r6, -inv_r6);
}
static inline double lj_first_derivative(double r2) {  /* :30-47 */
    const double inv_r2 = 1.0 / r2;
    const double inv_r4 = inv_r2 * inv_r2;
    const double inv_r6 = inv_r4 * inv_r2;
    const double inv_r8 = inv_r4 * inv_r4;
    return -12.0 * fma(inv_r8, inv_r6 + inv_r6, -inv_r8);
}
static inline double lj_second_derivative(double r2) { /* :50-72 */
    const double inv_r2 = 1.0 / r2;
    const double inv_r4 = inv_r2 * inv_r2;
    const double inv_r8 = inv_r4 * inv_r4;
    const double inv_r10 = inv_r8 * inv_r2;
    return 48.0 * fma(3.5, inv_r8 * inv_r8, -inv_r10);
}
static int pairwise_args_ok(int potential, int order, int64_t n) {
    if (potential != DZO_POT_LENNARD_JONES) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown radial potential id");
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    if (n <= 0) return fail(DZO_ERR_INVALID_ARGUMENT, "n must be positive");
    return DZO_OK;
}
/* One work-item of the three kernels (:131-147, :239-260, :389-421).  `what`: 0 energy, 1 gradient,
 * 2 hvp.  The j loop is one segment (SEQUENTIAL) or segments of DZO_RIESZ_SEG (TREE). */
static void pairwise_item(int what, int order, int64_t n, int64_t i, const double* x, const double* y, const double* z,
                          const double* u, const double* v, const double* w, double out[3]) {
    const int64_t seg = (order == DZO_ORDER_SEQUENTIAL) ? n : DZO_RIESZ_SEG;
    const double xi = x[i], yi = y[i], zi = z[i];
    const double ui = u ? u[i] : 0.0, vi = v ? v[i] : 0.0, wi = w ? w[i] : 0.0;
    double acc[3] = {0.0, 0.0, 0.0};
    for (int64_t s0 = 0; s0 < n; s0 += seg) {
        const int64_t s1 = (s0 + seg < n) ? s0 + seg : n;
        double ax = 0.0, ay = 0.0, az = 0.0;
        for (int64_t j = s0; j < s1; ++j) {
            const double dx = xi - x[j], dy = yi - y[j], dz = zi - z[j];
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (what == 0) {
                ax += (i == j) ? 0.0 : lj_energy(r2);                       /* :145 */
            } else if (what == 1) {
                const double f = (i == j) ? 0.0 : lj_first_derivative(r2);  /* :253 */
                ax += f * dx; ay += f * dy; az += f * dz;                   /* :254-256 */
            } else {
                const double du = ui - u[j], dv = vi - v[j], dw = wi - w[j];
                const double f = (i == j) ? 0.0 : lj_first_derivative(r2);  /* :409 */
                const double s = (i == j) ? 0.0 : lj_second_derivative(r2); /* :410 */
                const double overlap = dx * du + dy * dv + dz * dw;         /* :411 */
                const double os = overlap * s;
                const double g = os + os;                                   /* :413 twice(overlap * s) */
                ax += f * du + g * dx; ay += f * dv + g * dy; az += f * dw + g * dz;   /* :416-418 */
            }
        }
        if (s0 == 0) { acc[0] = ax; acc[1] = ay; acc[2] = az; }
        else { acc[0] += ax; acc[1] += ay; acc[2] += az; }
    }
    if (what == 0) out[0] = 0.5 * acc[0];                                   /* :148 */
    else { out[0] = acc[0] + acc[0]; out[1] = acc[1] + acc[1]; out[2] = acc[2] + acc[2]; }   /* :258-260, :419-421 */
}
int dzo_cpu_pairwise_energy(int potential, int order, int64_t n, const double* x, const double* y, const double* z,
                            double* point_energies, double* energy) {
    int rc = pairwise_args_ok(potential, order, n);
    if (rc) return rc;
    if (!x || !y || !z || !energy) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    double part[DZO_TREE_WIDTH];
    for (int k = 0; k < DZO_TREE_WIDTH; ++k) part[k] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double o[3];
        pairwise_item(0, order, n, i, x, y, z, NULL, NULL, NULL, o);
        if (point_energies) point_energies[i] = o[0];
        part[i & (DZO_TREE_WIDTH - 1)] += o[0];                            /* [GLUE] sum(point_energies) :172 */
    }
    *energy = tree_combine(part);
    return DZO_OK;
}
int dzo_cpu_pairwise_gradient(int potential, int order, int64_t n, const double* x, const double* y, const double* z,
                              double* gx, double* gy, double* gz) {
    int rc = pairwise_args_ok(potential, order, n);
    if (rc) return rc;
    if (!x || !y || !z || !gx || !gy || !gz) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double o[3];
        pairwise_item(1, order, n, i, x, y, z, NULL, NULL, NULL, o);
        gx[i] = o[0]; gy[i] = o[1]; gz[i] = o[2];
    }
    return DZO_OK;
}
int dzo_cpu_pairwise_hvp(int potential, int order, int64_t n, const double* x, const double* y, const double* z,
                         const double* u, const double* v, const double* w, double* px, double* py, double* pz) {
    int rc = pairwise_args_ok(potential, order, n);
    if (rc) return rc;
    if (!x || !y || !z || !u || !v || !w || !px || !py || !pz) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double o[3];
        pairwise_item(2, order, n, i, x, y, z, u, v, w, o);
        px[i] = o[0]; py[i] = o[1]; pz[i] = o[2];
    }
    return DZO_OK;
}

/* ======================================================================= live LBFGSOptimizer
 * src/DZOptimization.jl:107-154 (take_backtracking_step!), :321-509.  History index 0 = newest
 * (the reference pushfirst!s).  dot / norm: BLAS in the reference -> dot_ / sqrt(norm2_) here. */
struct dzo_cpu_lbfgs {
    problem_t P;
    int m, count;
    double *x, *dx, *g, *dg, *d;
    double *S, *Y;                 /* m x n each, logical slot i at S + i*n (shifted on push) */
    double rho[DZO_LBFGS_MAX_HISTORY], alpha[DZO_LBFGS_MAX_HISTORY];
    double f, df;
    int64_t iter;
    int stuck;
};

static int julia_isequal(double a, double b) { /* isequal: NaN == NaN, -0.0 != 0.0 */
    if (a == b) return signbit(a) == signbit(b);
    return (a != a) && (b != b);
}

int dzo_cpu_lbfgs_create(dzo_cpu_lbfgs** out, int objective, int constraint, int64_t obj_param, int64_t n,
                         const double* x0, double initial_step_length, int history_length, int order) {
    if (!out || !x0) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    *out = NULL;
    int rc = check_problem(objective, constraint, obj_param, n, 1);
    if (rc) return rc;
    if (history_length < 1 || history_length > DZO_LBFGS_MAX_HISTORY) return fail(DZO_ERR_INVALID_ARGUMENT, "history_length must be in [1, 64]");
    if (!(initial_step_length > 0.0)) return fail(DZO_ERR_INVALID_ARGUMENT, "initial_step_length must be positive"); /* :375 */
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE && order != DZO_ORDER_TREE_BLOCKED) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    dzo_cpu_lbfgs* o = (dzo_cpu_lbfgs*)calloc(1, sizeof *o);
    if (!o) return fail(DZO_ERR_ALLOC, "out of memory");
    o->P.objective = objective; o->P.constraint = constraint; o->P.order = order; o->P.n = n;
    o->P.dim = obj_param > 0 ? obj_param : 1;
    o->m = history_length;
    o->x = (double*)malloc((size_t)n * 8); o->dx = (double*)calloc((size_t)n, 8);       /* :364-367 */
    o->g = (double*)malloc((size_t)n * 8); o->dg = (double*)calloc((size_t)n, 8);       /* :369-372 */
    o->d = (double*)calloc((size_t)n, 8);
    o->S = (double*)calloc((size_t)n * (size_t)(history_length > 0 ? history_length : 1), 8);
    o->Y = (double*)calloc((size_t)n * (size_t)(history_length > 0 ? history_length : 1), 8);
    if (!o->x || !o->dx || !o->g || !o->dg || !o->d || !o->S || !o->Y) { dzo_cpu_lbfgs_destroy(o); return fail(DZO_ERR_ALLOC, "out of memory"); }
    memcpy(o->x, x0, (size_t)n * 8);
    if (!constraint_(&o->P, o->x)) { dzo_cpu_lbfgs_destroy(o); return fail(DZO_ERR_CONSTRAINT_FAILED, "constraint_function! failed on the initial point"); } /* :413-415 */
    o->f = objective_(&o->P, o->x);                                                     /* :417 */
    gradient_(&o->P, o->g, o->x);                                                       /* :419-422 */
    const double gnorm = sqrt(norm2_(order, o->g, n));                                  /* :376 */
    o->stuck = (gnorm == 0.0);                                                          /* :377 */
    if (!o->stuck) {                                                                    /* :380-383 */
        const double c = -initial_step_length / gnorm;
        for (int64_t i = 0; i < n; ++i) o->d[i] = o->g[i] * c;
    }
    *out = o;
    return DZO_OK;
}

/* compute_lbfgs_step_direction!  :430-451 */
static void lbfgs_direction(dzo_cpu_lbfgs* o) {
    const int64_t n = o->P.n;
    const int order = o->P.order;
    memcpy(o->d, o->g, (size_t)n * 8);
    for (int i = 0; i < o->count; ++i) {
        o->alpha[i] = dot_(order, o->S + (size_t)i * n, o->d, n) / o->rho[i];           /* :439 */
        const double a = -o->alpha[i];
        const double* y = o->Y + (size_t)i * n;
        for (int64_t k = 0; k < n; ++k) o->d[k] += a * y[k];                            /* :440 axpy! */
    }
    if (o->count > 0) {
        const double c = -o->rho[0] / dot_(order, o->Y, o->Y, n);                       /* :443 */
        for (int64_t k = 0; k < n; ++k) o->d[k] *= c;
    }
    for (int i = o->count - 1; i >= 0; --i) {
        const double beta = dot_(order, o->Y + (size_t)i * n, o->d, n) / o->rho[i];     /* :446 */
        const double a = -(o->alpha[i] + beta);
        const double* s = o->S + (size_t)i * n;
        for (int64_t k = 0; k < n; ++k) o->d[k] += a * s[k];                            /* :447 */
    }
}

static void lbfgs_step_one(dzo_cpu_lbfgs* o) {
    if (o->stuck) return;                                                               /* :456-458 */
    const int64_t n = o->P.n;
    const int order = o->P.order;
    if (o->iter > 0) lbfgs_direction(o);                                                /* :463-471 */
    /* take_backtracking_step!(opt, 1, step_direction)  :107-154 */
    double step = 1.0;
    memcpy(o->dx, o->x, (size_t)n * 8);                                                 /* :118 */
    for (;;) {
        int same = 1;
        for (int64_t k = 0; k < n; ++k) {
            o->x[k] += step * o->d[k];                                                  /* :124 axpy! */
            same &= julia_isequal(o->x[k], o->dx[k]);
        }
        if (same) { o->stuck = 1; return; }                                             /* :128-131 */
        if (constraint_(&o->P, o->x)) {                                                 /* :134-135 */
            const double next = objective_(&o->P, o->x);                                /* :138 */
            if (next < o->f) {                                                          /* :139 */
                o->df = next - o->f;                                                    /* :142-143 */
                o->f = next;
                for (int64_t k = 0; k < n; ++k) o->dx[k] = 1.0 * o->x[k] + (-1.0) * o->dx[k];   /* :145 axpby! */
                break;
            }
        }
        memcpy(o->x, o->dx, (size_t)n * 8);                                             /* :151 */
        step *= 0.5;                                                                    /* :152 */
    }
    memcpy(o->dg, o->g, (size_t)n * 8);                                                 /* :478 */
    gradient_(&o->P, o->g, o->x);                                                       /* :479 */
    for (int64_t k = 0; k < n; ++k) o->dg[k] = 1.0 * o->g[k] + (-1.0) * o->dg[k];       /* :480 */
    if (o->m > 0) {                                                                     /* :482-505 pushfirst! */
        const int keep = (o->count < o->m) ? o->count : o->m - 1;
        memmove(o->S + n, o->S, (size_t)keep * n * 8);
        memmove(o->Y + n, o->Y, (size_t)keep * n * 8);
        memmove(o->rho + 1, o->rho, (size_t)keep * 8);
        memcpy(o->S, o->dx, (size_t)n * 8);
        memcpy(o->Y, o->dg, (size_t)n * 8);
        o->rho[0] = dot_(order, o->dx, o->dg, n);                                       /* :505 */
        o->count = keep + 1;
    }
    o->iter += 1;                                                                       /* :507 */
}

int dzo_cpu_lbfgs_step(dzo_cpu_lbfgs* o, int k) {
    if (!o || k < 0) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    for (int s = 0; s < k; ++s) lbfgs_step_one(o);
    return DZO_OK;
}
#define LGET(name, field)                                                               \
    int name(dzo_cpu_lbfgs* o, double* out) {                                           \
        if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");          \
        memcpy(out, o->field, (size_t)o->P.n * 8);                                      \
        return DZO_OK;                                                                  \
    }
LGET(dzo_cpu_lbfgs_get_point, x)
LGET(dzo_cpu_lbfgs_get_delta_point, dx)
LGET(dzo_cpu_lbfgs_get_gradient, g)
LGET(dzo_cpu_lbfgs_get_delta_gradient, dg)
LGET(dzo_cpu_lbfgs_get_direction, d)
#undef LGET
int dzo_cpu_lbfgs_get_objective(dzo_cpu_lbfgs* o, double* out) { if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer"); *out = o->f; return DZO_OK; }
int dzo_cpu_lbfgs_get_delta_objective(dzo_cpu_lbfgs* o, double* out) { if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer"); *out = o->df; return DZO_OK; }
int dzo_cpu_lbfgs_get_iteration_count(dzo_cpu_lbfgs* o, int64_t* out) { if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer"); *out = o->iter; return DZO_OK; }
int dzo_cpu_lbfgs_get_stuck(dzo_cpu_lbfgs* o, uint8_t* out) { if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer"); *out = (uint8_t)o->stuck; return DZO_OK; }
int dzo_cpu_lbfgs_get_rho_history(dzo_cpu_lbfgs* o, int64_t* count, double* rho) {
    if (!o || !count || !rho) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    *count = o->count;
    for (int i = 0; i < DZO_LBFGS_MAX_HISTORY; ++i) rho[i] = (i < o->count) ? o->rho[i] : 0.0;
    return DZO_OK;
}
void dzo_cpu_lbfgs_destroy(dzo_cpu_lbfgs* o) {
    if (!o) return;
    free(o->x); free(o->dx); free(o->g); free(o->dg); free(o->d); free(o->S); free(o->Y);
    free(o);
}

/* ======================================================================= live AdGDOptimizer
 * src/DZOptimization.jl:179-312 */
struct dzo_cpu_adgd {
    problem_t P;
    double *x, *dx, *g, *dg;
    double f, df, cur, prev;
    int64_t iter;
    int stuck;
};
static double julia_min(double a, double b) { /* Base.min for Float64: NaN-propagating, min(-0.0, 0.0) = -0.0 */
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? a : b;
    return a < b ? a : b;
}
int dzo_cpu_adgd_create(dzo_cpu_adgd** out, int objective, int constraint, int64_t obj_param, int64_t n,
                        const double* x0, double initial_step_length, int order) {
    if (!out || !x0) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    *out = NULL;
    int rc = check_problem(objective, constraint, obj_param, n, 1);
    if (rc) return rc;
    if (!(initial_step_length > 0.0)) return fail(DZO_ERR_INVALID_ARGUMENT, "initial_step_length must be positive"); /* :232 */
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE && order != DZO_ORDER_TREE_BLOCKED) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    dzo_cpu_adgd* o = (dzo_cpu_adgd*)calloc(1, sizeof *o);
    if (!o) return fail(DZO_ERR_ALLOC, "out of memory");
    o->P.objective = objective; o->P.constraint = constraint; o->P.order = order; o->P.n = n;
    o->P.dim = obj_param > 0 ? obj_param : 1;
    o->x = (double*)malloc((size_t)n * 8); o->dx = (double*)calloc((size_t)n, 8);
    o->g = (double*)malloc((size_t)n * 8); o->dg = (double*)calloc((size_t)n, 8);
    if (!o->x || !o->dx || !o->g || !o->dg) { dzo_cpu_adgd_destroy(o); return fail(DZO_ERR_ALLOC, "out of memory"); }
    memcpy(o->x, x0, (size_t)n * 8);
    if (!constraint_(&o->P, o->x)) { dzo_cpu_adgd_destroy(o); return fail(DZO_ERR_CONSTRAINT_FAILED, "constraint_function! failed on the initial point"); }
    o->f = objective_(&o->P, o->x);                                                     /* :260 */
    gradient_(&o->P, o->g, o->x);                                                       /* :265 */
    const double gnorm = sqrt(norm2_(order, o->g, n));                                  /* :233 */
    o->stuck = (gnorm == 0.0);                                                          /* :234 */
    o->cur = o->prev = o->stuck ? 0.0 : initial_step_length / gnorm;                    /* :235-236 */
    *out = o;
    return DZO_OK;
}
static void adgd_step_one(dzo_cpu_adgd* o) {
    if (o->stuck) return;                                                               /* :276-278 */
    const int64_t n = o->P.n;
    const int order = o->P.order;
    const double inv_sqrt_two = sqrt(0.5);
    const double previous = o->prev, current = o->cur;
    double next = current;
    if (o->iter > 0) {                                                                  /* :288-297 */
        const double theta = current / previous;
        next *= sqrt(1.0 + theta);
        const double dgn = sqrt(norm2_(order, o->dg, n));
        if (dgn != 0.0) {
            const double inv_L = sqrt(norm2_(order, o->dx, n)) / dgn;
            next = julia_min(next, inv_sqrt_two * inv_L);
        }
    }
    o->prev = current;                                                                  /* :298-299 */
    o->cur = next;
    /* take_backtracking_step!(opt, -next_step_size, opt.current_gradient)  :301, :107-154 */
    double step = -next;
    memcpy(o->dx, o->x, (size_t)n * 8);
    for (;;) {
        int same = 1;
        for (int64_t k = 0; k < n; ++k) {
            o->x[k] += step * o->g[k];
            same &= julia_isequal(o->x[k], o->dx[k]);
        }
        if (same) { o->stuck = 1; return; }
        if (constraint_(&o->P, o->x)) {
            const double nxt = objective_(&o->P, o->x);
            if (nxt < o->f) {
                o->df = nxt - o->f;
                o->f = nxt;
                for (int64_t k = 0; k < n; ++k) o->dx[k] = 1.0 * o->x[k] + (-1.0) * o->dx[k];
                break;
            }
        }
        memcpy(o->x, o->dx, (size_t)n * 8);
        step *= 0.5;
    }
    memcpy(o->dg, o->g, (size_t)n * 8);                                                 /* :306 */
    gradient_(&o->P, o->g, o->x);                                                       /* :307 */
    for (int64_t k = 0; k < n; ++k) o->dg[k] = 1.0 * o->g[k] + (-1.0) * o->dg[k];       /* :308 */
    o->iter += 1;                                                                       /* :310 */
}
int dzo_cpu_adgd_step(dzo_cpu_adgd* o, int k) {
    if (!o || k < 0) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    for (int s = 0; s < k; ++s) adgd_step_one(o);
    return DZO_OK;
}
#define AGET(name, field)                                                               \
    int name(dzo_cpu_adgd* o, double* out) {                                            \
        if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");          \
        memcpy(out, o->field, (size_t)o->P.n * 8);                                      \
        return DZO_OK;                                                                  \
    }
AGET(dzo_cpu_adgd_get_point, x)
AGET(dzo_cpu_adgd_get_delta_point, dx)
AGET(dzo_cpu_adgd_get_gradient, g)
AGET(dzo_cpu_adgd_get_delta_gradient, dg)
#undef AGET
int dzo_cpu_adgd_get_scalars(dzo_cpu_adgd* o, double* s) {
    if (!o || !s) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    s[0] = o->f; s[1] = o->df; s[2] = o->cur; s[3] = o->prev; s[4] = (double)o->iter; s[5] = (double)o->stuck;
    return DZO_OK;
}
void dzo_cpu_adgd_destroy(dzo_cpu_adgd* o) {
    if (!o) return;
    free(o->x); free(o->dx); free(o->g); free(o->dg);
    free(o);
}

/* ======================================================================= legacy LBFGSOptimizer
 * legacy/DZOptimization.jl:458-695 (struct :458-486, ctor :489-548, step! :565-695): QuadraticLineSearch along the stored
 * direction, retry along the rescaled gradient when it fails (:589-610, history reset), cyclic
 * history buffer (column c = (iteration_count - 1) mod m, :641-643), two-loop correction exactly as
 * written (:656-680 -- note `axpy!(d, alpha, dg_c)` ADDS alpha*dg, unlike the textbook recursion;
 * the descent check :683-692 then falls back to the scaled negative gradient), natural step size
 * delta_overlap / norm2(delta_gradient) (:669-670). */
struct dzo_cpu_legacy_lbfgs {
    problem_t P;
    int m, max_increases;
    double *x, *dx, *g, *dg, *d, *S, *Y, *scratch;
    double rho[DZO_LBFGS_MAX_HISTORY], alpha[DZO_LBFGS_MAX_HISTORY];
    double f, df, L;
    int64_t iter, hist_count;
    int term;
};

static int check_decor(int objective, int decor, double l2_lambda, double lo, double hi) {
    if (decor & ~(DZO_DECOR_L2 | DZO_DECOR_BOX)) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown decorator bits");
    if (decor && objective != DZO_OBJ_ROSENBROCK) return fail(DZO_ERR_UNSUPPORTED, "decorators: DZO_OBJ_ROSENBROCK only");
    if ((decor & DZO_DECOR_L2) && l2_lambda != l2_lambda) return fail(DZO_ERR_INVALID_ARGUMENT, "lambda is NaN");
    if ((decor & DZO_DECOR_BOX) && !(lo <= hi)) return fail(DZO_ERR_INVALID_ARGUMENT, "box needs lower_bound <= upper_bound");
    return DZO_OK;
}

int dzo_cpu_legacy_lbfgs_create(dzo_cpu_legacy_lbfgs** out, int objective, int constraint, int64_t obj_param, int64_t n,
                                const double* x0, double initial_step_length, int history_length, int max_increases,
                                int decor, double l2_lambda, double box_lower, double box_upper, int order) {
    if (!out || !x0) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    *out = NULL;
    int rc = check_problem(objective, constraint, obj_param, n, 1);
    if (rc) return rc;
    rc = check_decor(objective, decor, l2_lambda, box_lower, box_upper);
    if (rc) return rc;
    if (history_length < 1 || history_length > DZO_LBFGS_MAX_HISTORY) return fail(DZO_ERR_INVALID_ARGUMENT, "history_length must be in [1, 64]"); /* :528 */
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE && order != DZO_ORDER_TREE_BLOCKED) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    dzo_cpu_legacy_lbfgs* o = (dzo_cpu_legacy_lbfgs*)calloc(1, sizeof *o);
    if (!o) return fail(DZO_ERR_ALLOC, "out of memory");
    o->P.objective = objective; o->P.constraint = constraint; o->P.order = order; o->P.n = n;
    o->P.dim = obj_param > 0 ? obj_param : 1;
    o->P.decor = decor; o->P.l2_lambda = l2_lambda; o->P.box_lo = box_lower; o->P.box_hi = box_upper;
    o->m = history_length; o->max_increases = max_increases;
    o->x = (double*)malloc((size_t)n * 8); o->dx = (double*)calloc((size_t)n, 8);
    o->g = (double*)malloc((size_t)n * 8); o->dg = (double*)calloc((size_t)n, 8);
    o->d = (double*)calloc((size_t)n, 8);
    o->S = (double*)calloc((size_t)n * (size_t)history_length, 8);
    o->Y = (double*)calloc((size_t)n * (size_t)history_length, 8);
    o->scratch = (double*)malloc((size_t)n * 2 * 8);
    if (!o->x || !o->dx || !o->g || !o->dg || !o->d || !o->S || !o->Y || !o->scratch) { dzo_cpu_legacy_lbfgs_destroy(o); return fail(DZO_ERR_ALLOC, "out of memory"); }
    memcpy(o->x, x0, (size_t)n * 8);                                                    /* :499 collect */
    if (!constraint_(&o->P, o->x)) { dzo_cpu_legacy_lbfgs_destroy(o); return fail(DZO_ERR_CONSTRAINT_FAILED, "constraint_function! failed on the initial point"); } /* :500 */
    o->f = objective_(&o->P, o->x);                                                     /* :503 */
    gradient_(&o->P, o->g, o->x);                                                       /* :507-508 */
    const double inv_gradient_norm = 1.0 / sqrt(norm2_(order, o->g, n));                /* :512 */
    if (isfinite(inv_gradient_norm)) {                                                  /* :514-517 */
        const double a = -initial_step_length * inv_gradient_norm;
        for (int64_t i = 0; i < n; ++i) o->d[i] = o->g[i];
        for (int64_t i = 0; i < n; ++i) o->d[i] *= a;
    }
    o->term = (!isfinite(o->f)) || (!isfinite(inv_gradient_norm));                      /* :524-526 */
    *out = o;
    return DZO_OK;
}

static int legacy_search_(dzo_cpu_legacy_lbfgs* o, double* step_size, double* objective_value) {
    ray_t r = {&o->P, o->x, o->d, o->scratch, o->scratch + o->P.n, +1.0, 0};
    quadratic_line_search(&r, o->f, 1.0, o->max_increases, step_size, objective_value);
    return !(*step_size == 0.0 || !(*objective_value < o->f));                          /* :589-590 */
}

static void legacy_lbfgs_step_one(dzo_cpu_legacy_lbfgs* o) {
    if (o->term) return;                                                                /* :578 */
    const problem_t* P = &o->P;
    const int64_t n = P->n;
    const int order = P->order, m = o->m;
    double *x = o->x, *dx = o->dx, *g = o->g, *dg = o->dg, *d = o->d;
    double step_size, objective_value;
    if (!legacy_search_(o, &step_size, &objective_value)) {                             /* :584-590 */
        const double a = -o->L * (1.0 / sqrt(norm2_(order, g, n)));                     /* :594-595 */
        for (int64_t i = 0; i < n; ++i) d[i] = g[i];                                    /* :593 */
        for (int64_t i = 0; i < n; ++i) d[i] *= a;
        if (!legacy_search_(o, &step_size, &objective_value)) {                         /* :596-605 */
            o->term = 1;
            return;
        }
        o->hist_count = 0;                                                              /* :609 */
    }
    o->iter += 1;                                                                       /* :611 */
    memcpy(dx, x, (size_t)n * 8);                                                       /* :614 */
    for (int64_t i = 0; i < n; ++i) x[i] += step_size * d[i];                           /* :615 */
    constraint_(P, x);                                                                  /* :616 */
    for (int64_t i = 0; i < n; ++i) dx[i] = x[i] - dx[i];                               /* :619 */
    const double step_length = sqrt(norm2_(order, dx, n));                              /* :620-621 */
    o->L = step_length;
    o->df = objective_value - o->f;                                                     /* :624-626 */
    o->f = objective_value;
    memcpy(dg, g, (size_t)n * 8);                                                       /* :629 */
    gradient_(P, g, x);                                                                 /* :630 */
    for (int64_t i = 0; i < n; ++i) dg[i] = g[i] - dg[i];                               /* :631 */
    const double inv_gradient_norm = 1.0 / sqrt(norm2_(order, g, n));                   /* :634 */
    if (!isfinite(inv_gradient_norm)) {                                                 /* :635-638 */
        o->term = 1;
        return;
    }
    int c = (int)((o->iter - 1) % m);                                                   /* :641 (0-based) */
    memcpy(o->S + (size_t)c * n, dx, (size_t)n * 8);                                    /* :642-643 */
    memcpy(o->Y + (size_t)c * n, dg, (size_t)n * 8);
    const double delta_overlap = dot_(order, dx, dg, n);                                /* :646 */
    o->rho[c] = 1.0 / delta_overlap;                                                    /* :647 */
    const int64_t hist_count = (o->hist_count + 1 < m) ? o->hist_count + 1 : m;         /* :650-653 */
    const int64_t hist_end = o->iter, hist_begin = hist_end - hist_count + 1;
    o->hist_count = hist_count;
    memcpy(d, g, (size_t)n * 8);                                                        /* :656 */
    for (int64_t it = hist_end; it >= hist_begin; --it) {                               /* :659-666 */
        c = (int)((it - 1) % m);
        const double alpha = o->rho[c] * dot_(order, d, o->S + (size_t)c * n, n);
        o->alpha[c] = alpha;
        const double* y = o->Y + (size_t)c * n;
        for (int64_t i = 0; i < n; ++i) d[i] += alpha * y[i];
    }
    const double gamma = delta_overlap / norm2_(order, dg, n);                          /* :669-670 */
    for (int64_t i = 0; i < n; ++i) d[i] *= gamma;
    for (int64_t it = hist_begin; it <= hist_end; ++it) {                               /* :673-680 */
        c = (int)((it - 1) % m);
        const double beta = o->alpha[c] - o->rho[c] * dot_(order, d, o->Y + (size_t)c * n, n);
        const double* s = o->S + (size_t)c * n;
        for (int64_t i = 0; i < n; ++i) d[i] += beta * s[i];
    }
    for (int64_t i = 0; i < n; ++i) d[i] = -d[i];                                       /* :683-684 negate! */
    const double gradient_overlap = dot_(order, d, g, n);
    if (!isfinite(gradient_overlap)) {                                                  /* :687-688 */
        o->term = 1;
    } else if (gradient_overlap >= 0.0) {                                               /* :689-692 */
        const double a = -step_length * inv_gradient_norm;
        for (int64_t i = 0; i < n; ++i) d[i] = a * g[i];
    }
}

int dzo_cpu_legacy_lbfgs_step(dzo_cpu_legacy_lbfgs* o, int k) {
    if (!o || k < 0) return fail(DZO_ERR_INVALID_ARGUMENT, "bad arguments");
    for (int s = 0; s < k; ++s) legacy_lbfgs_step_one(o);
    return DZO_OK;
}
#define LLGET(name, field)                                                              \
    int name(dzo_cpu_legacy_lbfgs* o, double* out) {                                    \
        if (!o || !out) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");          \
        memcpy(out, o->field, (size_t)o->P.n * 8);                                      \
        return DZO_OK;                                                                  \
    }
LLGET(dzo_cpu_legacy_lbfgs_get_point, x)
LLGET(dzo_cpu_legacy_lbfgs_get_delta_point, dx)
LLGET(dzo_cpu_legacy_lbfgs_get_gradient, g)
LLGET(dzo_cpu_legacy_lbfgs_get_delta_gradient, dg)
LLGET(dzo_cpu_legacy_lbfgs_get_direction, d)
#undef LLGET
int dzo_cpu_legacy_lbfgs_get_scalars(dzo_cpu_legacy_lbfgs* o, double* s) {
    if (!o || !s) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    s[0] = o->f; s[1] = o->df; s[2] = o->L; s[3] = (double)o->iter; s[4] = (double)o->term; s[5] = (double)o->hist_count;
    return DZO_OK;
}
int dzo_cpu_legacy_lbfgs_get_history(dzo_cpu_legacy_lbfgs* o, double* rho, double* alpha) {
    if (!o || !rho || !alpha) return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    for (int i = 0; i < o->m; ++i) { rho[i] = o->rho[i]; alpha[i] = o->alpha[i]; }
    return DZO_OK;
}
void dzo_cpu_legacy_lbfgs_destroy(dzo_cpu_legacy_lbfgs* o) {
    if (!o) return;
    free(o->x); free(o->dx); free(o->g); free(o->dg); free(o->d); free(o->S); free(o->Y); free(o->scratch);
    free(o);
}

/* ======================================================================= live LineSearchEvaluator
 * (lse::LineSearchEvaluator)(step_size, compute_gradient)  src/DZOptimization.jl:66-92.  LinearAlgebra.axpy! / dot
 * are BLAS in the reference -> un-fused x + step*dir and dot_ here.  results3 = { f_new (:81), improvement_ratio
 * (:85), slope_ratio (:90; 0 when compute_gradient is false -- the reference leaves the field untouched) }.
 * A failing constraint (:72-79) cannot occur with the device objectives (NONE, or SPHERE which returns true). */
int dzo_cpu_line_search_evaluate(int objective, int constraint, int64_t obj_param, int order, int64_t n,
                                 const double* x, double f_old, const double* dir, double overlap, double step_size,
                                 int compute_gradient, double* trial_point, double* trial_gradient, double* results3) {
    if (!x || !dir || !trial_point || !results3 || (compute_gradient && !trial_gradient))
        return fail(DZO_ERR_INVALID_ARGUMENT, "null pointer");
    int rc = check_problem(objective, constraint, obj_param, n, 1);
    if (rc) return rc;
    if (order != DZO_ORDER_SEQUENTIAL && order != DZO_ORDER_TREE) return fail(DZO_ERR_INVALID_ARGUMENT, "unknown summation order");
    problem_t P = {objective, constraint, order, n, obj_param > 0 ? obj_param : 1, 0, 0.0, 0.0, 0.0};
    for (int64_t i = 0; i < n; ++i) trial_point[i] = x[i];                              /* :70 copy! */
    for (int64_t i = 0; i < n; ++i) trial_point[i] += step_size * dir[i];               /* :71 axpy! */
    if (!constraint_(&P, trial_point)) {                                                /* :72-79 */
        results3[0] = INFINITY; results3[1] = -INFINITY; results3[2] = INFINITY;
        return DZO_OK;
    }
    const double f_new = objective_(&P, trial_point);                                   /* :81 */
    results3[0] = f_new;
    results3[1] = (f_new - f_old) / (step_size * overlap);                              /* :85 */
    results3[2] = 0.0;
    if (compute_gradient) {
        gradient_(&P, trial_gradient, trial_point);                                     /* :88 */
        results3[2] = dot_(order, trial_gradient, dir, n) / overlap;                    /* :89-90 */
    }
    return DZO_OK;
}
