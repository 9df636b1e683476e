// large_bfgs.cuh -- single large-n BFGS / GD step! as a chain of device-controlled kernels.
//
// One step! (legacy/DZOptimization.jl:891-994) is five asynchronous launches on
// one stream with NO host round trip; every kernel reads the LargeCtrl block to learn what
// the step turned out to be and exits at once when it has nothing to do:
//
//   vec_bfgs_search_kernel   1 CTA x 1024   norms, both line searches, the BFGS/GD/terminate
//                                           decision, x/g/dx/dg bookkeeping, overlap, d/overlap
//   gemv_kernel              tiles          t = H * delta_gradient              (reads 8 n^2 B)
//   vec_delta_kernel         1 CTA          delta = alpha*overlap + dg . t
//   update_gemv_kernel       tiles          rank-2 sweep fused with d = H' * g  (r+w 16 n^2 B)
//   identity_kernel          tiles          H = I, only after a gradient-descent step
//
// Summation order is the canonical TREE order of include/dzopt.h (oracle: DZO_ORDER_TREE):
// vectors through 4096 virtual threads + fixed butterfly; GEMV rows as one strictly sequential
// partial per 1024-column chunk, chunks added in ascending order.  No FMA (-fmad=false).
//
// H is the rows x n column-major slab of approximate_inverse_hessian owned by this GPU
// (rows == n on one GPU; rows = n / nranks, global row offset row0, when row-sharded).
#pragma once
#include "common.cuh"

namespace dzo {

// ============================================================================= objectives on a CTA
// Extended Rosenbrock (legacy/ExampleFunctions.jl:10-24 per consecutive pair).  Pair k belongs
// to virtual thread k mod 4096; thread tid emulates virtual threads tid + 1024 q.
struct RosenbrockVec {
    static DZO_DEVINL double term(double x, double y) {
        const double t1 = 1 - x;
        const double t2 = y - x * x;
        return t1 * t1 + 100 * (t2 * t2);
    }
    static DZO_DEVINL double2 grad(double x, double y) {
        const double t1 = 1 - x;
        const double t2 = y - x * x;
        return make_double2(-2 * t1 - 400 * x * t2, 200 * t2);
    }
};

struct ProbeFlags {
    bool changed;   // any(x != w)
    bool same_ref;  // all(w == w_ref)
};

// f(x + alpha*dir) over the canonical tree; optionally compares the trial point with the
// trial point of step alpha_ref (the reference_point of legacy/DZOptimization.jl:136,150).
// MODE 0: plain probe; 1: also compare with the trial point of alpha_ref; 2: evaluate at x itself
template <int MODE>
DZO_DEVINL double cta_probe_rosenbrock(const double* __restrict__ x, const double* __restrict__ dir, long long m,
                                       double alpha, double alpha_ref, double* sm, ProbeFlags& fl) {
    double p[1][4] = {{0.0, 0.0, 0.0, 0.0}};
    int changed = 0, differs = 0;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const double2* d2 = reinterpret_cast<const double2*>(dir);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double acc = 0.0;
        for (long long k = threadIdx.x + 1024 * q; k < m; k += DZO_TREE_WIDTH) {
            const double2 xx = x2[k];
            const double2 dd = d2[k];
            const double w0 = (MODE == 2) ? xx.x : xx.x + alpha * dd.x;  // legacy/Kernels.jl:127-135 axpy!/5 order
            const double w1 = (MODE == 2) ? xx.y : xx.y + alpha * dd.y;
            changed |= (xx.x != w0) | (xx.y != w1);
            if (MODE == 1) {
                const double r0 = xx.x + alpha_ref * dd.x;
                const double r1 = xx.y + alpha_ref * dd.y;
                differs |= (!(w0 == r0)) | (!(w1 == r1));
            }
            acc += RosenbrockVec::term(w0, w1);
        }
        p[0][q] = acc;
    }
    double out[1];
    cta1024_tree_reduce<1>(p, sm, out);
    fl.changed = __syncthreads_or(changed) != 0;
    fl.same_ref = (MODE == 1) ? (__syncthreads_or(differs) == 0) : false;
    return out[0];
}

// only the "did the point move" test of find_three_point_bracket (:73-80, :95-100)
DZO_DEVINL bool cta_point_changed(const double* __restrict__ x, const double* __restrict__ dir, long long n,
                                  double alpha) {
    int changed = 0;
    for (long long e = threadIdx.x; e < n; e += 1024) {
        const double xx = x[e];
        changed |= (xx != xx + alpha * dir[e]);
    }
    return __syncthreads_or(changed) != 0;
}

// quadratic_line_search(functor, f0, t1) [GLUE: :49-172 with first step t1, then :191-216],
// trial point x + (sign*t)*dir, unconstrained extended Rosenbrock.  Uniform control flow: every
// thread of the CTA takes the same branches because every reduced value is broadcast.
DZO_DEVINL void cta_line_search_rosenbrock(const double* __restrict__ x, const double* __restrict__ dir, long long n,
                                           double f0, double t1, double sign, int max_increases, double* sm,
                                           double& t_best, double& f_best, long long& evals) {
    const long long m = n >> 1;
    double x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0;
    ProbeFlags fl;
    do {
        if (!isfinite(f0)) break;                                     // :64-66
        if (!isfinite(t1) || t1 == 0.0) break;                        // [GLUE]
        {                                                             // :71-85 step_is_zero
            int nz = 0;
            for (long long e = threadIdx.x; e < n; e += 1024) nz |= !(dir[e] == 0.0);
            if (!__syncthreads_or(nz)) break;
        }
        double step = t1;
        bool small = false;
        bool changed = cta_point_changed(x, dir, n, sign * step);    // :73-80
        int cap = DZO_LINESEARCH_CAP;
        bool capped = false;
        while (!changed) {                                            // :91-101
            step += step;
            small = true;
            changed = cta_point_changed(x, dir, n, sign * step);
            if (--cap == 0) { capped = true; break; }
        }
        if (capped) break;
        (void)small;  // :107-123: with the null constraint new_point == initial_point cannot hold here
        double fa = cta_probe_rosenbrock<0>(x, dir, m, sign * step, 0.0, sm, fl);  // :126
        ++evals;
        if (fa <= f0) {                                               // :130
            int num_increases = 0;
            cap = DZO_LINESEARCH_CAP;
            for (;;) {                                                // :143-156
                const double ds = step + step;
                num_increases += 1;
                const double fb = cta_probe_rosenbrock<1>(x, dir, m, sign * ds, sign * step, sm, fl);
                ++evals;
                --cap;
                if (((max_increases > 0) && (num_increases >= max_increases)) || !isfinite(fb) || fb > fa ||
                    fl.same_ref || cap == 0) {
                    x1 = step; f1 = fa; x2 = ds; f2 = fb;
                    break;
                }
                step = ds;
                fa = fb;
            }
        } else {                                                      // :157-171
            cap = DZO_LINESEARCH_CAP;
            for (;;) {
                const double hs = 0.5 * step;
                const double fb = cta_probe_rosenbrock<0>(x, dir, m, sign * hs, 0.0, sm, fl);
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
    double xb = 0.0, fb = f0;                                         // :196-202
    if (f1 < fb) { xb = x1; fb = f1; }
    if (f2 < fb) { xb = x2; fb = f2; }
    const double delta_1 = f0 - f1;                                   // :203-205
    const double delta_2 = f2 - f1;
    const double sum_deltas = delta_1 + delta_2;
    if (delta_1 >= 0.0 && delta_2 >= 0.0 && sum_deltas > 0.0) {       // :206-214
        const double twice_delta_1 = delta_1 + delta_1;
        const double delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas);
        const double xq = delta_ratio * x1;
        const double fq = cta_probe_rosenbrock<0>(x, dir, m, sign * xq, 0.0, sm, fl);
        ++evals;
        if (fq < fb) { xb = xq; fb = fq; }
    }
    t_best = xb;
    f_best = fb;
}

// Kernels.dot over the canonical tree (element e -> virtual thread (e/2) mod 4096)
DZO_DEVINL double cta_tree_dot(const double* __restrict__ v, const double* __restrict__ w, long long n, double* sm) {
    double p[1][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double acc = 0.0;
        for (long long k = threadIdx.x + 1024 * q; 2 * k < n; k += DZO_TREE_WIDTH) {
            acc += v[2 * k] * w[2 * k];
            if (2 * k + 1 < n) acc += v[2 * k + 1] * w[2 * k + 1];
        }
        p[0][q] = acc;
    }
    double out[1];
    cta1024_tree_reduce<1>(p, sm, out);
    return out[0];
}

// ============================================================================= step! stage 1
struct LargeVecs {
    double *x, *g, *d, *dx, *dg;  // replicated n-vectors (fields of the optimizer struct)
    double *sd;                   // step_direction / overlap (:874) kept beside d: d itself is
                                  // overwritten by the fused GEMV while other CTAs still read sd
    double *t;                    // _scratch_space = H * delta_gradient (:875)
    LargeCtrl* ctrl;
    long long n;
    // row-sharded mode with fused peer-memory gathers: LOCAL flag words raised by the producers
    const unsigned long long* flags_t;   // t = H*dg slabs of step `calls`
    const unsigned long long* flags_d;   // d = H'*g slabs of step `calls`
    int nranks;
};
// batch > 1 with n > 32 ("run multiple optimizers in parallel" for medium n): problem q's vectors and control block are
// q * n doubles / q blocks further; a kernel instance (a CTA, or a cluster) serves one problem
DZO_DEVINL LargeVecs for_problem(LargeVecs a, long long q) {
    const long long o = q * a.n;
    a.x += o; a.g += o; a.d += o; a.dx += o; a.dg += o; a.sd += o; a.t += o;
    a.ctrl += q;
    return a;
}

// Constructor, legacy/DZOptimization.jl:762-810 (x already holds copy(x0)).
static __global__ void __launch_bounds__(1024, 1) vec_bfgs_init_kernel(LargeVecs a_, double initial_step_length) {
    const LargeVecs a = for_problem(a_, blockIdx.x);
    __shared__ double sm[132];
    const long long n = a.n, m = n >> 1;
    ProbeFlags fl;
    const double f0 = cta_probe_rosenbrock<2>(a.x, a.x, m, 0.0, 0.0, sm, fl);   // :772 (x + 0*x == x)
    for (long long k = threadIdx.x; k < m; k += 1024) {
        const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
        const double2 gg = RosenbrockVec::grad(xx.x, xx.y);                          // :775-776
        reinterpret_cast<double2*>(a.g)[k] = gg;
        reinterpret_cast<double2*>(a.d)[k] = gg;                                     // :784
        reinterpret_cast<double2*>(a.dx)[k] = make_double2(0.0, 0.0);                // :777
        reinterpret_cast<double2*>(a.dg)[k] = make_double2(0.0, 0.0);                // :778
    }
    if (threadIdx.x == 0) {
        LargeCtrl c;
        c.f = f0; c.L = initial_step_length; c.iter = 0; c.type = DZO_STEP_NULL; c.term = 0;
        c.kind = DZO_STEP_NULL; c.pad = 0; c.step_length = 0.0; c.overlap = 0.0; c.delta_norm = 0.0;
        c.evals = 1;
        c.calls = 0;
        for (int i = 0; i < 64; ++i) c.kind_log[i] = 0;
        *a.ctrl = c;
    }
}

// set_state (:819-862): recompute f and g at the restored point; d = H*g follows as a GEMV.
static __global__ void __launch_bounds__(1024, 1) vec_bfgs_restore_kernel(LargeVecs a_) {
    const LargeVecs a = for_problem(a_, blockIdx.x);
    __shared__ double sm[132];
    const long long n = a.n, m = n >> 1;
    ProbeFlags fl;
    const double f0 = cta_probe_rosenbrock<2>(a.x, a.x, m, 0.0, 0.0, sm, fl);   // :828
    for (long long k = threadIdx.x; k < m; k += 1024) {
        const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
        reinterpret_cast<double2*>(a.g)[k] = RosenbrockVec::grad(xx.x, xx.y);        // :830-831
    }
    if (threadIdx.x == 0) {
        a.ctrl->f = f0;
        a.ctrl->term = 0;                                                            // :849
        a.ctrl->kind = DZO_STEP_NULL;
    }
}

// step! :891-960 up to (and including) the O(n) part of update_inverse_hessian! (:873-874).
static __global__ void __launch_bounds__(1024, 1) vec_bfgs_search_kernel(LargeVecs a_) {
    const LargeVecs a = for_problem(a_, blockIdx.x);
    __shared__ double sm[2 * 132];
    __shared__ LargeCtrl sc;
    const long long n = a.n, m = n >> 1;
    if (threadIdx.x == 0) sc = *a.ctrl;
    __syncthreads();
    if (sc.term) {                                                    // :893
        if (threadIdx.x == 0) a.ctrl->kind = DZO_STEP_NULL;
        return;
    }
    const double f0 = sc.f;
    const double step_length = sc.L;                                  // :918
    long long evals = 0;

    // :921, :928  norm(g), norm(d) = sqrt(tree sum of squares)
    double grad_norm, bfgs_norm;
    {
        double p[2][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double ag = 0.0, ad = 0.0;
            for (long long k = threadIdx.x + 1024 * q; k < m; k += DZO_TREE_WIDTH) {
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
                ag += gg.x * gg.x; ag += gg.y * gg.y;
                ad += dd.x * dd.x; ad += dd.y * dd.y;
            }
            p[0][q] = ag; p[1][q] = ad;
        }
        double out[2];
        cta1024_tree_reduce<2>(p, sm, out);
        grad_norm = sqrt(out[0]);
        bfgs_norm = sqrt(out[1]);
    }
    double grad_step_length, grad_obj, bfgs_step_length, bfgs_obj;
    cta_line_search_rosenbrock(a.x, a.g, n, f0, step_length / grad_norm, -1.0, 0, sm, grad_step_length, grad_obj, evals);  // :922-925
    cta_line_search_rosenbrock(a.x, a.d, n, f0, step_length / bfgs_norm, -1.0, 0, sm, bfgs_step_length, bfgs_obj, evals);  // :929-932

    int kind;
    double alpha, fnew, Lnew;
    if (bfgs_obj < f0 && !(bfgs_obj > grad_obj)) {                    // :934
        kind = DZO_STEP_BFGS; alpha = -bfgs_step_length; fnew = bfgs_obj; Lnew = bfgs_step_length * bfgs_norm;  // :937-938
    } else if (grad_obj < f0) {                                       // :962
        kind = DZO_STEP_GRADIENT_DESCENT; alpha = -grad_step_length; fnew = grad_obj; Lnew = grad_step_length * grad_norm;  // :965-966
    } else {
        if (threadIdx.x == 0) {                                       // :989
            a.ctrl->term = 1;
            a.ctrl->kind = DZO_STEP_NULL;
            a.ctrl->evals = sc.evals + evals;
            a.ctrl->kind_log[sc.calls & 63] = DZO_STEP_NULL;
            a.ctrl->calls = sc.calls + 1;
        }
        return;
    }
    const double* dir = (kind == DZO_STEP_BFGS) ? a.d : a.g;

    // :943-950 / :971-978 and the overlap dot (:873); every thread touches only its own pairs
    double p[1][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double acc = 0.0;
        for (long long k = threadIdx.x + 1024 * q; k < m; k += DZO_TREE_WIDTH) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
            const double2 dd = reinterpret_cast<const double2*>(dir)[k];
            double2 xn, dxv, dgv;
            xn.x = xx.x + alpha * dd.x;                               // :945 add!(point, -step, direction)
            xn.y = xx.y + alpha * dd.y;
            const double2 gn = RosenbrockVec::grad(xn.x, xn.y);       // :948
            dxv.x = (-xx.x) + xn.x; dxv.y = (-xx.y) + xn.y;           // :943, :949
            dgv.x = (-gg.x) + gn.x; dgv.y = (-gg.y) + gn.y;           // :944, :950
            reinterpret_cast<double2*>(a.x)[k] = xn;
            reinterpret_cast<double2*>(a.g)[k] = gn;
            reinterpret_cast<double2*>(a.dx)[k] = dxv;
            reinterpret_cast<double2*>(a.dg)[k] = dgv;
            if (kind == DZO_STEP_BFGS) {
                acc += dd.x * dgv.x;                                  // :873 dot(step_direction, delta_gradient)
                acc += dd.y * dgv.y;
            } else {
                reinterpret_cast<double2*>(a.d)[k] = gn;              // :984-986 direction = copy of the new gradient
            }
        }
        p[0][q] = acc;
    }
    double overlap = 0.0;
    if (kind == DZO_STEP_BFGS) {
        double out[1];
        cta1024_tree_reduce<1>(p, sm, out);
        overlap = out[0];
        const double inv_overlap = 1.0 / overlap;                     // :874 inv(overlap)
        for (long long k = threadIdx.x; k < m; k += 1024) {
            const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
            reinterpret_cast<double2*>(a.sd)[k] = make_double2(dd.x * inv_overlap, dd.y * inv_overlap);
        }
    }
    if (threadIdx.x == 0) {
        LargeCtrl c = sc;
        c.f = fnew; c.L = Lnew; c.type = kind; c.iter = sc.iter + 1;  // :937-940 / :965-968
        c.kind = kind; c.step_length = alpha; c.overlap = overlap; c.delta_norm = 0.0;
        c.evals = sc.evals + evals;
        c.kind_log[sc.calls & 63] = (unsigned char)kind;
        c.calls = sc.calls + 1;
        *a.ctrl = c;
    }
}

// :876  delta_norm = step_length*overlap + dot(delta_gradient, scratch)
static __global__ void __launch_bounds__(1024, 1) vec_delta_kernel(LargeVecs a_) {
    const LargeVecs a = for_problem(a_, blockIdx.x);
    __shared__ double sm[132];
    if (a.ctrl->kind != DZO_STEP_BFGS) return;
    const double s = cta_tree_dot(a.dg, a.t, a.n, sm);
    if (threadIdx.x == 0) a.ctrl->delta_norm = a.ctrl->step_length * a.ctrl->overlap + s;
}

// overlap + in-place rescale for the kernel-level entry dzo_dev_update_inverse_hessian (:873-874)
static __global__ void __launch_bounds__(1024, 1) vec_overlap_scale_kernel(LargeVecs a, double step_length) {
    __shared__ double sm[132];
    const double overlap = cta_tree_dot(a.d, a.dg, a.n, sm);
    const double inv_overlap = 1.0 / overlap;
    for (long long e = threadIdx.x; e < a.n; e += 1024) a.sd[e] = a.d[e] * inv_overlap;
    if (threadIdx.x == 0) {
        a.ctrl->kind = DZO_STEP_BFGS;
        a.ctrl->step_length = step_length;
        a.ctrl->overlap = overlap;
    }
}

static __global__ void __launch_bounds__(1024, 1) vec_dot_kernel(const double* v, const double* w, long long n, double* out) {
    __shared__ double sm[132];
    const double s = cta_tree_dot(v, w, n, sm);
    if (threadIdx.x == 0) *out = s;
}

// batch of extended-Rosenbrock objective / gradient evaluations in TREE order (dzo_dev_objective)
static __global__ void __launch_bounds__(1024, 1) vec_rosenbrock_objective_kernel(const double* x, long long n, double* f) {
    __shared__ double sm[132];
    ProbeFlags fl;
    const double* xp = x + (long long)blockIdx.x * n;
    const double v = cta_probe_rosenbrock<2>(xp, xp, n >> 1, 0.0, 0.0, sm, fl);
    if (threadIdx.x == 0) f[blockIdx.x] = v;
}
static __global__ void vec_rosenbrock_gradient_kernel(const double* x, long long total_pairs, double* g) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total_pairs) return;
    const double2 xx = reinterpret_cast<const double2*>(x)[k];
    reinterpret_cast<double2*>(g)[k] = RosenbrockVec::grad(xx.x, xx.y);
}

// ============================================================================= n^2 sweeps
// cache operators of the update sweep's streaming H accesses (compile-time A/B knobs of tools/build_variant.sh; measured at
// n = 16384, fraction of the copy peak: loads __ldcs 0.943, __ldcg 0.949, __ldlu 0.948, plain 0.930 (with plain stores);
// stores __stcs 0.943, __stcg 0.942, __stwt 0.943, plain 0.919 -- profiles/r02_update_sweep_cache_operators.jsonl)
#ifndef DZO_SWEEP_LD
#define DZO_SWEEP_LD 1      // 0 __ldcs, 1 __ldcg, 2 plain, 3 __ldlu
#endif
#ifndef DZO_SWEEP_ST
#define DZO_SWEEP_ST 0      // 0 __stcs, 1 plain, 2 __stcg, 3 __stwt
#endif
DZO_DEVINL double2 sweep_ld(const double2* p) {
#if DZO_SWEEP_LD == 1
    return __ldcg(p);
#elif DZO_SWEEP_LD == 2
    return *p;
#elif DZO_SWEEP_LD == 3
    return __ldlu(p);
#else
    return __ldcs(p);
#endif
}
DZO_DEVINL void sweep_st(double2* p, double2 v) {
#if DZO_SWEEP_ST == 1
    *p = v;
#elif DZO_SWEEP_ST == 2
    __stcg(p, v);
#elif DZO_SWEEP_ST == 3
    __stwt(p, v);
#else
    __stcs(p, v);
#endif
}
// A sweep CTA has blockDim.x in {32, 64, 128, 256} threads, each owning two adjacent rows: the host picks
// the size so that even a thin row slab (row-sharded mode) yields enough tiles to fill the GPU.
constexpr int kSweepMaxThreads = 256;
constexpr int kSweepMinRows = 64;                  // rows per tile of the smallest CTA (counter arrays are sized for it)
constexpr int kSweepThreads = 128;                 // default (n x n on one GPU)
constexpr int kSweepRows = 2 * kSweepThreads;
constexpr int kSweepUnroll = 8;                    // columns in flight per thread (default; tunable 4 / 8 / 16)

struct SweepArgs {
    double* H;              // rows x n slab, column-major, leading dimension ld
    long long ld, rows, n;
    long long row0;         // global index of local row 0 (row-sharded mode)
    const double* v;        // GEMV operand (delta_gradient for t, gradient for d)      [n]
    const double* s;        // step_direction / overlap                                  [n]
    const double* t;        // H * delta_gradient                                        [n]
    double* partial;        // nchunks x rows chunk partials (unused when nchunks == 1)
    double* out;            // GEMV result; local row i is written to out[row0 + i]
    unsigned* counters;     // one per row block, zero between launches
    const LargeCtrl* ctrl;  // device-side predicate (may be null = always run)
    int need_kind;
    int nchunks;
    PeerSet peers;          // fused all-gather over peer memory (nranks == 1: plain local store)
};
// batch > 1: blockIdx.z selects the problem (its n x n matrix, vectors, chunk partials, tile counters, control block)
DZO_DEVINL SweepArgs sweep_for_problem(SweepArgs a) {
    const long long q = blockIdx.z;
    if (q == 0) return a;
    a.H += q * a.ld * a.n;
    if (a.v) a.v += q * a.n;
    if (a.s) a.s += q * a.n;
    if (a.t) a.t += q * a.n;
    if (a.out) a.out += q * a.n;
    if (a.partial) a.partial += q * (long long)a.nchunks * a.rows;
    a.counters += q * gridDim.x;
    if (a.ctrl) a.ctrl += q;
    return a;
}

// store one finished row: locally and, when sharded with the fused gather, into every peer's copy
DZO_DEVINL void sweep_store_row(const SweepArgs& a, long long gi, double r) {
    a.out[gi] = r;
    if (a.peers.nranks > 1) {
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < a.peers.nranks && p != a.peers.rank) a.peers.out[p][gi] = r;
    }
}
// called by the CTA that finished a row block, after its stores: the LAST row block of the launch raises
// flag[rank] = seq in every peer (release at system scope; the fences make every CTA's row stores visible first)
DZO_DEVINL void sweep_signal_peers(const SweepArgs& a) {
    if (a.peers.nranks <= 1) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(a.peers.done, 1u);
        if (prev == gridDim.x - 1) {
            *a.peers.done = 0;
            __threadfence_system();
            const unsigned long long seq = (unsigned long long)a.ctrl->calls;
            for (int p = 0; p < a.peers.nranks; ++p) st_release_sys(a.peers.flags[p] + a.peers.rank, seq);
        }
    }
}

// Chunk partials -> out, ascending chunk order starting from partial 0 (oracle gemv_rows_).
// Executed by the LAST tile of a row block to finish (threadfence + counter), so the result
// does not depend on which tile that is.
DZO_DEVINL void sweep_finish_rows(const SweepArgs& a, long long i0, double acc0, double acc1, bool two, bool one) {
    const int c = blockIdx.y;
    if (a.nchunks == 1) {
        if (one) sweep_store_row(a, a.row0 + i0, acc0);
        if (two) sweep_store_row(a, a.row0 + i0 + 1, acc1);
        sweep_signal_peers(a);
        return;
    }
    if (one) a.partial[(long long)c * a.rows + i0] = acc0;
    if (two) a.partial[(long long)c * a.rows + i0 + 1] = acc1;
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&a.counters[blockIdx.x], 1u);
        s_last = (prev == (unsigned)(a.nchunks - 1));
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int k = 0; k < 2; ++k) {
        const long long i = i0 + k;
        if (k == 0 ? !one : !two) continue;
        const volatile double* pp = a.partial + i;
        double r = pp[0];
        for (int cc = 1; cc < a.nchunks; ++cc) r += pp[(long long)cc * a.rows];
        sweep_store_row(a, a.row0 + i, r);
    }
    if (threadIdx.x == 0) a.counters[blockIdx.x] = 0;  // ready for the next launch
    sweep_signal_peers(a);
}

// mul!(out, H, v)  legacy/DZOptimization.jl:875, :958-960.  Thread-per-row walk over one
// 1024-column chunk: a warp reads 512 contiguous bytes of every column (H is column-major),
// and each thread's accumulator is exactly the sequential chunk partial of the oracle.
template <int U>
static __global__ void __launch_bounds__(kSweepMaxThreads) gemv_kernel(SweepArgs a_) {
    const SweepArgs a = sweep_for_problem(a_);
    if (a.ctrl && a.ctrl->kind != a.need_kind) return;
    __shared__ double sv[DZO_GEMV_CHUNK];
    const long long c0 = (long long)blockIdx.y * DZO_GEMV_CHUNK;
    const int nc = (int)((a.n - c0 < DZO_GEMV_CHUNK) ? (a.n - c0) : DZO_GEMV_CHUNK);
    for (int j = threadIdx.x; j < nc; j += blockDim.x) sv[j] = a.v[c0 + j];
    __syncthreads();
    const long long i0 = (long long)blockIdx.x * (2 * blockDim.x) + 2 * threadIdx.x;
    const bool one = i0 < a.rows, two = i0 + 1 < a.rows;
    double acc0 = 0.0, acc1 = 0.0;
    const double* base = a.H + i0 + c0 * a.ld;
    if (two && ((a.ld & 1) == 0)) {
        int j = 0;
        for (; j + U <= nc; j += U) {
            double2 h[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                h[u] = __ldcs(reinterpret_cast<const double2*>(base + (long long)(j + u) * a.ld));
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double vj = sv[j + u];
                acc0 += h[u].x * vj;
                acc1 += h[u].y * vj;
            }
        }
        for (; j < nc; ++j) {
            const double2 h = __ldcs(reinterpret_cast<const double2*>(base + (long long)j * a.ld));
            acc0 += h.x * sv[j];
            acc1 += h.y * sv[j];
        }
    } else if (one) {
        for (int j = 0; j < nc; ++j) {
            const double vj = sv[j];
            acc0 += base[(long long)j * a.ld] * vj;
            if (two) acc1 += base[(long long)j * a.ld + 1] * vj;
        }
    }
    sweep_finish_rows(a, i0, acc0, acc1, two, one);
}

// update_inverse_hessian! rank-2 sweep (legacy/DZOptimization.jl:878-886, exact operation
// order of :882-884) fused with next_step_direction = H' * gradient (:958-960): every element
// of H is read once, updated, written once, and contributes to the new direction on the way.
DZO_DEVINL void identity_tile(const SweepArgs& a);

template <int U>
static __global__ void __launch_bounds__(kSweepMaxThreads) update_gemv_kernel(SweepArgs a_) {
    const SweepArgs a = sweep_for_problem(a_);
    if (a.need_kind == DZO_STEP_GRADIENT_DESCENT + 100 && a.ctrl->kind == DZO_STEP_GRADIENT_DESCENT) {
        identity_tile(a);   // :981 -- step! resets H after a gradient-descent step; same launch, same tiling
        return;
    }
    if (a.ctrl->kind != DZO_STEP_BFGS) return;
    __shared__ double ss[DZO_GEMV_CHUNK], st[DZO_GEMV_CHUNK], sg[DZO_GEMV_CHUNK];
    const double delta = a.ctrl->delta_norm;
    const long long c0 = (long long)blockIdx.y * DZO_GEMV_CHUNK;
    const int nc = (int)((a.n - c0 < DZO_GEMV_CHUNK) ? (a.n - c0) : DZO_GEMV_CHUNK);
    for (int j = threadIdx.x; j < nc; j += blockDim.x) {
        ss[j] = a.s[c0 + j];
        st[j] = a.t[c0 + j];
        sg[j] = a.v ? a.v[c0 + j] : 0.0;
    }
    __syncthreads();
    const long long i0 = (long long)blockIdx.x * (2 * blockDim.x) + 2 * threadIdx.x;
    const bool one = i0 < a.rows, two = i0 + 1 < a.rows;
    double acc0 = 0.0, acc1 = 0.0;
    double* base = a.H + i0 + c0 * a.ld;
    const double si0 = one ? a.s[a.row0 + i0] : 0.0, ti0 = one ? a.t[a.row0 + i0] : 0.0;
    const double si1 = two ? a.s[a.row0 + i0 + 1] : 0.0, ti1 = two ? a.t[a.row0 + i0 + 1] : 0.0;
    if (two && ((a.ld & 1) == 0)) {
        int j = 0;
        for (; j + U <= nc; j += U) {
            double2 h[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                h[u] = sweep_ld(reinterpret_cast<const double2*>(base + (long long)(j + u) * a.ld));
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double sj = ss[j + u], tj = st[j + u], gj = sg[j + u];
                h[u].x = h[u].x + (delta * (si0 * sj) - (ti0 * sj + si0 * tj));   // :882-884
                h[u].y = h[u].y + (delta * (si1 * sj) - (ti1 * sj + si1 * tj));
                sweep_st(reinterpret_cast<double2*>(base + (long long)(j + u) * a.ld), h[u]);
                acc0 += h[u].x * gj;                                               // :958-960
                acc1 += h[u].y * gj;
            }
        }
        for (; j < nc; ++j) {
            double2 h = __ldcs(reinterpret_cast<const double2*>(base + (long long)j * a.ld));
            const double sj = ss[j], tj = st[j], gj = sg[j];
            h.x = h.x + (delta * (si0 * sj) - (ti0 * sj + si0 * tj));
            h.y = h.y + (delta * (si1 * sj) - (ti1 * sj + si1 * tj));
            __stcs(reinterpret_cast<double2*>(base + (long long)j * a.ld), h);
            acc0 += h.x * gj;
            acc1 += h.y * gj;
        }
    } else if (one) {
        for (int j = 0; j < nc; ++j) {
            const double sj = ss[j], tj = st[j], gj = sg[j];
            double h0 = base[(long long)j * a.ld];
            h0 = h0 + (delta * (si0 * sj) - (ti0 * sj + si0 * tj));
            base[(long long)j * a.ld] = h0;
            acc0 += h0 * gj;
            if (two) {
                double h1 = base[(long long)j * a.ld + 1];
                h1 = h1 + (delta * (si1 * sj) - (ti1 * sj + si1 * tj));
                base[(long long)j * a.ld + 1] = h1;
                acc1 += h1 * gj;
            }
        }
    }
    if (a.out) sweep_finish_rows(a, i0, acc0, acc1, two, one);
}

// identity_matrix!  legacy/DZOptimization.jl:712-720 (after a gradient-descent step, :981).
// Same tiling as the sweeps; pure streaming stores.
DZO_DEVINL void identity_tile(const SweepArgs& a) {
    const long long c0 = (long long)blockIdx.y * DZO_GEMV_CHUNK;
    const int nc = (int)((a.n - c0 < DZO_GEMV_CHUNK) ? (a.n - c0) : DZO_GEMV_CHUNK);
    const long long i0 = (long long)blockIdx.x * (2 * blockDim.x) + 2 * threadIdx.x;
    const bool one = i0 < a.rows, two = i0 + 1 < a.rows;
    double* base = a.H + i0 + c0 * a.ld;
    const long long gi = a.row0 + i0;  // global row of the first of my two rows
    if (two && ((a.ld & 1) == 0)) {
#pragma unroll 8
        for (int j = 0; j < nc; ++j) {
            const long long gj = c0 + j;
            __stcs(reinterpret_cast<double2*>(base + (long long)j * a.ld),
                   make_double2(gj == gi ? 1.0 : 0.0, gj == gi + 1 ? 1.0 : 0.0));
        }
    } else if (one) {
        for (int j = 0; j < nc; ++j) {
            const long long gj = c0 + j;
            base[(long long)j * a.ld] = (gj == gi) ? 1.0 : 0.0;
            if (two) base[(long long)j * a.ld + 1] = (gj == gi + 1) ? 1.0 : 0.0;
        }
    }
}
static __global__ void __launch_bounds__(kSweepMaxThreads) identity_kernel(SweepArgs a_) {
    const SweepArgs a = sweep_for_problem(a_);
    if (a.ctrl && a.ctrl->kind != a.need_kind) return;
    identity_tile(a);
}

}  // namespace dzo
