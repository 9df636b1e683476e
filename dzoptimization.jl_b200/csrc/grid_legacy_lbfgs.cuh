// grid_legacy_lbfgs.cuh -- the LEGACY LBFGSOptimizer step! (legacy/DZOptimization.jl:565-695, decorators :222-296)
// for n > DZO_TREE_BLOCK on the whole GPU: the cooperative grid / blocked canonical tree machinery of grid_lbfgs.cuh
// (eight 512-thread CTAs per 65536-element block, one grid barrier per reduction) under the control flow of
// cluster_legacy_lbfgs_kernel (legacy_lbfgs.cuh).  Every pass over the direction of the two-loop correction also
// accumulates the dot product the next stage needs.
// a.algo == 1 runs step!(::GradientDescentOptimizer) (legacy/DZOptimization.jl:393-449) on the same machinery: the same
// constructor and line search, no gradient retry, no history, next_step_direction = -(step length / |g|) * g.
#pragma once
#include "grid_lbfgs.cuh"
#include "legacy_lbfgs.cuh"

namespace dzo {

struct GridLegacyArgs {
    LegacyArgs a;                            // vectors, control block, decorators (legacy_lbfgs.cuh)
    double* part;                            // [2][kGridQ][kGridMaxParts]
    unsigned* fpart;                         // [2][kGridMaxParts]
    int nblocks;
    int stage;                               // 1: kGridStageBytes of dynamic shared memory are there (OWN == 1 launches)
};

// bit 0: any(x != x + alpha*dir)   bit 1: any(dir != 0)
// (out of line: the line search calls these from a dozen sites and each expands to an OWN-way unrolled pass)
template <int OWN>
static __device__ __noinline__ unsigned grid_point_flags(GridCtx& c, long long m2, const double* __restrict__ x, const double* __restrict__ dir,
                                     double alpha) {
    constexpr int kGridOwn = OWN;
    double acc[1][kGridOwn];
    unsigned fl[kGridOwn];
#pragma unroll
    for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
    DZO_GRID_OWN_PAIRS(c, m2, j, k) {
        const double2 xx = reinterpret_cast<const double2*>(x)[k];
        const double2 dd = reinterpret_cast<const double2*>(dir)[k];
        if ((xx.x != xx.x + alpha * dd.x) | (xx.y != xx.y + alpha * dd.y)) fl[j] |= 1u;
        if ((!(dd.x == 0.0)) | (!(dd.y == 0.0))) fl[j] |= 2u;
    }
    double out[1];
    unsigned f;
    grid_reduce<1, kGridOwn>(c, acc, fl, out, f);
    return f;
}

// lse(alpha), see legacy_probe (legacy_lbfgs.cuh) for the flag bits
template <int MODE, int OWN>
static __device__ __noinline__ double grid_legacy_probe(GridCtx& c, long long m2, const LegacyDecor& D, const double* __restrict__ x,
                                    const double* __restrict__ dir, double alpha, double alpha_ref, unsigned& flags) {
    constexpr int kGridOwn = OWN;
    double acc[2][kGridOwn];
    unsigned fl[kGridOwn];
#pragma unroll
    for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; fl[j] = 0; }
    DZO_GRID_OWN_PAIRS(c, m2, j, k) {
        const double2 xx = reinterpret_cast<const double2*>(x)[k];
        double w0 = xx.x, w1 = xx.y;
        if (MODE != 2) {
            const double2 dd = reinterpret_cast<const double2*>(dir)[k];
            w0 = xx.x + alpha * dd.x;
            w1 = xx.y + alpha * dd.y;
            if ((xx.x != w0) | (xx.y != w1)) fl[j] |= 1u;
            w0 = D.clamp(w0);
            w1 = D.clamp(w1);
            if ((!(xx.x == w0)) | (!(xx.y == w1))) fl[j] |= 4u;
            if (MODE == 1) {
                const double r0 = D.clamp(xx.x + alpha_ref * dd.x);
                const double r1 = D.clamp(xx.y + alpha_ref * dd.y);
                if ((!(w0 == r0)) | (!(w1 == r1))) fl[j] |= 2u;
            }
        }
        acc[0][j] += RosenbrockVec::term(w0, w1);
        if (D.l2) { acc[1][j] += w0 * w0; acc[1][j] += w1 * w1; }
    }
    double out[2];
    grid_reduce<2, kGridOwn>(c, acc, fl, out, flags);
    return D.l2 ? out[0] + D.lam * out[1] : out[0];                        // :233-234
}

// QuadraticLineSearch(max_increases)(lse, f0, _)  :191-216 with find_three_point_bracket :49-172 (first step 1);
// the same control flow as legacy_line_search (legacy_lbfgs.cuh)
template <int OWN>
DZO_DEVINL void grid_legacy_line_search(GridCtx& c, long long m2, const LegacyDecor& D, const double* __restrict__ x,
                                        const double* __restrict__ dir, double f0, int max_increases, double& t_best,
                                        double& f_best, long long& evals) {
    double x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0;
    unsigned pf;
    do {
        if (!isfinite(f0)) break;                                         // :64-66
        double step = 1.0;
        unsigned fl = grid_point_flags<OWN>(c, m2, x, dir, step);
        if (!(fl & 2u)) break;                                            // :71-85 step_is_zero
        int cap = DZO_LINESEARCH_CAP;
        bool capped = false, small = false;
        while (!(fl & 1u)) {                                              // :91-101
            step += step;
            small = true;
            fl = grid_point_flags<OWN>(c, m2, x, dir, step);
            if (--cap == 0) { capped = true; break; }
        }
        if (capped) break;
        double fa = grid_legacy_probe<0, OWN>(c, m2, D, x, dir, step, 0.0, pf);   // :104, :126
        if (small && !(pf & 4u)) break;                                   // :107-123
        ++evals;
        if (fa <= f0) {                                                   // :130
            int num_increases = 0;
            cap = DZO_LINESEARCH_CAP;
            for (;;) {                                                    // :143-156
                const double ds = step + step;
                num_increases += 1;
                const double fb = grid_legacy_probe<1, OWN>(c, m2, D, x, dir, ds, step, pf);
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
                const double fb = grid_legacy_probe<0, OWN>(c, m2, D, x, dir, hs, 0.0, pf);
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
        const double fq = grid_legacy_probe<0, OWN>(c, m2, D, x, dir, xq, 0.0, pf);
        ++evals;
        if (fq < fb) { xb = xq; fb = fq; }
    }
    t_best = xb;
    f_best = fb;
}

template <int OWN>   // eighths one CTA may own (1: n <= CTAs * 8192); compiled in ONE translation unit (grid_legacy_tu.cu)
static __global__ void __launch_bounds__(kClusterThreads, 1) grid_legacy_lbfgs_kernel(GridLegacyArgs ga) {
    extern __shared__ __align__(16) unsigned char grid_stage_raw[];
    constexpr int kGridOwn = OWN;
    const LegacyArgs& a = ga.a;
    __shared__ LegacyCtrl sc;
    __shared__ double s_warp[kGridQ * 16];
    __shared__ unsigned s_wflag[16];
    __shared__ double s_out[kGridQ];
    __shared__ unsigned s_flags;
    GridCtx c{cg::this_grid(), (int)gridDim.x, (int)blockIdx.x, ga.nblocks, 0, ga.part, ga.fpart, s_warp, s_wflag, s_out, &s_flags};
    grid_ctx_begin(c);
    const long long n = a.n, m2 = n >> 1;
    const bool leader = (blockIdx.x == 0 && threadIdx.x == 0);
    LegacyDecor D;
    D.l2 = (a.decor & DZO_DECOR_L2) != 0; D.box = (a.decor & DZO_DECOR_BOX) != 0;
    D.lam = a.l2; D.lo = a.lo; D.hi = a.hi;
    if (threadIdx.x == 0 && a.mode == 0) sc = *a.ctrl;
    __syncthreads();
    c.grid.sync();                 // every CTA holds the control block before the leader may rewrite it

    if (a.mode == 1) {
        // LBFGSOptimizer(c!, f, g!, linesearch, x0, L0, m)  :489-548 (x already holds collect(x0))
        double acc[3][kGridOwn];
        unsigned fl[kGridOwn];
#pragma unroll
        for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; acc[2][j] = 0.0; fl[j] = 0; }
        DZO_GRID_OWN_PAIRS(c, m2, j, k) {
            double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            xx.x = D.clamp(xx.x); xx.y = D.clamp(xx.y);                                         // :500
            const double2 gg = D.grad(xx.x, xx.y);                                              // :507-508
            reinterpret_cast<double2*>(a.x)[k] = xx;
            reinterpret_cast<double2*>(a.g)[k] = gg;
            reinterpret_cast<double2*>(a.dx)[k] = make_double2(0.0, 0.0);                       // :501
            reinterpret_cast<double2*>(a.dg)[k] = make_double2(0.0, 0.0);                       // :509
            acc[0][j] += RosenbrockVec::term(xx.x, xx.y);                                       // :503
            if (D.l2) { acc[1][j] += xx.x * xx.x; acc[1][j] += xx.y * xx.y; }
            acc[2][j] += gg.x * gg.x; acc[2][j] += gg.y * gg.y;
        }
        double out[3];
        unsigned f;
        grid_reduce<3, kGridOwn>(c, acc, fl, out, f);
        const double f0 = D.l2 ? out[0] + D.lam * out[1] : out[0];
        const double inv_gradient_norm = 1.0 / sqrt(out[2]);                                    // :512
        const bool ok = isfinite(inv_gradient_norm);
        const double cc = -a.initial_step_length * inv_gradient_norm;
        DZO_GRID_OWN_PAIRS(c, m2, j, k) {                                                       // :513-517
            const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
            reinterpret_cast<double2*>(a.d)[k] = ok ? make_double2(gg.x * cc, gg.y * cc) : make_double2(0.0, 0.0);
        }
        if (leader) {
            LegacyCtrl t;
            t.f = f0; t.df = 0.0; t.L = 0.0; t.iter = 0; t.hist_count = 0; t.evals = 1;
            t.term = (!isfinite(f0)) || (!ok);                                                  // :524-526
            for (int i = 0; i < DZO_LBFGS_MAX_HISTORY; ++i) { t.rho[i] = 0.0; t.alpha[i] = 0.0; }
            *a.ctrl = t;
        }
        grid_ctx_end(c);
        return;
    }

    // step!  :565-695, k times; sc is the CTA-local copy of the control block (identical on every CTA)
    const int m = a.m;
    for (int step_i = 0; step_i < a.ksteps; ++step_i) {
        if (sc.term) break;                                                                     // :578
        const double f0 = sc.f;
        long long evals = 0;
        double step_size, objective_value;
        grid_legacy_line_search<OWN>(c, m2, D, a.x, a.d, f0, a.max_increases, step_size, objective_value, evals);   // :584-586
        bool reset_history = false;
        if (a.algo == 1 && (step_size == 0.0 || !(objective_value < f0))) {                     // GD :410-414
            if (threadIdx.x == 0) { sc.term = 1; sc.evals += evals; }
            __syncthreads();
            break;
        }
        if (step_size == 0.0 || !(objective_value < f0)) {                                      // :589-590
            double acc[1][kGridOwn];
            unsigned fl[kGridOwn];
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
            DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                acc[0][j] += gg.x * gg.x; acc[0][j] += gg.y * gg.y;
            }
            double out[1];
            unsigned f;
            grid_reduce<1, kGridOwn>(c, acc, fl, out, f);
            const double cc = -sc.L * (1.0 / sqrt(out[0]));                                     // :594-595
            DZO_GRID_OWN_PAIRS(c, m2, j, k) {                                                   // :593
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                reinterpret_cast<double2*>(a.d)[k] = make_double2(gg.x * cc, gg.y * cc);
            }
            grid_legacy_line_search<OWN>(c, m2, D, a.x, a.d, f0, a.max_increases, step_size, objective_value, evals);   // :596-598
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
        double q4[4];
        {
            double acc[4][kGridOwn];   // norm2(dx), norm2(g), dot(dx, dg), norm2(dg)
            unsigned fl[kGridOwn];
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; acc[2][j] = 0.0; acc[3][j] = 0.0; fl[j] = 0; }
            DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
                const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
                const double2 go = reinterpret_cast<const double2*>(a.g)[k];
                double2 xn, dxv, dgv;
                xn.x = D.clamp(xx.x + step_size * dd.x);                                        // :615-616
                xn.y = D.clamp(xx.y + step_size * dd.y);
                dxv.x = xn.x - xx.x; dxv.y = xn.y - xx.y;                                       // :614, :619 delta!
                const double2 gn = D.grad(xn.x, xn.y);                                          // :630
                dgv.x = gn.x - go.x; dgv.y = gn.y - go.y;                                       // :629, :631
                reinterpret_cast<double2*>(a.x)[k] = xn;
                reinterpret_cast<double2*>(a.dx)[k] = dxv;
                reinterpret_cast<double2*>(a.g)[k] = gn;
                reinterpret_cast<double2*>(a.dg)[k] = dgv;
                if (a.algo == 0) {
                    reinterpret_cast<double2*>(Snew)[k] = dxv;                                  // :642-643 (unused once terminated)
                    reinterpret_cast<double2*>(Ynew)[k] = dgv;
                }
                acc[0][j] += dxv.x * dxv.x; acc[0][j] += dxv.y * dxv.y;
                acc[1][j] += gn.x * gn.x;   acc[1][j] += gn.y * gn.y;
                acc[2][j] += dxv.x * dgv.x; acc[2][j] += dxv.y * dgv.y;
                acc[3][j] += dgv.x * dgv.x; acc[3][j] += dgv.y * dgv.y;
            }
            unsigned f;
            grid_reduce<4, kGridOwn>(c, acc, fl, q4, f);
        }
        const double step_length = sqrt(q4[0]);                                                 // :620-621
        const double inv_gradient_norm = 1.0 / sqrt(q4[1]);                                     // :634
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
        if (a.algo == 1) {                                                                      // GD :445-446
            const double cc = -step_length * inv_gradient_norm;
            DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                reinterpret_cast<double2*>(a.d)[k] = make_double2(cc * gg.x, cc * gg.y);
            }
            __syncthreads();
            continue;
        }
        const double delta_overlap = q4[2];                                                     // :646
        const double rho_new = 1.0 / delta_overlap;                                             // :647
        const int hist_prev = reset_history ? 0 : sc.hist_count;
        const int hist_count = (hist_prev + 1 < m) ? hist_prev + 1 : m;                         // :650
        const long long hist_end = iter, hist_begin = hist_end - hist_count + 1;                // :651-652
        __syncthreads();                       // everyone has read sc.hist_count / sc.rho before thread 0 rewrites them
        if (threadIdx.x == 0) { sc.rho[cnew] = rho_new; sc.hist_count = hist_count; }
        __syncthreads();
        // two-loop correction as written (:656-680); every pass also accumulates the dot product of the NEXT stage.
        // Stage list: first loop it = hist_end .. hist_begin (dot with S_c, then d += alpha*Y_c), scale by gamma,
        // second loop it = hist_begin .. hist_end (dot with Y_c, then d += beta*S_c), negate, dot with g.
        double acc[1][kGridOwn];
        unsigned fl[kGridOwn];
        double out[1];
        unsigned f;
#pragma unroll
        for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
        DirRegs<OWN> Dr;                       // OWN == 1: the direction lives in registers across the 2m + 1 passes
        PassStage<OWN> St{(OWN == 1 && ga.stage) ? reinterpret_cast<double2*>(grid_stage_raw) : nullptr, false};
        // the two vectors a pass reads depend only on loop indices: the next pass's are fetched into shared memory while
        // the reduction of the current one is in flight (grid_lbfgs.cuh: PassStage)
        auto first_loop_vectors = [&](long long it, const double*& y, const double*& nxt) {
            const int cc = (int)((it - 1) % m);
            const bool last = (it == hist_begin);
            const int cn = last ? (int)((hist_begin - 1) % m) : (int)((it - 2) % m);
            y = a.Y + (long long)cc * n;
            nxt = last ? (a.Y + (long long)cn * n) : (a.S + (long long)cn * n);
        };
        auto second_loop_vectors = [&](long long it, const double*& sp, const double*& nxt) {
            const int cc = (int)((it - 1) % m);
            sp = a.S + (long long)cc * n;
            nxt = (it == hist_end) ? a.g : (a.Y + (long long)((it) % m) * n);                    // Y of it+1, or g for :683-684
        };
        {
            const double* s0 = a.S + (long long)((hist_end - 1) % m) * n;
            own_pairs_idx<OWN>(c, m2, [&](int j, int q, long long k) {                          // :656 d = g, and d . S_c of the first stage
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                const double2 ss = reinterpret_cast<const double2*>(s0)[k];
                Dr.set(a.d, q, k, gg, false);
                acc[0][j] += gg.x * ss.x; acc[0][j] += gg.y * ss.y;
            });
        }
        grid_reduce_then<1, kGridOwn>(c, acc, fl, out, f, [&] {
            const double *fa, *fb;
            first_loop_vectors(hist_end, fa, fb);
            St.fetch(c, m2, fa, fb);
        });
        const double gamma = delta_overlap / q4[3];                                             // :669-670
        for (long long it = hist_end; it >= hist_begin; --it) {                                 // :659-666
            const int cc = (int)((it - 1) % m);
            const double alpha = sc.rho[cc] * out[0];
            if (threadIdx.x == 0) sc.alpha[cc] = alpha;
            const double *y, *nxt;
            first_loop_vectors(it, y, nxt);
            const bool last = (it == hist_begin);
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) acc[0][j] = 0.0;
            St.wait();
            own_pairs_idx<OWN>(c, m2, [&](int j, int q, long long k) {
                double2 dd = Dr.get(a.d, q, k);
                const double2 yy = St.get(0, q, y, k);
                const double2 nn = St.get(1, q, nxt, k);
                dd.x += alpha * yy.x; dd.y += alpha * yy.y;
                if (last) { dd.x *= gamma; dd.y *= gamma; }                                     // :669-670 scale!
                Dr.set(a.d, q, k, dd, false);
                acc[0][j] += dd.x * nn.x; acc[0][j] += dd.y * nn.y;
            });
            St.done();
            grid_reduce_then<1, kGridOwn>(c, acc, fl, out, f, [&] {
                const double *fa, *fb;
                if (!last) first_loop_vectors(it - 1, fa, fb);
                else second_loop_vectors(hist_begin, fa, fb);
                St.fetch(c, m2, fa, fb);
            });
        }
        __syncthreads();                       // sc.alpha[] visible
        double gradient_overlap = 0.0;
        for (long long it = hist_begin; it <= hist_end; ++it) {                                 // :673-680
            const int cc = (int)((it - 1) % m);
            const double beta = sc.alpha[cc] - sc.rho[cc] * out[0];
            const double *sp, *nxt;
            second_loop_vectors(it, sp, nxt);
            const bool last = (it == hist_end);
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) acc[0][j] = 0.0;
            St.wait();
            own_pairs_idx<OWN>(c, m2, [&](int j, int q, long long k) {
                double2 dd = Dr.get(a.d, q, k);
                const double2 ss = St.get(0, q, sp, k);
                const double2 nn = St.get(1, q, nxt, k);
                dd.x += beta * ss.x; dd.y += beta * ss.y;
                if (last) { dd.x = -dd.x; dd.y = -dd.y; }                                       // :683 negate!
                Dr.set(a.d, q, k, dd, last);                                                    // the finished direction goes to memory
                acc[0][j] += dd.x * nn.x; acc[0][j] += dd.y * nn.y;
            });
            St.done();
            grid_reduce_then<1, kGridOwn>(c, acc, fl, out, f, [&] {
                if (!last) {
                    const double *fa, *fb;
                    second_loop_vectors(it + 1, fa, fb);
                    St.fetch(c, m2, fa, fb);
                }
            });
            if (last) gradient_overlap = out[0];                                                // :684
        }
        if (!isfinite(gradient_overlap)) {                                                      // :687-688
            if (threadIdx.x == 0) sc.term = 1;
        } else if (gradient_overlap >= 0.0) {                                                   // :689-692
            const double cc = -step_length * inv_gradient_norm;
            DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                reinterpret_cast<double2*>(a.d)[k] = make_double2(cc * gg.x, cc * gg.y);
            }
        }
        __syncthreads();
    }
    if (leader) *a.ctrl = sc;
    grid_ctx_end(c);
}

}  // namespace dzo
