// batched_hybrid.cuh -- batched small-n BFGS step!, second-generation mapping (n in {2,4,8,16}).
//
// Same arithmetic, same SEQUENTIAL summation order and therefore the same bits as
// batched_bfgs.cuh / the sequential oracle; only the thread mapping differs.  ncu on the
// first-generation kernel (profiles/r01_ncu_full_summary.csv) showed it issue-bound: one
// 16-lane group per problem means every line-search probe costs a private instruction stream
// per problem (~960 warp instructions per problem-step, 24 % of them FP64 math).  Here a warp
// owns 32 problems and works in two phases per step!:
//
//   phase 1  ONE THREAD PER PROBLEM (32 problems per instruction): norms, both bracketing line
//            searches as a single per-thread state machine whose loop body is exactly one probe
//            evaluation (so lanes in different stages of different searches share the expensive
//            code), the BFGS / GD / terminate decision, x, g, dx, dg, overlap, d/overlap.
//            Vectors live in a per-warp shared-memory tile [32 problems][17] (conflict-free).
//   phase 2  N LANES PER PROBLEM, uniform control flow: lane r recomputes element r of x, g, dx, dg
//            (bit-identical elementwise formulas) and stores them coalesced, streams ROW r of the
//            2 KB inverse Hessian from HBM into registers with coalesced loads (for a fixed column
//            the N lanes read N consecutive doubles; the next round's rows are prefetched while the
//            current round computes), t = H*dg, delta, the rank-2 update fused with d = H'*g, rows
//            streamed back.  A gradient-descent step writes the identity without reading H; a
//            terminated problem moves no H bytes.  (A first version let every lane read its row as
//            128 contiguous bytes: ncu showed the L1TEX tag stage at 80 % -- 32 lines per request.)
//
// legacy/DZOptimization.jl:891-994 (step!), :864-889 (update_inverse_hessian!), :49-216 (line search).
#pragma once
#include "batched_bfgs.cuh"

namespace dzo {

constexpr int kHybridWarps = 4;
constexpr int kHybridThreads = 32 * kHybridWarps;
constexpr int kHybridStride = 18;  // doubles per problem row in the shared tile: 16-byte aligned rows (phase 2 reads
                                   // pairs with one 128-bit broadcast load); phase 1's per-thread reads are 2-way conflicted

template <int N>
struct HybridSmem {
    alignas(16) double X[32][kHybridStride];   // current_point          -> after phase 1: delta_gradient
    alignas(16) double G[32][kHybridStride];   // current_gradient       -> after phase 1: new gradient
    alignas(16) double D[32][kHybridStride];   // next_step_direction    -> after phase 1: step_direction / overlap
    alignas(16) double T[32 / N][N];           // scratch = H * delta_gradient of the problems of the current round
    alignas(16) double P[32 / N][N];           // products for the sequential dot of :876
};

enum : int { HS_INIT = 0, HS_EXPAND = 1, HS_SHRINK = 2, HS_QUAD = 3, HS_DONE = 4 };

DZO_DEVINL void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int N>
__global__ void __launch_bounds__(kHybridThreads, 4) bfgs_batched_hybrid_kernel(BatchedArgs A) {
    static_assert(N == 2 || N == 4 || N == 8 || N == 16, "hybrid mapping: n in {2,4,8,16}");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NN = N * N;
    constexpr int PPR = 32 / N;  // problems per phase-2 round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    HybridSmem<N>& S = reinterpret_cast<HybridSmem<N>*>(smem_raw)[warp];
    const long long p0 = ((long long)blockIdx.x * kHybridWarps + warp) * 32;  // first problem of this warp
    if (p0 >= A.batch) return;
    const long long p = p0 + lane;
    const bool valid = p < A.batch;
    const int nprob = (int)((A.batch - p0 < 32) ? (A.batch - p0) : 32);

    bool term = valid ? (A.term[p] != 0) : true;
    double f0 = valid ? A.f[p] : 0.0;
    double L = valid ? A.L[p] : 0.0;
    long long iter = valid ? A.iter[p] : 0;
    int type = DZO_STEP_NULL;
    bool moved = false;

    for (int s = 0; s < A.ksteps; ++s) {
        if (!__any_sync(FULL, !term)) break;
        // ---------------------------------------------------------------- stage the vectors (coalesced)
        __syncwarp();
        for (int e = lane; e < nprob * N; e += 32) {
            const int q = e / N, i = e - q * N;
            S.X[q][i] = A.x[p0 * N + e];
            S.G[q][i] = A.g[p0 * N + e];
            S.D[q][i] = A.d[p0 * N + e];
        }
        __syncwarp();
        double* X = S.X[lane];
        double* G = S.G[lane];
        double* D = S.D[lane];
        // Pull the inverse Hessians of the first phase-2 rounds into L2 now: they arrive while phase 1
        // computes.  One 128-byte line per lane and request; tiles of terminated problems are skipped.
        constexpr int LINES_PER_ROUND = PPR * NN * 8 / 128 > 0 ? PPR * NN * 8 / 128 : 1;   // 32 at n = 16
        const unsigned live = __ballot_sync(FULL, !term);
        auto prefetch_round = [&](int round) {
            if (round >= N) return;
            for (int l = lane; l < LINES_PER_ROUND; l += 32) {
                const int q = round * PPR + (l * 128) / (NN * 8 > 128 ? NN * 8 : 128) % PPR;
                const char* base = reinterpret_cast<const char*>(A.H + (p0 + round * PPR) * NN);
                if (NN * 8 >= 128) {
                    if ((live >> q) & 1u) prefetch_l2(base + l * 128);
                } else {
                    prefetch_l2(base + l * 128);
                }
            }
        };
        for (int rr = 0; rr < A.prefetch_rounds; ++rr) prefetch_round(rr);

        // ---------------------------------------------------------------- phase 1: one thread per problem
        int kind = DZO_STEP_NULL;
        double alpha = 0.0, overlap = 0.0;
        {
            double grad_norm = 0.0, bfgs_norm = 0.0;
            if (!term) {
                double sg = 0.0, sd = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) { sg += G[i] * G[i]; sd += D[i] * D[i]; }   // :921, :928
                grad_norm = sqrt(sg);
                bfgs_norm = sqrt(sd);
            }
            // Per-thread line-search state machine over BOTH searches (0: gradient direction, 1: BFGS
            // direction); the loop body is exactly one probe evaluation, so lanes in different stages of
            // different searches share the expensive code.
            int search = 0, st = term ? HS_DONE : HS_INIT;
            double res_t0 = 0.0, res_f0 = f0, res_t1 = 0.0, res_f1 = f0;
            double step = 0.0, fa = 0.0, x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0, xb = 0.0, fb = f0, trial = 0.0, tref = 0.0;
            int cap = DZO_LINESEARCH_CAP;
            bool fresh = true;        // the current search has not been set up yet
            bool bracket_done = false;
            for (;;) {
                if (st != HS_DONE && fresh) {
                    // find_three_point_bracket prologue  :64-85 (+ [GLUE] guards), first trial step t1 = L / norm
                    fresh = false;
                    const double* dir = search ? D : G;
                    const double t1 = L / (search ? bfgs_norm : grad_norm);              // :922, :929
                    x1 = 0.0; f1 = f0; x2 = 0.0; f2 = f0;
                    bool zero = true;
#pragma unroll
                    for (int i = 0; i < N; ++i) zero &= (dir[i] == 0.0);
                    st = HS_INIT; step = t1; trial = t1; tref = 0.0; cap = DZO_LINESEARCH_CAP;
                    if (!isfinite(f0) || !isfinite(t1) || t1 == 0.0 || zero) bracket_done = true;   // bracket (0,f0,0,f0)
                }
                if (bracket_done) {                                                       // :196-214
                    bracket_done = false;
                    xb = 0.0; fb = f0;
                    if (f1 < fb) { xb = x1; fb = f1; }
                    if (f2 < fb) { xb = x2; fb = f2; }
                    const double delta_1 = f0 - f1;
                    const double delta_2 = f2 - f1;
                    const double sum_deltas = delta_1 + delta_2;
                    if (delta_1 >= 0.0 && delta_2 >= 0.0 && sum_deltas > 0.0) {
                        const double twice_delta_1 = delta_1 + delta_1;
                        const double delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas);
                        trial = delta_ratio * x1;
                        st = HS_QUAD;
                    } else {
                        if (search == 0) { res_t0 = xb; res_f0 = fb; } else { res_t1 = xb; res_f1 = fb; }
                        st = (search == 1) ? HS_DONE : HS_INIT;
                        search += 1; fresh = true;
                        continue;    // set the next search up before probing (uniformity is restored at the vote)
                    }
                }
                if (!__any_sync(FULL, st != HS_DONE)) break;
                double fv = 0.0;
                bool changed = false, same = true;
                if (st != HS_DONE) {
                    // lse(t): w = x + (-t)*dir ; f(w)   (sign fixed by :945,:973)
                    const double* dir = search ? D : G;
                    const double a = -trial, ar = -tref;
#pragma unroll
                    for (int k = 0; k < N / 2; ++k) {
                        const double xa = X[2 * k], xc = X[2 * k + 1], da = dir[2 * k], dc = dir[2 * k + 1];
                        const double w0 = xa + a * da, w1 = xc + a * dc;
                        changed |= (xa != w0) | (xc != w1);
                        const double r0 = xa + ar * da, r1 = xc + ar * dc;
                        same &= (w0 == r0) & (w1 == r1);
                        const double t1_ = 1 - w0;
                        const double t2_ = w1 - w0 * w0;
                        fv += t1_ * t1_ + 100 * (t2_ * t2_);                            // legacy/ExampleFunctions.jl:10-15
                    }
                }
                // ---- advance the state machine (cheap, divergent)
                if (st == HS_INIT) {
                    if (!changed) {                                                       // :91-101
                        step += step; trial = step;
                        if (--cap == 0) bracket_done = true;                              // [GLUE]
                    } else {                                                              // :126
                        fa = fv;
                        cap = DZO_LINESEARCH_CAP;
                        if (fa <= f0) { st = HS_EXPAND; tref = step; trial = step + step; }   // :130-136
                        else { st = HS_SHRINK; trial = 0.5 * step; }                          // :157
                    }
                } else if (st == HS_EXPAND) {                                             // :143-156
                    --cap;
                    if (!isfinite(fv) || fv > fa || same || cap == 0) {
                        x1 = step; f1 = fa; x2 = trial; f2 = fv;
                        bracket_done = true;
                    } else {
                        step = trial; fa = fv; tref = step; trial = step + step;
                    }
                } else if (st == HS_SHRINK) {                                             // :162-170
                    --cap;
                    if (fv <= f0 || cap == 0) {
                        x1 = trial; f1 = fv; x2 = step; f2 = fa;
                        bracket_done = true;
                    } else {
                        step = trial; fa = fv; trial = 0.5 * step;
                    }
                } else if (st == HS_QUAD) {                                               // :210-213
                    if (fv < fb) { xb = trial; fb = fv; }
                    if (search == 0) { res_t0 = xb; res_f0 = fb; } else { res_t1 = xb; res_f1 = fb; }
                    st = (search == 1) ? HS_DONE : HS_INIT;
                    search += 1; fresh = true;
                }
            }
            const double res_t[2] = {res_t0, res_t1}, res_f[2] = {res_f0, res_f1};
            // ---- decision and bookkeeping  :934-990
            if (!term) {
                const double grad_step_length = res_t[0], grad_obj = res_f[0];
                const double bfgs_step_length = res_t[1], bfgs_obj = res_f[1];
                if (bfgs_obj < f0 && !(bfgs_obj > grad_obj)) {                            // :934
                    kind = DZO_STEP_BFGS; alpha = -bfgs_step_length;
                    L = bfgs_step_length * bfgs_norm; f0 = bfgs_obj;                      // :937-938
                } else if (grad_obj < f0) {                                               // :962
                    kind = DZO_STEP_GRADIENT_DESCENT; alpha = -grad_step_length;
                    L = grad_step_length * grad_norm; f0 = grad_obj;                      // :965-966
                } else {
                    term = true;                                                          // :989
                }
            }
            if (kind != DZO_STEP_NULL) { type = kind; iter += 1; moved = true; }         // :939-940 / :967-968
            if (kind == DZO_STEP_BFGS) {
                // overlap = dot(step_direction, delta_gradient)  :873 -- thread-local and strictly sequential.
                // The elementwise results (x, g, dx, dg, d/overlap) are recomputed bit-identically by the
                // phase-2 lanes, which can store them coalesced.
#pragma unroll
                for (int k = 0; k < N / 2; ++k) {
                    const double xa = X[2 * k], xc = X[2 * k + 1];
                    const double ga = G[2 * k], gc = G[2 * k + 1];
                    const double da = D[2 * k], dc = D[2 * k + 1];
                    const double na = xa + alpha * da, nc = xc + alpha * dc;              // :945
                    const double t1_ = 1 - na;
                    const double t2_ = nc - na * na;
                    const double gna = -2 * t1_ - 400 * na * t2_;                         // :948 rosenbrock_gradient!
                    const double gnc = 200 * t2_;
                    const double dga = (-ga) + gna, dgc = (-gc) + gnc;                    // :944, :950
                    overlap += da * dga;
                    overlap += dc * dgc;
                }
            }
        }
        __syncwarp();

        // ---------------------------------------------------------------- phase 2: N lanes per problem
        // Lane r of problem q owns element r of every vector and row r of H.  Every global access below
        // is coalesced: for a fixed instruction the N lanes of a problem touch N consecutive doubles.
        const int r = lane % N, sub = lane / N;
        const unsigned gmask = (N == 32) ? FULL : (((1u << N) - 1u) << (sub * N));   // lanes of my problem
        const double ao = alpha * overlap;                                            // first operand of :876
        const double inv_overlap = 1.0 / overlap;                                     // :874 inv(overlap)
        auto load_rows = [&](int round, double (&row)[N], int& kq) {
            const int q = round * PPR + sub;
            kq = __shfl_sync(FULL, kind, q);
            if (kq == DZO_STEP_BFGS) {
                const double* Hp = A.H + (p0 + q) * NN + r;          // H[r, j] at Hp[j*N] (column-major)
#pragma unroll
                for (int j = 0; j < N; ++j) row[j] = __ldcs(Hp + j * N);
            }
        };
        double cur[N], nxt[N];
        int kcur, knxt = DZO_STEP_NULL;
        load_rows(0, cur, kcur);
#pragma unroll 1
        for (int round = 0; round < N; ++round) {
            const int q = round * PPR + sub;
            const double alpha_q = __shfl_sync(FULL, alpha, q);
            const double ao_q = __shfl_sync(FULL, ao, q);
            const double inv_overlap_q = __shfl_sync(FULL, inv_overlap, q);
            if (round + 1 < N) load_rows(round + 1, nxt, knxt);
            prefetch_round(round + A.prefetch_rounds);
            if (kcur != DZO_STEP_NULL) {
                const long long e = (p0 + q) * N + r;
                const double xo = S.X[q][r], go = S.G[q][r], dol = S.D[q][r];
                const double dirv = (kcur == DZO_STEP_BFGS) ? dol : go;
                const double xn = xo + alpha_q * dirv;                                    // :945 / :973
                const double xp = __shfl_xor_sync(gmask, xn, 1);                          // the other element of my pair
                const double xe = (r & 1) ? xp : xn, xod = (r & 1) ? xn : xp;
                const double t1_ = 1 - xe;
                const double t2_ = xod - xe * xe;
                const double gn = (r & 1) ? (200 * t2_) : (-2 * t1_ - 400 * xe * t2_);    // :948 rosenbrock_gradient!
                const double dgv = (-go) + gn;                                            // :944, :950
                A.x[e] = xn;
                A.g[e] = gn;
                A.dx[e] = (-xo) + xn;                                                     // :943, :949
                A.dg[e] = dgv;
                double* Hp = A.H + (p0 + q) * NN + r;
                if (kcur == DZO_STEP_BFGS) {
                    // update_inverse_hessian!  :874-886 fused with mul!(d, H, g)  :958-960
                    const double sd = dol * inv_overlap_q;                                // :874
                    __syncwarp(gmask);              // every lane has read the old tile values
                    S.X[q][r] = dgv;
                    S.G[q][r] = gn;
                    S.D[q][r] = sd;
                    __syncwarp(gmask);
                    const double2* dg2 = reinterpret_cast<const double2*>(S.X[q]);
                    const double2* g2 = reinterpret_cast<const double2*>(S.G[q]);
                    const double2* sd2 = reinterpret_cast<const double2*>(S.D[q]);
                    const double2* t2 = reinterpret_cast<const double2*>(S.T[sub]);
                    const double2* p2 = reinterpret_cast<const double2*>(S.P[sub]);
                    double t = 0.0;
#pragma unroll
                    for (int j = 0; j < N; j += 2) {                                      // :875
                        const double2 v = dg2[j >> 1];
                        t += cur[j] * v.x;
                        t += cur[j + 1] * v.y;
                    }
                    S.T[sub][r] = t;
                    S.P[sub][r] = dgv * t;
                    __syncwarp(gmask);
                    double dot = 0.0;
#pragma unroll
                    for (int j = 0; j < N; j += 2) {                                      // Kernels.dot order
                        const double2 v = p2[j >> 1];
                        dot += v.x;
                        dot += v.y;
                    }
                    const double delta_norm = ao_q + dot;                                 // :876
                    double dnew = 0.0;
#pragma unroll
                    for (int j = 0; j < N; j += 2) {
                        const double2 sj = sd2[j >> 1], tj = t2[j >> 1], gj = g2[j >> 1];
                        cur[j] += (delta_norm * (sd * sj.x) - (t * sj.x + sd * tj.x));    // :882-884
                        cur[j + 1] += (delta_norm * (sd * sj.y) - (t * sj.y + sd * tj.y));
                        dnew += cur[j] * gj.x;                                            // :958-960
                        dnew += cur[j + 1] * gj.y;
                    }
#pragma unroll
                    for (int j = 0; j < N; ++j) __stcs(Hp + j * N, cur[j]);
                    A.d[e] = dnew;
                    __syncwarp(gmask);
                } else {
#pragma unroll
                    for (int j = 0; j < N; ++j) __stcs(Hp + j * N, (j == r) ? 1.0 : 0.0);  // :981 identity_matrix!
                    A.d[e] = gn;                                                          // :984-986
                }
            }
#pragma unroll
            for (int j = 0; j < N; ++j) cur[j] = nxt[j];
            kcur = knxt;
        }
        __syncwarp();
    }

    if (valid) {
        if (moved) {
            A.f[p] = f0;
            A.L[p] = L;
            A.iter[p] = iter;
            A.type[p] = type;
            if (A.f_host) A.f_host[p] = f0;       // 32 lanes = 256 contiguous bytes of posted PCIe writes
        }
        if (term) {
            A.term[p] = 1;
            if (A.term_host) A.term_host[p] = 1;
        }
    }
}

// ============================================================================= third-generation kernel
// Same two-phase mapping and the same bits as above; what changed (round 2):
//
//   * DUAL   the gradient-direction and the BFGS-direction line searches (:922-925, :929-932) are independent,
//            so one thread runs BOTH state machines side by side: every loop iteration evaluates one probe of each.
//            The loop runs max(probes_g, probes_d) instead of probes_g + probes_d times per lane, the two
//            dependent FP64 chains interleave (ncu on the second-generation kernel: 1.9 `wait` + 1.8 short-scoreboard
//            stalls per issue), and the point x is read from shared memory once for both.
//   * LAZY   identity_matrix! (:981) after a gradient-descent step is not written to HBM: a per-problem byte
//            `hid` says "H is the identity, the 2 KB in HBM are stale".  The next BFGS-type step builds its rows of I in
//            registers instead of loading them (bit-identical: the same arithmetic runs on the same values), the
//            constructor writes no H at all, and dzo_bfgs_get_inverse_hessian materialises I for the caller.
//            A GD-type step therefore moves 1.1 KB instead of 3.1 KB and a BFGS-type step after it 3.1 KB instead of 5.2 KB.
//   * the (w == reference point) test of the expansion loop (:150) is evaluated only while some lane is expanding.
//   * per-launch step-kind counters (warp ballots -> one atomicAdd per warp and kind) feed bench.py's roofline.
struct HybridSearch {
    int st;
    int cap;
    double step, fa, xb, fb, trial, tref;
};

// QuadraticLineSearch  :196-214 on a finished bracket (0, f0), (x1, f1), (x2, f2): best of the three, then either the
// parabola-vertex probe (HS_QUAD) or done.  The bracket itself never outlives this call.
DZO_DEVINL void hs_bracket(HybridSearch& s, double f0, double x1, double f1, double x2, double f2) {
    s.xb = 0.0; s.fb = f0;
    if (f1 < s.fb) { s.xb = x1; s.fb = f1; }
    if (f2 < s.fb) { s.xb = x2; s.fb = f2; }
    const double delta_1 = f0 - f1;
    const double delta_2 = f2 - f1;
    const double sum_deltas = delta_1 + delta_2;
    if (delta_1 >= 0.0 && delta_2 >= 0.0 && sum_deltas > 0.0) {
        const double twice_delta_1 = delta_1 + delta_1;
        const double delta_ratio = (twice_delta_1 + sum_deltas) / (sum_deltas + sum_deltas);
        s.trial = delta_ratio * x1;
        s.st = HS_QUAD;
    } else {
        s.st = HS_DONE;
    }
}
// find_three_point_bracket prologue  :64-85 (+ [GLUE] guards), first trial step t1 = L / norm.  The degenerate
// bracket (0,f0,0,f0) of the early returns resolves to (t*, f*) = (0, f0) without a probe (:196-206 with all deltas 0).
template <int N>
DZO_DEVINL void hs_setup(HybridSearch& s, const double* dir, double f0, double t1, bool active) {
    s.xb = 0.0; s.fb = f0; s.fa = 0.0;
    s.step = t1; s.trial = t1; s.tref = 0.0; s.cap = DZO_LINESEARCH_CAP;
    s.st = active ? HS_INIT : HS_DONE;
    if (active) {
        bool zero = true;
#pragma unroll
        for (int i = 0; i < N; ++i) zero &= (dir[i] == 0.0);
        if (!isfinite(f0) || !isfinite(t1) || t1 == 0.0 || zero) s.st = HS_DONE;
    }
}
// one probe result -> next state  (:91-101, :126-170, :210-213)
DZO_DEVINL void hs_advance(HybridSearch& s, double f0, double fv, bool changed, bool same) {
    if (s.st == HS_INIT) {
        if (!changed) {                                                       // :91-101
            s.step += s.step; s.trial = s.step;
            if (--s.cap == 0) s.st = HS_DONE;                                 // [GLUE] bracket (0,f0,0,f0)
        } else {                                                              // :126
            s.fa = fv;
            s.cap = DZO_LINESEARCH_CAP;
            if (s.fa <= f0) { s.st = HS_EXPAND; s.tref = s.step; s.trial = s.step + s.step; }   // :130-136
            else { s.st = HS_SHRINK; s.trial = 0.5 * s.step; }                                  // :157
        }
    } else if (s.st == HS_EXPAND) {                                           // :143-156
        --s.cap;
        if (!isfinite(fv) || fv > s.fa || same || s.cap == 0) {
            hs_bracket(s, f0, s.step, s.fa, s.trial, fv);
        } else {
            s.step = s.trial; s.fa = fv; s.tref = s.step; s.trial = s.step + s.step;
        }
    } else if (s.st == HS_SHRINK) {                                           // :162-170
        --s.cap;
        if (fv <= f0 || s.cap == 0) {
            hs_bracket(s, f0, s.trial, fv, s.step, s.fa);
        } else {
            s.step = s.trial; s.fa = fv; s.trial = 0.5 * s.step;
        }
    } else if (s.st == HS_QUAD) {                                             // :210-213
        if (fv < s.fb) { s.xb = s.trial; s.fb = fv; }
        s.st = HS_DONE;
    }
}

// lse(t): w = x + (-t)*dir ; f(w)   (sign fixed by :945,:973) for both searches at once; x is read once.
template <int N, bool SAME>
DZO_DEVINL void hs_probe2(const double* X, const double* G, const double* D, double a0, double ar0, double a1, double ar1,
                          double& fv0, bool& ch0, bool& sm0, double& fv1, bool& ch1, bool& sm1) {
    const double2* X2 = reinterpret_cast<const double2*>(X);
    const double2* G2 = reinterpret_cast<const double2*>(G);
    const double2* D2 = reinterpret_cast<const double2*>(D);
    fv0 = 0.0; fv1 = 0.0; ch0 = false; ch1 = false; sm0 = true; sm1 = true;
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        const double2 x = X2[k], g = G2[k], d = D2[k];
        {
            const double w0 = x.x + a0 * g.x, w1 = x.y + a0 * g.y;
            ch0 |= (x.x != w0) | (x.y != w1);
            if (SAME) {
                const double r0 = x.x + ar0 * g.x, r1 = x.y + ar0 * g.y;
                sm0 &= (w0 == r0) & (w1 == r1);
            }
            const double t1_ = 1 - w0;
            const double t2_ = w1 - w0 * w0;
            fv0 += t1_ * t1_ + 100 * (t2_ * t2_);                              // legacy/ExampleFunctions.jl:10-15
        }
        {
            const double w0 = x.x + a1 * d.x, w1 = x.y + a1 * d.y;
            ch1 |= (x.x != w0) | (x.y != w1);
            if (SAME) {
                const double r0 = x.x + ar1 * d.x, r1 = x.y + ar1 * d.y;
                sm1 &= (w0 == r0) & (w1 == r1);
            }
            const double t1_ = 1 - w0;
            const double t2_ = w1 - w0 * w0;
            fv1 += t1_ * t1_ + 100 * (t2_ * t2_);
        }
    }
}
template <int N, bool SAME>
DZO_DEVINL void hs_probe1(const double* X, const double* Dir, double a, double ar, double& fv, bool& ch, bool& sm) {
    const double2* X2 = reinterpret_cast<const double2*>(X);
    const double2* D2 = reinterpret_cast<const double2*>(Dir);
    fv = 0.0; ch = false; sm = true;
#pragma unroll
    for (int k = 0; k < N / 2; ++k) {
        const double2 x = X2[k], d = D2[k];
        const double w0 = x.x + a * d.x, w1 = x.y + a * d.y;
        ch |= (x.x != w0) | (x.y != w1);
        if (SAME) {
            const double r0 = x.x + ar * d.x, r1 = x.y + ar * d.y;
            sm &= (w0 == r0) & (w1 == r1);
        }
        const double t1_ = 1 - w0;
        const double t2_ = w1 - w0 * w0;
        fv += t1_ * t1_ + 100 * (t2_ * t2_);
    }
}

// step kinds as counted for the roofline (include/dzopt.h, dzo_bfgs_get_step_kind_counts)
enum : int { HK_BFGS_READ = 0, HK_BFGS_IDENT = 1, HK_GD = 2, HK_TERMINATE = 3, HK_IDLE = 4, HK_IDLE_WARP = 5, HK_COUNT = 6 };

// Row q of the shared tile is 18 doubles wide for 16 (or fewer) elements: the two spare doubles of the three rows of a
// problem carry what phase 1 hands to phase 2 (the shared tile is exactly what fits four CTAs per SM, and registers are
// the other limit: every value kept in a register across phase 2 showed up as a local-memory spill that misses the
// 28 KB of L1 left beside the tile -- ncu: 17 % of all stall samples on two spill reloads).
//   X[q][16]  lane q < HK_COUNT: running step-kind counter q of this warp     X[q][17]  alpha = -step length   (:945/:973)
//   G[q][16]  alpha * overlap  (first operand of :876)                         G[q][17]  1 / overlap             (:874)
//   D[q][16]  low word: step kind | where H comes from << 8                    D[q][17]  (free)
template <int N, bool DUAL>
__global__ void __launch_bounds__(kHybridThreads, 4) bfgs_batched_hybrid3_kernel(BatchedArgs A) {
    static_assert(N == 2 || N == 4 || N == 8 || N == 16, "hybrid mapping: n in {2,4,8,16}");
    static_assert(kHybridStride >= 18 && HK_COUNT <= 32, "hand-off slots live in the row padding");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NN = N * N;
    constexpr int PPR = 32 / N;  // problems per phase-2 round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    HybridSmem<N>& S = reinterpret_cast<HybridSmem<N>*>(smem_raw)[warp];
    const bool lazy = (A.hid != nullptr);
    double& cnt_slot = S.X[lane][16];
    cnt_slot = 0.0;                                            // exact for counts < 2^53
    auto count = [&](int k, unsigned ballot) { if (lane == k) cnt_slot += (double)__popc(ballot); };

    // Optional persistent mode ("batched_persistent", off by default): tiles of 32 problems are handed out to WARPS through a
    // global counter (4 CTAs per SM), so a warp whose line searches finish early takes the next tile instead of waiting for
    // the slowest warp of its CTA to free the CTA slot, and the partial last wave of the grid disappears.  Measured on the
    // 1 M x n=16 workload it LOSES to one tile per warp (0.80 vs 0.74 ms per launch; ncu: 2.0 instead of 0.3 no-instruction
    // stall cycles per issue -- 16 warps at 16 different places of a 54 KB kernel body thrash the instruction cache, while
    // the warps of freshly launched CTAs walk through it together); kept as a tested variant.  The counter returns to
    // zero with the last fetch of a launch.
    const long long ntiles = (A.batch + 31) / 32;
    const unsigned nwarps_total = gridDim.x * kHybridWarps;
    auto fetch_tile = [&]() -> long long {
        if (!A.tile_counter) return ntiles;                    // one tile per warp (A/B baseline)
        unsigned t = 0;
        if (lane == 0) {
            t = atomicAdd(A.tile_counter, 1u);
            if ((long long)t == ntiles + (long long)nwarps_total - 1) atomicExch(A.tile_counter, 0u);   // the launch's last fetch
        }
        t = __shfl_sync(FULL, t, 0);
        return ((long long)t < ntiles) ? (long long)t : ntiles;
    };
    auto read_flags = [&](long long t) -> unsigned {            // has_terminated | (H is an implicit identity) << 1
        const long long pp = t * 32 + lane;
        if (t >= ntiles || pp >= A.batch) return 1u;
        return (A.term[pp] != 0 ? 1u : 0u) | ((lazy && A.hid[pp] != 0) ? 2u : 0u);
    };
    long long tile = A.tile_counter ? fetch_tile() : (long long)blockIdx.x * kHybridWarps + warp;
    if (tile >= ntiles) return;
    unsigned flags = read_flags(tile);

  while (tile < ntiles) {
    const long long p0 = tile * 32;                            // first problem of this tile
    const long long p = p0 + lane;
    const bool valid = p < A.batch;
    const int nprob = (int)((A.batch - p0 < 32) ? (A.batch - p0) : 32);
    long long next_tile = ntiles;
    unsigned next_flags = 1u;
    bool have_next = !A.tile_counter;                          // the next tile has been fetched (or there is none to fetch)

    // A tile starts cold: everything it needs first is requested at once -- the three vectors as 16-byte asynchronous
    // copies straight into the shared tile (24 per lane, no registers held), the scalars as ordinary loads in flight
    // beside them; the flags were read while the previous tile was in phase 2.
    auto stage_vectors = [&]() {
        for (int e2 = lane; e2 < nprob * (N / 2); e2 += 32) {
            const int q = (2 * e2) / N, i = 2 * e2 - q * N;
            cp_async16(&S.X[q][i], A.x + p0 * N + 2 * e2);
            cp_async16(&S.G[q][i], A.g + p0 * N + 2 * e2);
            cp_async16(&S.D[q][i], A.d + p0 * N + 2 * e2);
        }
        cp_async_commit();
    };
    __syncwarp();                                              // the previous tile's phase 2 is done with the shared tile
    stage_vectors();
    bool term = (flags & 1u) != 0;
    bool ident = (flags & 2u) != 0;                            // H == I and the copy in HBM is stale
    bool moved_any = false;

    for (int s = 0; s < A.ksteps; ++s) {
        if (!__any_sync(FULL, !term)) {
            if (lane == HK_IDLE_WARP) cnt_slot += (double)nprob * (double)(A.ksteps - s);
            cp_async_wait_all();                               // nothing may still be landing in the tile when it is reused
            break;
        }
        if (s > 0) {
            // a later step of the same launch: the vectors were stored by other lanes of this warp in phase 2
            __threadfence_block();
            __syncwarp();
            stage_vectors();
        }
        double f0 = valid ? A.f[p] : 0.0;              // (a later step of the same launch reads back its own stores)
        double L = valid ? A.L[p] : 0.0;
        const double* X = S.X[lane];
        const double* G = S.G[lane];
        const double* D = S.D[lane];
        // Pull the inverse Hessians of the first phase-2 rounds into L2 now: they arrive while phase 1
        // computes.  One 128-byte line per lane and request; tiles of terminated problems and of problems whose
        // H is the (unmaterialised) identity are skipped.
        constexpr int LINES_PER_ROUND = PPR * NN * 8 / 128 > 0 ? PPR * NN * 8 / 128 : 1;   // 32 at n = 16
        const bool term_in = term;
        const unsigned live = __ballot_sync(FULL, !term && !ident);
        auto prefetch_round = [&](int round) {
            if (round >= N) return;
            for (int l = lane; l < LINES_PER_ROUND; l += 32) {
                const int q = round * PPR + (l * 128) / (NN * 8 > 128 ? NN * 8 : 128) % PPR;
                const char* base = reinterpret_cast<const char*>(A.H + (p0 + round * PPR) * NN);
                if (NN * 8 >= 128) {
                    if ((live >> q) & 1u) prefetch_l2(base + l * 128);
                } else {
                    if (live) prefetch_l2(base + l * 128);
                }
            }
        };
        for (int rr = 0; rr < A.prefetch_rounds; ++rr) prefetch_round(rr);
        cp_async_wait_all();
        __syncwarp();

        // ---------------------------------------------------------------- phase 1: one thread per problem
        {
            int kind = DZO_STEP_NULL;
            double alpha = 0.0, overlap = 0.0;
            double grad_norm = 0.0, bfgs_norm = 0.0;
            if (!term) {
                double sg = 0.0, sd = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) { sg += G[i] * G[i]; sd += D[i] * D[i]; }   // :921, :928
                grad_norm = sqrt(sg);
                bfgs_norm = sqrt(sd);
            }
            double grad_step_length, grad_obj, bfgs_step_length, bfgs_obj;
            if (DUAL) {
                // both searches side by side: one probe of each per iteration
                HybridSearch s0, s1;
                hs_setup<N>(s0, G, f0, L / grad_norm, !term);                             // :922
                hs_setup<N>(s1, D, f0, L / bfgs_norm, !term);                             // :929
                for (;;) {
                    if (!__any_sync(FULL, (s0.st != HS_DONE) | (s1.st != HS_DONE))) break;
                    double fv0, fv1;
                    bool ch0, ch1, sm0, sm1;
                    if (__any_sync(FULL, (s0.st == HS_EXPAND) | (s1.st == HS_EXPAND)))
                        hs_probe2<N, true>(X, G, D, -s0.trial, -s0.tref, -s1.trial, -s1.tref, fv0, ch0, sm0, fv1, ch1, sm1);
                    else
                        hs_probe2<N, false>(X, G, D, -s0.trial, -s0.tref, -s1.trial, -s1.tref, fv0, ch0, sm0, fv1, ch1, sm1);
                    hs_advance(s0, f0, fv0, ch0, sm0);
                    hs_advance(s1, f0, fv1, ch1, sm1);
                }
                grad_step_length = s0.xb; grad_obj = s0.fb; bfgs_step_length = s1.xb; bfgs_obj = s1.fb;
            } else {
                // one search after the other, but as ONE per-thread state machine whose loop body is exactly one probe
                // evaluation: lanes in different stages of different searches share the expensive code
                HybridSearch sc;
                int which = 0;
                grad_step_length = 0.0; grad_obj = f0;
                hs_setup<N>(sc, G, f0, L / grad_norm, !term);                             // :922
                for (;;) {
                    if (sc.st == HS_DONE && which == 0 && !term) {
                        grad_step_length = sc.xb; grad_obj = sc.fb;
                        which = 1;
                        hs_setup<N>(sc, D, f0, L / bfgs_norm, true);                      // :929
                    }
                    if (!__any_sync(FULL, sc.st != HS_DONE)) break;
                    const double* dir = which ? D : G;
                    double fv;
                    bool ch, sm;
                    if (__any_sync(FULL, sc.st == HS_EXPAND)) hs_probe1<N, true>(X, dir, -sc.trial, -sc.tref, fv, ch, sm);
                    else hs_probe1<N, false>(X, dir, -sc.trial, -sc.tref, fv, ch, sm);
                    hs_advance(sc, f0, fv, ch, sm);
                }
                bfgs_step_length = sc.xb; bfgs_obj = sc.fb;
            }
            // ---- decision and bookkeeping  :934-990
            if (!term) {
                if (bfgs_obj < f0 && !(bfgs_obj > grad_obj)) {                            // :934
                    kind = DZO_STEP_BFGS; alpha = -bfgs_step_length;
                    L = bfgs_step_length * bfgs_norm; f0 = bfgs_obj;                      // :937-938
                } else if (grad_obj < f0) {                                               // :962
                    kind = DZO_STEP_GRADIENT_DESCENT; alpha = -grad_step_length;
                    L = grad_step_length * grad_norm; f0 = grad_obj;                      // :965-966
                } else {
                    term = true;                                                          // :989
                }
            }
            if (kind == DZO_STEP_BFGS) {
                // overlap = dot(step_direction, delta_gradient)  :873 -- thread-local and strictly sequential.
                // The elementwise results (x, g, dx, dg, d/overlap) are recomputed bit-identically by the
                // phase-2 lanes, which can store them coalesced.
#pragma unroll
                for (int k = 0; k < N / 2; ++k) {
                    const double xa = X[2 * k], xc = X[2 * k + 1];
                    const double ga = G[2 * k], gc = G[2 * k + 1];
                    const double da = D[2 * k], dc = D[2 * k + 1];
                    const double na = xa + alpha * da, nc = xc + alpha * dc;              // :945
                    const double t1_ = 1 - na;
                    const double t2_ = nc - na * na;
                    const double gna = -2 * t1_ - 400 * na * t2_;                         // :948 rosenbrock_gradient!
                    const double gnc = 200 * t2_;
                    const double dga = (-ga) + gna, dgc = (-gc) + gnc;                    // :944, :950
                    overlap += da * dga;
                    overlap += dc * dgc;
                }
            }
            // what this step does to each problem, for the roofline (uniform across the warp)
            count(HK_BFGS_READ, __ballot_sync(FULL, kind == DZO_STEP_BFGS && !ident));
            count(HK_BFGS_IDENT, __ballot_sync(FULL, kind == DZO_STEP_BFGS && ident));
            count(HK_GD, __ballot_sync(FULL, kind == DZO_STEP_GRADIENT_DESCENT));
            count(HK_TERMINATE, __ballot_sync(FULL, valid && term && !term_in));
            count(HK_IDLE, __ballot_sync(FULL, valid && term_in));
            // ---- the per-problem scalars are final for this step: store them now, hand the rest to phase 2 through the tile
            const int hsrc = (kind == DZO_STEP_BFGS) ? (ident ? 2 : 1) : 0;   // rows from HBM (1) or the implicit identity (2)
            if (valid) {
                if (kind != DZO_STEP_NULL) {
                    A.f[p] = f0;
                    A.L[p] = L;
                    atomicAdd(reinterpret_cast<unsigned long long*>(A.iter + p), 1ull);  // :940 / :968 (no load to wait for)
                    A.type[p] = kind;                                                     // :939 / :967
                    moved_any = true;
                }
                if (term && !term_in) {
                    A.term[p] = 1;
                    if (A.term_host) A.term_host[p] = 1;
                }
                const bool ident_next = (kind == DZO_STEP_BFGS) ? false : ((kind == DZO_STEP_GRADIENT_DESCENT) ? lazy : ident);
                if (lazy && ident_next != ident) A.hid[p] = ident_next ? 1 : 0;
                ident = ident_next;
            }
            S.X[lane][17] = alpha;
            S.G[lane][16] = alpha * overlap;                                              // first operand of :876
            S.G[lane][17] = 1.0 / overlap;                                                // :874 inv(overlap)
            reinterpret_cast<int*>(&S.D[lane][16])[0] = kind | (hsrc << 8);
        }
        __syncwarp();
        if (!have_next) {
            // the next tile of this warp: its index and its flags are in flight during phase 2.  (Prefetching its vectors
            // into L2 as well was measured and removed: 57 us later the lines are gone again -- +0.38 GB of DRAM reads per launch.)
            have_next = true;
            next_tile = fetch_tile();
            next_flags = read_flags(next_tile);
        }

        // ---------------------------------------------------------------- phase 2: N lanes per problem
        // Lane r of problem q owns element r of every vector and row r of H.  Every global access below
        // is coalesced: for a fixed instruction the N lanes of a problem touch N consecutive doubles.
        const int r = lane % N, sub = lane / N;
        const unsigned gmask = (N == 32) ? FULL : (((1u << N) - 1u) << (sub * N));   // lanes of my problem
        auto load_rows = [&](int round, double (&row)[N], int& meta) {
            const int q = round * PPR + sub;
            meta = reinterpret_cast<const int*>(&S.D[q][16])[0];
            if ((meta >> 8) == 1) {
                const double* Hp = A.H + (p0 + q) * NN + r;          // H[r, j] at Hp[j*N] (column-major)
#pragma unroll
                for (int j = 0; j < N; ++j) row[j] = __ldcs(Hp + j * N);
            }
        };
        auto process = [&](int round, double (&cur)[N], int meta) {
            const int kcur = meta & 0xff;
            if (kcur == DZO_STEP_NULL) return;
            const int q = round * PPR + sub;
            const long long e = (p0 + q) * N + r;
            const double xo = S.X[q][r], go = S.G[q][r], dol = S.D[q][r];
            const double alpha_q = S.X[q][17];
            const double dirv = (kcur == DZO_STEP_BFGS) ? dol : go;
            const double xn = xo + alpha_q * dirv;                                    // :945 / :973
            const double xp = __shfl_xor_sync(gmask, xn, 1);                          // the other element of my pair
            const double xe = (r & 1) ? xp : xn, xod = (r & 1) ? xn : xp;
            const double t1_ = 1 - xe;
            const double t2_ = xod - xe * xe;
            const double gn = (r & 1) ? (200 * t2_) : (-2 * t1_ - 400 * xe * t2_);    // :948 rosenbrock_gradient!
            const double dgv = (-go) + gn;                                            // :944, :950
            A.x[e] = xn;
            A.g[e] = gn;
            A.dx[e] = (-xo) + xn;                                                     // :943, :949
            A.dg[e] = dgv;
            double* Hp = A.H + (p0 + q) * NN + r;
            if (kcur == DZO_STEP_BFGS) {
                // update_inverse_hessian!  :874-886 fused with mul!(d, H, g)  :958-960
                const double ao_q = S.G[q][16], inv_overlap_q = S.G[q][17];
                if ((meta >> 8) == 2) {
#pragma unroll
                    for (int j = 0; j < N; ++j) cur[j] = (j == r) ? 1.0 : 0.0;        // the rows identity_matrix! would have left
                }
                const double sd = dol * inv_overlap_q;                                // :874
                __syncwarp(gmask);              // every lane has read the old tile values
                S.X[q][r] = dgv;
                S.G[q][r] = gn;
                S.D[q][r] = sd;
                __syncwarp(gmask);
                const double2* dg2 = reinterpret_cast<const double2*>(S.X[q]);
                const double2* g2 = reinterpret_cast<const double2*>(S.G[q]);
                const double2* sd2 = reinterpret_cast<const double2*>(S.D[q]);
                const double2* t2 = reinterpret_cast<const double2*>(S.T[sub]);
                const double2* p2 = reinterpret_cast<const double2*>(S.P[sub]);
                double t = 0.0;
#pragma unroll
                for (int j = 0; j < N; j += 2) {                                      // :875
                    const double2 v = dg2[j >> 1];
                    t += cur[j] * v.x;
                    t += cur[j + 1] * v.y;
                }
                S.T[sub][r] = t;
                S.P[sub][r] = dgv * t;
                __syncwarp(gmask);
                double dot = 0.0;
#pragma unroll
                for (int j = 0; j < N; j += 2) {                                      // Kernels.dot order
                    const double2 v = p2[j >> 1];
                    dot += v.x;
                    dot += v.y;
                }
                const double delta_norm = ao_q + dot;                                 // :876
                double dnew = 0.0;
#pragma unroll
                for (int j = 0; j < N; j += 2) {
                    const double2 sj = sd2[j >> 1], tj = t2[j >> 1], gj = g2[j >> 1];
                    cur[j] += (delta_norm * (sd * sj.x) - (t * sj.x + sd * tj.x));    // :882-884
                    cur[j + 1] += (delta_norm * (sd * sj.y) - (t * sj.y + sd * tj.y));
                    dnew += cur[j] * gj.x;                                            // :958-960
                    dnew += cur[j + 1] * gj.y;
                }
#pragma unroll
                for (int j = 0; j < N; ++j) __stcs(Hp + j * N, cur[j]);
                A.d[e] = dnew;
                __syncwarp(gmask);
            } else {
                if (!lazy) {
#pragma unroll
                    for (int j = 0; j < N; ++j) __stcs(Hp + j * N, (j == r) ? 1.0 : 0.0);  // :981 identity_matrix!
                }
                A.d[e] = gn;                                                          // :984-986
            }
        };
        // two rounds per trip, ping-pong between two register tiles: the loads of round i+1 are in flight while
        // round i computes, and no tile is ever copied
        double rowsA[N], rowsB[N];
        int metaA, metaB = 0;
        load_rows(0, rowsA, metaA);
#pragma unroll 1
        for (int round = 0; round < N; round += 2) {
            load_rows(round + 1, rowsB, metaB);
            prefetch_round(round + A.prefetch_rounds);
            process(round, rowsA, metaA);
            if (round + 2 < N) load_rows(round + 2, rowsA, metaA);
            prefetch_round(round + 1 + A.prefetch_rounds);
            process(round + 1, rowsB, metaB);
        }
        __syncwarp();
    }

    if (valid && moved_any && A.f_host) A.f_host[p] = A.f[p];     // 32 lanes = 256 contiguous bytes of posted PCIe writes
    if (!have_next) { next_tile = fetch_tile(); next_flags = read_flags(next_tile); }   // (a tile that was idle on entry)
    tile = next_tile;
    flags = next_flags;
  }
    if (A.stats && lane < HK_COUNT && cnt_slot != 0.0) atomicAdd(A.stats + lane, (unsigned long long)cnt_slot);
}

template <int N>
inline size_t hybrid_smem() { return sizeof(HybridSmem<N>) * kHybridWarps; }

}  // namespace dzo
