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

template <int N>
inline size_t hybrid_smem() { return sizeof(HybridSmem<N>) * kHybridWarps; }

}  // namespace dzo
