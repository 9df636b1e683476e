// grid_lbfgs.cuh -- the live LBFGSOptimizer step! (src/DZOptimization.jl:454-509) for n > DZO_TREE_BLOCK on the
// WHOLE GPU: a cooperative grid of 8-CTA clusters, one cluster per block of DZO_TREE_BLOCK elements.
//
// cluster_lbfgs_kernel (lbfgs_kernels.cuh) runs a step on one cluster = 8 SMs, which caps an O(n*m) step at what
// 8 SMs can pull out of L2 (n = 2^20, m = 10: 1.98 ms per step, ~0.5 TB/s).  DZO_ORDER_TREE_BLOCKED (include/dzopt.h)
// keeps the canonical tree inside a block of 65536 elements and adds the block results in ascending order, so
// every block can live on its own cluster: thread v of the cluster IS virtual thread v of that block's tree and
// owns its 8 element pairs of every vector (between reductions a thread only re-reads what it wrote itself).
// A reduction = in-cluster DSMEM tree (cluster_search.cuh) -> one double per block in global memory -> ONE grid
// barrier -> every thread adds the <= 64 block values in ascending order.  The block values are double-buffered
// by the parity of the reduction count so that a cluster racing ahead cannot overwrite what a slower one still reads.
#pragma once
#include "lbfgs_kernels.cuh"

namespace dzo {

constexpr int kGridMaxBlocks = 1024;                     // n <= 64 Mi elements
constexpr long long kBlockPairs = DZO_TREE_BLOCK / 2;    // 32768 pairs = 8 per virtual thread

struct GridLbfgsArgs {
    double *x, *dx, *g, *dg, *d;
    double *S, *Y;                           // m x n, physical slot p at S + p*n
    LbfgsCtrl* ctrl;
    double* part;                            // [2][3][kGridMaxBlocks] block results (parity, quantity, block)
    unsigned* fpart;                         // [2][kGridMaxBlocks]    block flag words
    long long n;
    int m, ksteps, nblocks;
    int mode;                                // 0 = steps, 1 = constructor
    double initial_step_length;
};

struct GridCtx {
    cg::grid_group grid;
    cg::cluster_group cluster;
    ClusterRed* R;
    int nclusters, cid, nblocks;
    long long v;                             // virtual thread of this thread inside a block
    int red;                                 // reductions so far (parity)
    double* s_out;                           // shared: totals of the last reduction
    unsigned* s_flags;
};

// Reduce K per-block accumulators (acc[k][j] = this thread's partial of quantity k for its j-th owned block) and OR
// the per-block flag words.  Returns the totals in out[k], the OR in flags; identical on every thread of the grid.
template <int K, int MAXB>
DZO_DEVINL void grid_reduce(GridCtx& c, const GridLbfgsArgs& a, double (&acc)[K][MAXB], unsigned (&fl)[MAXB], double (&out)[K],
                            unsigned& flags) {
    const int par = c.red & 1;
    c.red += 1;
#pragma unroll
    for (int j = 0; j < MAXB; ++j) {
        const int b = c.cid + j * c.nclusters;
        if (b < c.nblocks) {                 // uniform over the cluster
            double p[K];
#pragma unroll
            for (int k = 0; k < K; ++k) p[k] = acc[k][j];
            unsigned f = fl[j];
            cluster_tree_reduce<K>(c.cluster, *c.R, p, f);
            if (c.cluster.block_rank() == 0 && threadIdx.x == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) a.part[(par * 3 + k) * kGridMaxBlocks + b] = p[k];
                a.fpart[par * kGridMaxBlocks + b] = f;
            }
        }
    }
    c.grid.sync();
    // warp 0 of every CTA fetches the block values side by side (one L2 round trip per 32 blocks -- a thread adding
    // them one load at a time would wait out a full L2 latency per block) and adds them in ascending block order
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        double tot[K];
        unsigned f = 0;
        for (int b0 = 0; b0 < c.nblocks; b0 += 32) {
            const int b = b0 + lane;
            double val[K];
#pragma unroll
            for (int k = 0; k < K; ++k) val[k] = (b < c.nblocks) ? __ldcg(&a.part[(par * 3 + k) * kGridMaxBlocks + b]) : 0.0;
            const unsigned fv = (b < c.nblocks) ? __ldcg(&a.fpart[par * kGridMaxBlocks + b]) : 0u;
            f |= __reduce_or_sync(0xffffffffu, fv);
            const int cnt = min(32, c.nblocks - b0);
            for (int i = 0; i < cnt; ++i) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double x = __shfl_sync(0xffffffffu, val[k], i);
                    tot[k] = (b0 == 0 && i == 0) ? x : tot[k] + x;
                }
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) c.s_out[k] = tot[k];
            *c.s_flags = f;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = c.s_out[k];
    flags = *c.s_flags;
    // the next write to s_out sits behind the next grid barrier
}

constexpr int kGridOwn = 8;   // blocks one cluster may own: n <= 8 * nclusters * 65536 (9.4 Mi elements on 18 clusters)

// for (own pairs) body: j = index of the owned block (compile-time after unrolling, so per-block accumulators stay in
// registers), k = global pair index
#define DZO_GRID_OWN_PAIRS(c, m2, j, k)                                                                              \
    _Pragma("unroll") for (int j = 0; j < kGridOwn; ++j)                                                             \
        if ((c).cid + j * (c).nclusters < (c).nblocks)                                                               \
            for (long long k = (long long)((c).cid + j * (c).nclusters) * kBlockPairs + (c).v,                       \
                           e__ = (long long)((c).cid + j * (c).nclusters + 1) * kBlockPairs;                         \
                 k < e__ && k < (m2); k += DZO_TREE_WIDTH)

DZO_DEVINL double grid_dot(GridCtx& c, const GridLbfgsArgs& a, const double* __restrict__ u, const double* __restrict__ w) {
    const long long m2 = a.n >> 1;
    double acc[1][kGridOwn];
    unsigned fl[kGridOwn];
#pragma unroll
    for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
    DZO_GRID_OWN_PAIRS(c, m2, j, k) {
        const double2 uu = reinterpret_cast<const double2*>(u)[k];
        const double2 ww = reinterpret_cast<const double2*>(w)[k];
        acc[0][j] += uu.x * ww.x;
        acc[0][j] += uu.y * ww.y;
    }
    double out[1];
    unsigned f;
    grid_reduce<1, kGridOwn>(c, a, acc, fl, out, f);
    return out[0];
}

static __global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
    grid_lbfgs_kernel(GridLbfgsArgs a) {
    __shared__ ClusterRed R;
    __shared__ LbfgsCtrl sc;
    __shared__ double alpha[DZO_LBFGS_MAX_HISTORY];
    __shared__ double s_out[3];
    __shared__ unsigned s_flags;
    GridCtx c{cg::this_grid(), cg::this_cluster(), &R, 0, 0, a.nblocks, 0, 0, s_out, &s_flags};
    c.nclusters = (int)(gridDim.x / kClusterCtas);
    c.cid = (int)(blockIdx.x / kClusterCtas);
    c.v = (long long)c.cluster.block_rank() * kClusterThreads + threadIdx.x;
    const long long n = a.n, m2 = n >> 1;
    const bool leader = (blockIdx.x == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0) { sc = *a.ctrl; R.parity = 0; }
    __syncthreads();
    c.cluster.sync();
    c.grid.sync();                 // every CTA holds the control block before the leader may rewrite it

    if (a.mode == 1) {
        // LBFGSOptimizer(c!, f, g!, x0, L0, m)  :347-427 (x already holds the initial point)
        double acc[2][kGridOwn];
        unsigned fl[kGridOwn];
#pragma unroll
        for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; fl[j] = 0; }
        DZO_GRID_OWN_PAIRS(c, m2, j, k) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            const double2 gg = RosenbrockVec::grad(xx.x, xx.y);                                 // :422
            reinterpret_cast<double2*>(a.g)[k] = gg;
            reinterpret_cast<double2*>(a.dx)[k] = make_double2(0.0, 0.0);                       // :367
            reinterpret_cast<double2*>(a.dg)[k] = make_double2(0.0, 0.0);                       // :372
            acc[0][j] += RosenbrockVec::term(xx.x, xx.y);                                       // :417
            acc[1][j] += gg.x * gg.x;
            acc[1][j] += gg.y * gg.y;
        }
        double out[2];
        unsigned f;
        grid_reduce<2, kGridOwn>(c, a, acc, fl, out, f);
        const double gnorm = sqrt(out[1]);                                                      // :376
        const bool stuck = (gnorm == 0.0);                                                      // :377
        const double cc = -a.initial_step_length / gnorm;
        DZO_GRID_OWN_PAIRS(c, m2, j, k) {                                                       // :378-383
            const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
            reinterpret_cast<double2*>(a.d)[k] = stuck ? make_double2(0.0, 0.0) : make_double2(gg.x * cc, gg.y * cc);
        }
        if (leader) {
            LbfgsCtrl t;
            t.f = out[0]; t.df = 0.0; t.iter = 0; t.stuck = stuck; t.count = 0; t.head = 0; t.pad = 0; t.evals = 1; t.yy = 0.0;
            for (int i = 0; i < DZO_LBFGS_MAX_HISTORY; ++i) t.rho[i] = 0.0;
            *a.ctrl = t;
        }
        return;
    }

    // step! :454-509, k times.  sc is the CTA-local copy of the control block; every CTA updates its copy identically
    // (all values come out of grid-wide reductions) and the leader publishes it at the end.
    for (int step_i = 0; step_i < a.ksteps; ++step_i) {
        if (sc.stuck) break;                                                                    // :456-458
        const int cnt = sc.count, head = sc.head, m = a.m;
        double step = 1.0, next = 0.0;
        bool accepted = false, have_probe = false;
        long long evals = 0;
        double pr_out[1];
        unsigned pr_flags = 0;
        if (sc.iter > 0 && cnt > 0) {
            // compute_lbfgs_step_direction!  :430-451 (logical slot i -> physical (head + i) mod m), restructured so that
            // every pass over the direction also accumulates the dot product the NEXT stage needs: 2*cnt + 1 passes
            // and reductions per direction instead of 4*cnt + 3 passes.  Per element the operations and their order
            // are those of the reference loop (d = g; d += -alpha*y; d *= c; d += -(alpha+beta)*s).
            double acc[1][kGridOwn];
            unsigned fl[kGridOwn];
            double out[1];
            unsigned f;
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
            {
                const double* s0 = a.S + (long long)head * n;
                DZO_GRID_OWN_PAIRS(c, m2, j, k) {                                               // d = g, and s_0 . d  (:439)
                    const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                    const double2 ss = reinterpret_cast<const double2*>(s0)[k];
                    reinterpret_cast<double2*>(a.d)[k] = gg;
                    acc[0][j] += ss.x * gg.x;
                    acc[0][j] += ss.y * gg.y;
                }
            }
            grid_reduce<1, kGridOwn>(c, a, acc, fl, out, f);
            double al = out[0] / sc.rho[head];
            const double cc = -sc.rho[head] / sc.yy;                                            // :443 (dot(y_0, y_0) cached)
            double beta = 0.0;
            for (int i = 0; i < cnt; ++i) {
                const int p = (head + i) % m;
                if (threadIdx.x == 0) alpha[i] = al;
                const double na = -al;
                const double* y = a.Y + (long long)p * n;
                const bool last = (i == cnt - 1);
                const int pn = last ? p : (head + i + 1) % m;
                const double* nxt = last ? (a.Y + (long long)pn * n) : (a.S + (long long)pn * n);
#pragma unroll
                for (int j = 0; j < kGridOwn; ++j) acc[0][j] = 0.0;
                DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                    double2 dd = reinterpret_cast<double2*>(a.d)[k];
                    const double2 yy = reinterpret_cast<const double2*>(y)[k];
                    const double2 nn = reinterpret_cast<const double2*>(nxt)[k];
                    dd.x += na * yy.x; dd.y += na * yy.y;                                       // :440
                    if (last) { dd.x *= cc; dd.y *= cc; }                                       // :443
                    reinterpret_cast<double2*>(a.d)[k] = dd;
                    acc[0][j] += nn.x * dd.x;                                                   // s_{i+1} . d (:439) or y_{cnt-1} . d (:446)
                    acc[0][j] += nn.y * dd.y;
                }
                grid_reduce<1, kGridOwn>(c, a, acc, fl, out, f);
                if (last) beta = out[0] / sc.rho[p];
                else al = out[0] / sc.rho[pn];
            }
            __syncthreads();   // alpha[] visible
            for (int i = cnt - 1; i >= 0; --i) {
                const int p = (head + i) % m;
                const double na = -(alpha[i] + beta);
                const double* sp = a.S + (long long)p * n;
#pragma unroll
                for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
                if (i > 0) {
                    const int pn = (head + i - 1) % m;
                    const double* yn = a.Y + (long long)pn * n;
                    DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                        double2 dd = reinterpret_cast<double2*>(a.d)[k];
                        const double2 ss = reinterpret_cast<const double2*>(sp)[k];
                        const double2 yy = reinterpret_cast<const double2*>(yn)[k];
                        dd.x += na * ss.x; dd.y += na * ss.y;                                   // :447
                        reinterpret_cast<double2*>(a.d)[k] = dd;
                        acc[0][j] += yy.x * dd.x;                                               // y_{i-1} . d  (:446)
                        acc[0][j] += yy.y * dd.y;
                    }
                    grid_reduce<1, kGridOwn>(c, a, acc, fl, out, f);
                    beta = out[0] / sc.rho[pn];
                } else {
                    // last correction fused with the first trial of take_backtracking_step!(opt, 1, d)  (:124-138)
                    DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                        double2 dd = reinterpret_cast<double2*>(a.d)[k];
                        const double2 ss = reinterpret_cast<const double2*>(sp)[k];
                        const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
                        dd.x += na * ss.x; dd.y += na * ss.y;                                   // :447
                        reinterpret_cast<double2*>(a.d)[k] = dd;
                        const double w0 = xx.x + step * dd.x, w1 = xx.y + step * dd.y;         // :124 axpy!
                        if (!julia_isequal(w0, xx.x) || !julia_isequal(w1, xx.y)) fl[j] |= 1u;  // :128
                        acc[0][j] += RosenbrockVec::term(w0, w1);                               // :138
                    }
                    grid_reduce<1, kGridOwn>(c, a, acc, fl, pr_out, pr_flags);
                    have_probe = true;
                }
            }
        }
        // take_backtracking_step!(opt, 1, step_direction)  :107-154
        for (;;) {
            if (!have_probe) {
                double acc[1][kGridOwn];
                unsigned fl[kGridOwn];
#pragma unroll
                for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
                DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                    const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
                    const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
                    const double w0 = xx.x + step * dd.x, w1 = xx.y + step * dd.y;             // :124 axpy!
                    if (!julia_isequal(w0, xx.x) || !julia_isequal(w1, xx.y)) fl[j] |= 1u;      // :128
                    acc[0][j] += RosenbrockVec::term(w0, w1);                                   // :138
                }
                grid_reduce<1, kGridOwn>(c, a, acc, fl, pr_out, pr_flags);
            }
            have_probe = false;
            if (!(pr_flags & 1u)) break;                                                        // :128-131 stuck
            ++evals;
            next = pr_out[0];
            if (next < sc.f) { accepted = true; break; }                                        // :139
            step *= 0.5;                                                                        // :152
        }
        if (!accepted) {
            // the trial point equals the current point: delta_point keeps the COPY of the point (:118)
            DZO_GRID_OWN_PAIRS(c, m2, j, k) reinterpret_cast<double2*>(a.dx)[k] = reinterpret_cast<const double2*>(a.x)[k];
            if (threadIdx.x == 0) { sc.stuck = 1; sc.evals += evals; }
            __syncthreads();
            break;
        }
        // accept: x, delta_point (:145), gradient and delta_gradient (:478-480), history push (:482-505)
        const int slot = (head - 1 + m) % m;
        double* Snew = a.S + (long long)slot * n;
        double* Ynew = a.Y + (long long)slot * n;
        double acc[2][kGridOwn];
        unsigned fl[kGridOwn];
#pragma unroll
        for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; fl[j] = 0; }
        DZO_GRID_OWN_PAIRS(c, m2, j, k) {
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
            acc[0][j] += dxv.x * dgv.x;                                                         // :505
            acc[0][j] += dxv.y * dgv.y;
            acc[1][j] += dgv.x * dgv.x;                                                         // dot(y_0, y_0) of the NEXT step (:443)
            acc[1][j] += dgv.y * dgv.y;
        }
        double out[2];
        unsigned f;
        grid_reduce<2, kGridOwn>(c, a, acc, fl, out, f);
        const double rho_new = out[0];
        if (threadIdx.x == 0) {
            sc.yy = out[1];
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

}  // namespace dzo
