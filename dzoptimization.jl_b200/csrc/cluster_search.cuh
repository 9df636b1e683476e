// cluster_search.cuh -- the O(n) stage of the large-n step! on a thread-block CLUSTER.
//
// The first version (vec_bfgs_search_kernel, one 1024-thread CTA) spent ~85 us per step at
// n = 16384 (8 % of the step): a single SM executes 64 FP64 lanes per clock and every probe
// touches all n elements.  Here the 4096 virtual threads of the canonical tree ARE 4096 real
// threads: a cluster of 8 CTAs x 512 threads.  Virtual thread v = 512*cta + 32*warp + lane, so
// the tree levels map onto the hardware exactly:
//     bits 4..0  (lane)   xor-shuffle butterfly 16,8,4,2,1
//     bits 5..8  (warp)   16 warp partials through shared memory, combined in ascending bit order
//     bits 9..11 (CTA)    8 CTA partials exchanged through DISTRIBUTED SHARED MEMORY
//                         (st.shared::cluster into every peer + one cluster barrier per probe)
// and the result is bit-identical to the single-CTA kernel and to the oracle's DZO_ORDER_TREE.
// Control flow stays uniform over the whole cluster because every reduced value is broadcast.
#pragma once
#include <cooperative_groups.h>

#include "large_bfgs.cuh"

namespace dzo {
namespace cg = cooperative_groups;

constexpr int kClusterCtas = 8;
constexpr int kClusterThreads = 512;

// shared-memory exchange area of one CTA
struct ClusterRed {
    double warp_part[4][16];               // [quantity][warp]
    unsigned warp_flag[16];
    double cta_part[2][4][kClusterCtas];   // [parity][quantity][source CTA]   (written by peers)
    unsigned cta_flag[2][kClusterCtas];
    int parity;
};

// Reduce K doubles over the canonical tree and OR one flag word across the cluster.  Every thread of
// every CTA returns the same values.  One __syncthreads + one cluster barrier.
template <int K>
DZO_DEVINL void cluster_tree_reduce(cg::cluster_group& cluster, ClusterRed& R, double (&p)[K], unsigned& flags) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned rank = cluster.block_rank();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double s = warp_butterfly_desc(p[k]);                       // bits 4,3,2,1,0
        if (lane == 0) R.warp_part[k][warp] = s;
    }
    const unsigned wf = __reduce_or_sync(0xffffffffu, flags);
    if (lane == 0) R.warp_flag[warp] = wf;
    __syncthreads();
    const int par = R.parity;
    if (warp == 0) {
        double c[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double v = R.warp_part[k][lane & 15];
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);   // bits 5,6,7,8
            c[k] = v;
        }
        unsigned f = R.warp_flag[lane & 15];
        f = __reduce_or_sync(0xffffffffu, f);
        if (lane < kClusterCtas) {                                        // lane = destination CTA
            ClusterRed* peer = cluster.map_shared_rank(&R, lane);
#pragma unroll
            for (int k = 0; k < K; ++k) peer->cta_part[par][k][rank] = c[k];
            peer->cta_flag[par][rank] = f;
        }
    }
    cluster.sync();                                                       // release/acquire over the cluster
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const double* c = R.cta_part[par][k];
        const double a = (c[0] + c[1]) + (c[2] + c[3]);                   // bits 9, 10
        const double b = (c[4] + c[5]) + (c[6] + c[7]);
        p[k] = a + b;                                                     // bit 11
    }
    unsigned f = 0;
#pragma unroll
    for (int c = 0; c < kClusterCtas; ++c) f |= R.cta_flag[par][c];
    flags = f;
    __syncthreads();                       // everyone has read slot `par` and warp_part before they are reused
    if (threadIdx.x == 0) R.parity = par ^ 1;
    // the next call's first __syncthreads orders this write before its read
}

// f(x + alpha*dir) for extended Rosenbrock; flag bit 0: any(x != w), bit 1: any(w != w_ref)
template <int MODE>
DZO_DEVINL double cluster_probe(cg::cluster_group& cluster, ClusterRed& R, const double* __restrict__ x,
                                const double* __restrict__ dir, long long m, double alpha, double alpha_ref,
                                bool& changed, bool& same_ref) {
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    const double2* x2 = reinterpret_cast<const double2*>(x);
    const double2* d2 = reinterpret_cast<const double2*>(dir);
    double acc = 0.0;
    unsigned fl = 0;
    for (long long k = v; k < m; k += DZO_TREE_WIDTH) {
        const double2 xx = x2[k];
        const double2 dd = d2[k];
        const double w0 = (MODE == 2) ? xx.x : xx.x + alpha * dd.x;
        const double w1 = (MODE == 2) ? xx.y : xx.y + alpha * dd.y;
        if ((xx.x != w0) | (xx.y != w1)) fl |= 1u;
        if (MODE == 1) {
            const double r0 = xx.x + alpha_ref * dd.x;
            const double r1 = xx.y + alpha_ref * dd.y;
            if ((!(w0 == r0)) | (!(w1 == r1))) fl |= 2u;
        }
        acc += RosenbrockVec::term(w0, w1);
    }
    double p[1] = {acc};
    cluster_tree_reduce<1>(cluster, R, p, fl);
    changed = (fl & 1u) != 0;
    same_ref = (MODE == 1) ? ((fl & 2u) == 0) : false;
    return p[0];
}

// bit 0: any(x != x + alpha*dir)   bit 1: any(dir != 0)
DZO_DEVINL unsigned cluster_point_flags(cg::cluster_group& cluster, ClusterRed& R, const double* __restrict__ x,
                                        const double* __restrict__ dir, long long n, double alpha) {
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    unsigned fl = 0;
    for (long long e = v; e < n; e += DZO_TREE_WIDTH) {
        const double xx = x[e], dd = dir[e];
        if (xx != xx + alpha * dd) fl |= 1u;
        if (!(dd == 0.0)) fl |= 2u;
    }
    double p[1] = {0.0};
    cluster_tree_reduce<1>(cluster, R, p, fl);
    return fl;
}

// quadratic_line_search [GLUE: legacy/DZOptimization.jl:49-172 with first step t1, then :191-216]
DZO_DEVINL void cluster_line_search(cg::cluster_group& cluster, ClusterRed& R, const double* __restrict__ x,
                                    const double* __restrict__ dir, long long n, double f0, double t1, double sign,
                                    int max_increases, double& t_best, double& f_best, long long& evals) {
    const long long m = n >> 1;
    double x1 = 0.0, f1 = f0, x2 = 0.0, f2 = f0;
    bool changed, same;
    do {
        if (!isfinite(f0)) break;                                         // :64-66
        if (!isfinite(t1) || t1 == 0.0) break;                            // [GLUE]
        double step = t1;
        unsigned fl = cluster_point_flags(cluster, R, x, dir, n, sign * step);
        if (!(fl & 2u)) break;                                            // :71-85 step_is_zero
        int cap = DZO_LINESEARCH_CAP;
        bool capped = false;
        while (!(fl & 1u)) {                                              // :91-101
            step += step;
            fl = cluster_point_flags(cluster, R, x, dir, n, sign * step);
            if (--cap == 0) { capped = true; break; }
        }
        if (capped) break;
        double fa = cluster_probe<0>(cluster, R, x, dir, m, sign * step, 0.0, changed, same);   // :126
        ++evals;
        if (fa <= f0) {                                                   // :130
            int num_increases = 0;
            cap = DZO_LINESEARCH_CAP;
            for (;;) {                                                    // :143-156
                const double ds = step + step;
                num_increases += 1;
                const double fb = cluster_probe<1>(cluster, R, x, dir, m, sign * ds, sign * step, changed, same);
                ++evals;
                --cap;
                if (((max_increases > 0) && (num_increases >= max_increases)) || !isfinite(fb) || fb > fa || same ||
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
                const double fb = cluster_probe<0>(cluster, R, x, dir, m, sign * hs, 0.0, changed, same);
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
        const double fq = cluster_probe<0>(cluster, R, x, dir, m, sign * xq, 0.0, changed, same);
        ++evals;
        if (fq < fb) { xb = xq; fb = fq; }
    }
    t_best = xb;
    f_best = fb;
}

// step! :891-960 up to (and including) the O(n) part of update_inverse_hessian! (:873-874).
// Launch: grid = 8 CTAs (one cluster of 8), 512 threads each.
static __global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
    cluster_bfgs_search_kernel(LargeVecs a_) {
    const LargeVecs a = for_problem(a_, blockIdx.x / kClusterCtas);      // one cluster per problem
    __shared__ ClusterRed R;
    __shared__ LargeCtrl sc;
    cg::cluster_group cluster = cg::this_cluster();
    const long long n = a.n, m = n >> 1;
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    const bool leader = (cluster.block_rank() == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0) { sc = *a.ctrl; R.parity = 0; }
    __syncthreads();
    cluster.sync();          // every CTA has read ctrl (the leader rewrites it at the end) and initialised R
    // row-sharded mode: the direction slabs of the previous BFGS-type step are stored into this GPU's
    // memory by the peers' update kernels; wait for all of them (flag = that step's sequence number)
    if (a.nranks > 1 && sc.kind == DZO_STEP_BFGS) {
        peer_wait(a.flags_d, a.nranks, (unsigned long long)sc.calls, &a.ctrl->pad);
        cluster.sync();                                                   // every CTA's wait is over before anyone looks
    }
    if (a.nranks > 1 && *reinterpret_cast<volatile int*>(&a.ctrl->pad) != 0) {
        // a peer's rows never arrived (now or in an earlier step!): this step! is a no-op on every kernel of the chain, the
        // optimizer state stays at the last consistent step and the host reports DZO_ERR_NCCL at the next sync
        if (leader) a.ctrl->kind = DZO_STEP_NULL;
        return;
    }
    if (sc.term) {                                                        // :893
        if (leader) a.ctrl->kind = DZO_STEP_NULL;
        return;
    }
    const double f0 = sc.f;
    const double step_length = sc.L;                                      // :918
    long long evals = 0;

    double grad_norm, bfgs_norm;                                          // :921, :928
    {
        double ag = 0.0, ad = 0.0;
        for (long long k = v; k < m; k += DZO_TREE_WIDTH) {
            const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
            const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
            ag += gg.x * gg.x; ag += gg.y * gg.y;
            ad += dd.x * dd.x; ad += dd.y * dd.y;
        }
        double p[2] = {ag, ad};
        unsigned fl = 0;
        cluster_tree_reduce<2>(cluster, R, p, fl);
        grad_norm = sqrt(p[0]);
        bfgs_norm = sqrt(p[1]);
    }
    double grad_step_length, grad_obj, bfgs_step_length, bfgs_obj;
    cluster_line_search(cluster, R, a.x, a.g, n, f0, step_length / grad_norm, -1.0, 0, grad_step_length, grad_obj, evals);  // :922-925
    cluster_line_search(cluster, R, a.x, a.d, n, f0, step_length / bfgs_norm, -1.0, 0, bfgs_step_length, bfgs_obj, evals);  // :929-932

    int kind;
    double alpha, fnew, Lnew;
    if (bfgs_obj < f0 && !(bfgs_obj > grad_obj)) {                        // :934
        kind = DZO_STEP_BFGS; alpha = -bfgs_step_length; fnew = bfgs_obj; Lnew = bfgs_step_length * bfgs_norm;
    } else if (grad_obj < f0) {                                           // :962
        kind = DZO_STEP_GRADIENT_DESCENT; alpha = -grad_step_length; fnew = grad_obj; Lnew = grad_step_length * grad_norm;
    } else {
        if (leader) {                                                     // :989
            a.ctrl->term = 1;
            a.ctrl->kind = DZO_STEP_NULL;
            a.ctrl->evals = sc.evals + evals;
            a.ctrl->kind_log[sc.calls & 63] = DZO_STEP_NULL;
            a.ctrl->calls = sc.calls + 1;
        }
        return;
    }
    const double* dir = (kind == DZO_STEP_BFGS) ? a.d : a.g;
    double acc = 0.0;                                                     // :943-950 / :971-978, overlap :873
    for (long long k = v; k < m; k += DZO_TREE_WIDTH) {
        const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
        const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
        const double2 dd = reinterpret_cast<const double2*>(dir)[k];
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
            acc += dd.x * dgv.x;
            acc += dd.y * dgv.y;
        } else {
            reinterpret_cast<double2*>(a.d)[k] = gn;                      // :984-986
        }
    }
    double overlap = 0.0;
    if (kind == DZO_STEP_BFGS) {
        double p[1] = {acc};
        unsigned fl = 0;
        cluster_tree_reduce<1>(cluster, R, p, fl);
        overlap = p[0];
        const double inv_overlap = 1.0 / overlap;                         // :874
        for (long long k = v; k < m; k += DZO_TREE_WIDTH) {
            const double2 dd = reinterpret_cast<const double2*>(a.d)[k];
            reinterpret_cast<double2*>(a.sd)[k] = make_double2(dd.x * inv_overlap, dd.y * inv_overlap);
        }
    }
    if (leader) {
        LargeCtrl c = sc;
        c.f = fnew; c.L = Lnew; c.type = kind; c.iter = sc.iter + 1;
        c.kind = kind; c.step_length = alpha; c.overlap = overlap; c.delta_norm = 0.0;
        c.evals = sc.evals + evals;
        c.kind_log[sc.calls & 63] = (unsigned char)kind;
        c.calls = sc.calls + 1;
        c.pad = *reinterpret_cast<volatile int*>(&a.ctrl->pad);           // keep a peer-timeout mark set during this kernel
        *a.ctrl = c;
    }
}

// :876 on the cluster
static __global__ void __cluster_dims__(kClusterCtas, 1, 1) __launch_bounds__(kClusterThreads, 1)
    cluster_delta_kernel(LargeVecs a_) {
    const LargeVecs a = for_problem(a_, blockIdx.x / kClusterCtas);
    __shared__ ClusterRed R;
    cg::cluster_group cluster = cg::this_cluster();
    if (a.ctrl->kind != DZO_STEP_BFGS) return;
    if (threadIdx.x == 0) R.parity = 0;
    __syncthreads();
    cluster.sync();
    if (a.nranks > 1) peer_wait(a.flags_t, a.nranks, (unsigned long long)a.ctrl->calls, &a.ctrl->pad);   // all slabs of t have landed
    const long long v = (long long)cluster.block_rank() * kClusterThreads + threadIdx.x;
    double acc = 0.0;
    for (long long k = v; 2 * k < a.n; k += DZO_TREE_WIDTH) {
        acc += a.dg[2 * k] * a.t[2 * k];
        if (2 * k + 1 < a.n) acc += a.dg[2 * k + 1] * a.t[2 * k + 1];
    }
    double p[1] = {acc};
    unsigned fl = 0;
    cluster_tree_reduce<1>(cluster, R, p, fl);
    if (cluster.block_rank() == 0 && threadIdx.x == 0) {
        if (a.nranks > 1 && *reinterpret_cast<volatile int*>(&a.ctrl->pad) != 0)
            a.ctrl->kind = DZO_STEP_NULL;      // t is incomplete (a peer timed out): the update sweep must not touch H
        else
            a.ctrl->delta_norm = a.ctrl->step_length * a.ctrl->overlap + p[0];
    }
}

// Row-sharded fused mode, last launch of a step!: the peers' rows of next_step_direction are stored into this GPU's
// memory by THEIR update kernels, so a drained local stream alone would not mean d is complete.  This one-warp kernel
// waits for every rank's flag of this step; behind it "stream drained" implies "all slabs have landed"
// (dzo_bfgs_get_direction / dzo_bfgs_sync right after step!), and the next search kernel's wait is already satisfied.
static __global__ void __launch_bounds__(32, 1) peer_arrival_kernel(LargeVecs a) {
    if (a.nranks > 1 && a.ctrl->kind == DZO_STEP_BFGS)
        peer_wait(a.flags_d, a.nranks, (unsigned long long)a.ctrl->calls, &a.ctrl->pad);
}

}  // namespace dzo
