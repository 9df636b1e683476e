// grid_lbfgs.cuh -- the live LBFGSOptimizer step! (src/DZOptimization.jl:454-509) for n > DZO_TREE_BLOCK on the
// WHOLE GPU: one cooperative grid, eight 512-thread CTAs per block of DZO_TREE_BLOCK elements.
//
// cluster_lbfgs_kernel (lbfgs_kernels.cuh) runs a step on one cluster = 8 SMs, which caps an O(n*m) step at what
// 8 SMs can pull out of L2 (n = 2^20, m = 10: 1.98 ms per step, ~0.5 TB/s).  DZO_ORDER_TREE_BLOCKED (include/dzopt.h)
// keeps the canonical tree inside a block of 65536 elements and adds the block results in ascending order, so the
// blocks can live on different SMs.  A CTA owns an EIGHTH of a block: its 512 threads are the virtual threads
// 512*part .. 512*part+511 of that block's tree and own their 8 element pairs of every vector (between reductions a
// thread only re-reads what it wrote itself).  A reduction = lane bits by shuffle, warp bits through shared memory
// -> one double per CTA in global memory -> ONE grid barrier -> warp 0 of every CTA fetches all CTA values side by
// side, finishes each block's tree (the three CTA bits, ascending) and adds the blocks in ascending order.  The CTA
// values are double-buffered by the parity of the reduction count so that a CTA racing ahead cannot overwrite what a
// slower one still reads.  (A first version used one 8-CTA cluster per block with DSMEM for the CTA bits; only 15
// such clusters are co-resident on a B200, one short of the 16 blocks of n = 2^20.)
#pragma once
#include <cstdio>

#include "lbfgs_kernels.cuh"

namespace dzo {

constexpr int kGridMaxBlocks = 1024;                     // n <= 64 Mi elements
constexpr int kGridMaxParts = 8 * kGridMaxBlocks;        // CTA values per quantity
constexpr int kGridQ = 4;                                // quantities one reduction can carry
constexpr long long kBlockPairs = DZO_TREE_BLOCK / 2;    // 32768 pairs = 8 per virtual thread

struct GridLbfgsArgs {
    double *x, *dx, *g, *dg, *d;
    double *S, *Y;                           // m x n, physical slot p at S + p*n
    LbfgsCtrl* ctrl;
    double* part;                            // [2][kGridQ][kGridMaxParts] CTA results (parity, quantity, eighth)
    unsigned* fpart;                         // [2][kGridMaxParts]    CTA flag words
    long long n;
    int m, ksteps, nblocks;
    int mode;                                // 0 = steps, 1 = constructor
    double initial_step_length;
    int stage;                               // 1: kGridStageBytes of dynamic shared memory are there (OWN == 1 launches)
};

struct GridCtx {
    cg::grid_group grid;
    int nctas, cta, nblocks;                 // eighth e of the vector (block e / 8, part e % 8) belongs to CTA e mod nctas
    int red;                                 // reductions so far (parity)
    double* part;                            // global [2][kGridQ][kGridMaxParts] (barrier mode), then the flagged lines below
    unsigned* fpart;                         // global [2][kGridMaxParts]
    double* s_warp;                          // shared [kGridQ][16]: warp partials
    unsigned* s_wflag;                       // shared [16]
    double* s_out;                           // shared [kGridQ]: totals of the last reduction
    unsigned* s_flags;
    long long* prof = nullptr;               // optional (leader thread): cycles spent [0] in the CTA-level part of reductions,
                                             // [1] polling for the other CTAs' lines, [2] in the final combine, [3] reductions
    uint4* ll = nullptr;                     // flagged lines [2][kGridQ + 1][kGridMaxParts] (null: grid-barrier mode)
    unsigned seq = 0;                        // sequence number of the next reduction (flag value of its lines)
    unsigned backoff = 0;                    // ns a consumer sleeps after a poll that found a line missing
};

// ---- reductions WITHOUT a grid barrier (round 2).  A CTA value travels as one 16-byte line {lo, tag, hi, tag}: two 8-byte
// halves, each atomic and each carrying the sequence number of the reduction (the protocol NCCL calls LL); the flag bits
// of the reduction ride in the top four bits of quantity 0's tag.  Every CTA owns an INBOX; a producer pushes its line
// into the inbox of every CTA of the grid (148 fire-and-forget stores to 148 different addresses) and a consumer polls
// only its own inbox until both tags of every line show the current sequence number.  The producer's store and the
// consumer's load are the whole synchronisation: no atomic counter and no hot line -- a first version with ONE shared
// set of lines polled by all 128 CTAs was slower than grid.sync() (0.175 vs 0.135 ms per L-BFGS step! at n = 2^20:
// a thousand requests per line and poll round serialise in the L2 slice that owns it).  No other data crosses CTAs
// between reductions (a thread only re-reads vector elements it wrote itself), so nothing else needs the barrier's fence.
// Lines are double-buffered by the parity of the reduction count: the line of reduction r is overwritten by reduction
// r + 2, which every CTA reaches only after it has consumed r + 1, i.e. after every CTA has finished reading r.  The
// running sequence number lives behind the inboxes and survives across launches (stale lines never match).
#ifndef DZO_LL_SYS
#define DZO_LL_SYS 0            // 1: ld.volatile / st.cg lines (A/B)
#endif
#ifndef DZO_LL_COMBINE_SMEM
#define DZO_LL_COMBINE_SMEM 0   // 1: final combine through shared memory instead of a chain of shuffles (A/B)
#endif
#ifndef DZO_EARLY_FETCH
#define DZO_EARLY_FETCH 0       // 1: staged fetch issued before the reduction instead of behind the CTA's line stores (A/B)
#endif
constexpr int kLLParts = 1024;                            // CTA values per quantity the inbox mode handles (n <= 8 Mi)
constexpr int kLLMaxCtas = 192;                           // inboxes allocated (a B200 runs 148 CTAs of these kernels)
constexpr size_t kGridBarrierBytes = sizeof(double) * 2 * kGridQ * kGridMaxParts;
constexpr size_t kLLInboxLines = (size_t)2 * kGridQ * kLLParts;
constexpr size_t kGridLineCount = kLLInboxLines * kLLMaxCtas;
constexpr unsigned kLLSeqMask = 0x0fffffffu;
inline size_t grid_part_bytes() { return kGridBarrierBytes + sizeof(uint4) * (kGridLineCount + 1); }
// host: zero everything, sequence numbers start at 1; `enabled` = 0 keeps the grid-barrier reductions (A/B)
inline cudaError_t grid_part_init(double* part, int enabled, int backoff_ns = 0, unsigned first_seq = 1u) {
    cudaError_t e = cudaMemset(part, 0, grid_part_bytes());
    if (e != cudaSuccess) return e;
    const uint4 head = make_uint4(first_seq ? (first_seq & kLLSeqMask) : 1u, 0u, (unsigned)backoff_ns, enabled ? 1u : 0u);
    return cudaMemcpy(reinterpret_cast<char*>(part) + kGridBarrierBytes + sizeof(uint4) * kGridLineCount, &head, sizeof head,
                      cudaMemcpyHostToDevice);
}
DZO_DEVINL uint4* grid_seq_cell(double* part) {
    return reinterpret_cast<uint4*>(reinterpret_cast<char*>(part) + kGridBarrierBytes) + kGridLineCount;
}
// every thread, before the kernel's first grid barrier
DZO_DEVINL void grid_ctx_begin(GridCtx& c) {
    const uint4 head = __ldcg(grid_seq_cell(c.part));
    if (head.w && 8 * c.nblocks <= kLLParts && c.nctas <= kLLMaxCtas) {
        c.ll = reinterpret_cast<uint4*>(reinterpret_cast<char*>(c.part) + kGridBarrierBytes);
        c.seq = head.x;
        c.backoff = head.z;
        if (c.seq > (kLLSeqMask >> 1)) {
            // Half of the 28-bit tag range is used up: start over at 1 with clean inboxes, so that a line left behind by a
            // reduction long ago (a quantity or parity that has not come up since) can never carry the current tag again.
            // Every CTA clears its own inbox; the kernel's first grid barrier (right behind this call) separates the
            // clearing from the first line any producer stores.  Uniform: every CTA read the same counter.
            uint4* mine = c.ll + (size_t)c.cta * kLLInboxLines;
            for (size_t i = threadIdx.x; i < kLLInboxLines; i += blockDim.x) mine[i] = make_uint4(0u, 0u, 0u, 0u);
            __threadfence();
            c.seq = 1u;
        }
    }
}
// the leader, after the kernel's last reduction (every CTA read the cell before its first one)
DZO_DEVINL void grid_ctx_end(const GridCtx& c) {
    if (c.ll != nullptr && blockIdx.x == 0 && threadIdx.x == 0) grid_seq_cell(c.part)->x = c.seq;
}
DZO_DEVINL void ll_store(uint4* p, unsigned long long bits, unsigned seq) {
    #if DZO_LL_SYS
    asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};"
#else
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};"
#endif
                 ::"l"(p), "r"((unsigned)bits), "r"(seq), "r"((unsigned)(bits >> 32)), "r"(seq)
                 : "memory");
}
DZO_DEVINL uint4 ll_load(const uint4* p) {
    uint4 q;
    #if DZO_LL_SYS
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
#else
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
#endif
                 : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    return q;
}
DZO_DEVINL bool ll_ok(const uint4& q, unsigned seq) { return (q.y & kLLSeqMask) == seq && (q.w & kLLSeqMask) == seq; }
DZO_DEVINL double ll_value(const uint4& q) { return __longlong_as_double((long long)(((unsigned long long)q.z << 32) | q.x)); }

// Reduce K per-eighth accumulators (acc[k][j] = this thread's partial of quantity k for the j-th eighth its CTA owns)
// and OR the flag words.  Returns the totals in out[k], the OR in flags; identical on every thread of the grid.
// `after_store` runs on every thread once the CTA's own lines are on their way to the other CTAs and before the CTA
// starts waiting for theirs: the place to issue memory traffic that should overlap the reduction (the staged fetch of
// the next pass) -- issued earlier it would sit in the SM's request queue AHEAD of the lines everybody else is waiting for.
template <int K, int MAXB, class F>
DZO_DEVINL void grid_reduce_then(GridCtx& c, double (&acc)[K][MAXB], unsigned (&fl)[MAXB], double (&out)[K], unsigned& flags,
                                 F&& after_store) {
#if DZO_EARLY_FETCH
    after_store();
#endif
    const int par = c.red & 1;
    c.red += 1;
#ifndef DZO_GRID_PROF
#define DZO_GRID_PROF 1
#endif
    const bool prof = DZO_GRID_PROF && (c.prof != nullptr && threadIdx.x == 0);
    long long t_in = 0, t_local = 0, t_poll = 0;
    if (prof) t_in = clock64();
    const bool ll = (c.ll != nullptr);
    const unsigned seq = c.seq;
    if (ll) c.seq = ((seq + 1u) & kLLSeqMask) ? ((seq + 1u) & kLLSeqMask) : 1u;      // 28 bits, 0 is "never written"
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < MAXB; ++j) {
        const int e = c.cta + j * c.nctas;
        if (e < 8 * c.nblocks) {             // uniform over the CTA
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const double sv = warp_butterfly_desc(acc[k][j]);                 // bits 4,3,2,1,0
                if (lane == 0) c.s_warp[k * 16 + warp] = sv;
            }
            const unsigned wf = __reduce_or_sync(0xffffffffu, fl[j]);
            if (lane == 0) c.s_wflag[warp] = wf;
            __syncthreads();
            if (warp == 0) {
                unsigned f = c.s_wflag[lane & 15];
                f = __reduce_or_sync(0xffffffffu, f);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    double v = c.s_warp[k * 16 + (lane & 15)];
#pragma unroll
                    for (int o = 1; o < 16; o <<= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);   // bits 5,6,7,8 (every lane holds v)
                    if (ll) {
                        const unsigned tag = seq | ((k == 0) ? ((f & 0xfu) << 28) : 0u);   // flag words of these kernels use bits 0..2
                        const size_t slot = (size_t)(par * kGridQ + k) * kLLParts + e;
                        for (int dst = lane; dst < c.nctas; dst += 32)             // into every CTA's inbox
                            ll_store(c.ll + (size_t)dst * kLLInboxLines + slot, (unsigned long long)__double_as_longlong(v), tag);
                    } else if (lane == 0) {
                        c.part[(par * kGridQ + k) * kGridMaxParts + e] = v;
                    }
                }
                if (!ll && lane == 0) c.fpart[par * kGridMaxParts + e] = f;
            }
            __syncthreads();                 // s_warp may be reused by the next owned eighth
        }
    }
#if !DZO_EARLY_FETCH
    after_store();
#endif
    if (prof) t_local = clock64();
    if (!ll) c.grid.sync();
    // warp 0 of every CTA fetches the CTA values side by side (one L2 round trip per 16 blocks -- a thread adding
    // them one load at a time would wait out a full L2 latency per value): lane l holds CTA values 4l .. 4l+3 of
    // the round, i.e. half a block; bits 9, 10 inside the lane, bit 11 across the lane pair, blocks in ascending order
    if (threadIdx.x < 32) {
        double tot[K];
        unsigned f = 0;
        for (int b0 = 0; b0 < c.nblocks; b0 += 16) {
            const int e0 = 8 * b0 + 4 * lane;
            const bool in = (e0 < 8 * c.nblocks);
            double half[K];
            unsigned fv = 0;
            if (ll) {
                // poll this lane's lines of the CTA's own inbox, one quantity at a time, until both tags of each carry the
                // current sequence number
                const uint4* inbox = c.ll + (size_t)c.cta * kLLInboxLines;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint4* q = inbox + ((size_t)(par * kGridQ + k) * kLLParts + e0);
                    uint4 l0 = make_uint4(0, 0, 0, 0), l1 = l0, l2 = l0, l3 = l0;
                    // The loop is left by the WHOLE warp at once (vote).  With per-lane exits the lanes dropped out of the
                    // spin loop one by one and the warp reached the combine below diverged: the same binary measured 0.20 ms
                    // per L-BFGS step! (n = 2^20) with the shuffle combine and 0.13 ms with a shared-memory combine that
                    // needs no collective; with the vote both run at 0.106 - 0.109 ms (profiles/r02_grid_lbfgs_variants_*).
                    unsigned long long t0 = 0;
                    for (unsigned spins = 1;; ++spins) {
                        bool ok = true;
                        if (in) {
                            l0 = ll_load(q + 0); l1 = ll_load(q + 1); l2 = ll_load(q + 2); l3 = ll_load(q + 3);
                            ok = ll_ok(l0, seq) & ll_ok(l1, seq) & ll_ok(l2, seq) & ll_ok(l3, seq);
                        }
                        if (__all_sync(0xffffffffu, ok)) break;
                        if (c.backoff) __nanosleep(c.backoff);
                        if ((spins & 1023u) == 0u) {                                // a grid that lost a CTA must not hang the GPU
                            const unsigned long long now = global_timer_ns();
                            if (t0 == 0) t0 = now;
                            const bool give_up = (now - t0 > 5000000000ull);
                            if (__any_sync(0xffffffffu, give_up)) { if (lane == 0) grid_seq_cell(c.part)->y = 1u; break; }
                        }
                    }
                    half[k] = (ll_value(l0) + ll_value(l1)) + (ll_value(l2) + ll_value(l3));   // bits 9, 10 (lanes past the end: 0.0)
                    if (k == 0) fv = (l0.y | l1.y | l2.y | l3.y) >> 28;
                }
                __syncwarp();
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const double* q = c.part + (par * kGridQ + k) * kGridMaxParts + e0;
                    const double c0 = in ? __ldcg(&q[0]) : 0.0, c1 = in ? __ldcg(&q[1]) : 0.0;
                    const double c2 = in ? __ldcg(&q[2]) : 0.0, c3 = in ? __ldcg(&q[3]) : 0.0;
                    half[k] = (c0 + c1) + (c2 + c3);                                  // bits 9, 10
                }
                if (in) {
                    const unsigned* fq = c.fpart + par * kGridMaxParts + e0;
                    fv = __ldcg(&fq[0]) | __ldcg(&fq[1]) | __ldcg(&fq[2]) | __ldcg(&fq[3]);
                }
            }
            f |= __reduce_or_sync(0xffffffffu, fv);
            if (prof) t_poll = clock64();
            const int cnt = min(16, c.nblocks - b0);
#pragma unroll
#if DZO_LL_COMBINE_SMEM
            for (int k = 0; k < K; ++k) {
                const double blockv = half[k] + __shfl_down_sync(0xffffffffu, half[k], 1);   // bit 11 (valid on even lanes)
                if (!(lane & 1)) c.s_warp[k * 16 + (lane >> 1)] = blockv;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k)
                for (int i = 0; i < cnt; ++i) {
                    const double x = c.s_warp[k * 16 + i];
                    tot[k] = (b0 == 0 && i == 0) ? x : tot[k] + x;
                }
            __syncwarp();
#else
            for (int k = 0; k < K; ++k) {
                const double blockv = half[k] + __shfl_down_sync(0xffffffffu, half[k], 1);   // bit 11 (valid on even lanes)
                for (int i = 0; i < cnt; ++i) {      // (through shared memory instead of 16 shuffles: measured slower, legacy
                    const double x = __shfl_sync(0xffffffffu, blockv, 2 * i);   //  L-BFGS 0.212 -> 0.230 ms per step!)
                    tot[k] = (b0 == 0 && i == 0) ? x : tot[k] + x;
                }
            }
#endif
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) c.s_out[k] = tot[k];
            *c.s_flags = f;
        }
    }
    __syncthreads();
    if (prof) {
        const long long t_out = clock64();
        c.prof[0] += t_local - t_in; c.prof[1] += t_poll - t_local; c.prof[2] += t_out - t_poll; c.prof[3] += 1;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = c.s_out[k];
    flags = *c.s_flags;
    // the next write to s_out sits behind this CTA's own contribution to the next reduction, i.e. behind the
    // __syncthreads every warp passes after reading s_out
}
template <int K, int MAXB>
DZO_DEVINL void grid_reduce(GridCtx& c, double (&acc)[K][MAXB], unsigned (&fl)[MAXB], double (&out)[K], unsigned& flags) {
    grid_reduce_then<K, MAXB>(c, acc, fl, out, flags, [] {});
}

constexpr int kGridOwnMax = 8;   // eighths one CTA may own: n <= nctas * 65536 elements (148 CTAs: 9.7 Mi)

// for (own pairs) body: j = index of the owned eighth (compile-time after unrolling, so per-eighth accumulators stay in
// registers), k = global pair index.  Eighth e: block e / 8, virtual threads 512 * (e % 8) + threadIdx.x.
#define DZO_GRID_OWN_PAIRS(c, m2, j, k)                                                                              \
    _Pragma("unroll") for (int j = 0; j < kGridOwn; ++j)                                                             \
        if ((c).cta + j * (c).nctas < 8 * (c).nblocks)                                                               \
            for (long long e__ = (c).cta + j * (c).nctas, k = (e__ >> 3) * kBlockPairs + (e__ & 7) * 512 + threadIdx.x, \
                           i__ = 0;                                                                                  \
                 i__ < 8 && k < (m2); ++i__, k += DZO_TREE_WIDTH)

// Own pairs with a compile-time position: f(j, i, k) -- j = owned block, i = position of the pair inside the thread's
// share of the block (only meaningful for OWN == 1, where the loop is fully unrolled), k = global pair index.
template <int OWN, class F>
DZO_DEVINL void own_pairs_idx(const GridCtx& c, long long m2, F&& f) {
    if constexpr (OWN == 1) {
        if (c.cta < 8 * c.nblocks) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const long long k = (long long)(c.cta >> 3) * kBlockPairs + (c.cta & 7) * 512 + threadIdx.x + (long long)DZO_TREE_WIDTH * i;
                if (k < m2) f(0, i, k);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < OWN; ++j) {
            const long long e = c.cta + j * c.nctas;
            if (e < 8 * c.nblocks) {
                long long k = (e >> 3) * kBlockPairs + (e & 7) * 512 + threadIdx.x;
                for (int i = 0; i < 8 && k < m2; ++i, k += DZO_TREE_WIDTH) f(j, 0, k);
            }
        }
    }
}

// The direction vector while compute_lbfgs_step_direction! runs: with one eighth per CTA (OWN == 1: n <= nctas * 8192,
// e.g. n = 2^20 on 128 CTAs) a thread's 8 pairs stay in REGISTERS across the 2m + 1 passes, so a pass only streams
// the history vectors; otherwise the vector lives in global memory (L2-resident).
template <int OWN>
struct DirRegs {
    double2 r[OWN == 1 ? 8 : 1];
    DZO_DEVINL double2 get(const double* d, int i, long long k) const {
        if constexpr (OWN == 1) return r[i];
        else return reinterpret_cast<const double2*>(d)[k];
    }
    DZO_DEVINL void set(double* d, int i, long long k, double2 v, bool to_global) {
        if constexpr (OWN == 1) {
            r[i] = v;
            if (to_global) reinterpret_cast<double2*>(d)[k] = v;
        } else {
            reinterpret_cast<double2*>(d)[k] = v;
        }
    }
};

// The two history vectors of the NEXT pass of compute_lbfgs_step_direction!, fetched while the reduction of the current
// pass is in flight (OWN == 1 only): which vectors the next pass reads depends on loop indices, not on the reduced value,
// so each thread copies its own 8 pairs of both with 16-byte cp.async (128 KB of shared memory per CTA, one CTA per SM)
// right before it enters the reduction and finds them in shared memory afterwards -- the pass no longer waits out an
// HBM round trip behind every reduction.  A thread only ever reads back what it copied itself: no barrier involved.
constexpr size_t kGridStageBytes = (size_t)2 * 8 * kClusterThreads * sizeof(double2);
template <int OWN>
struct PassStage {
    double2* buf;                          // [2 vectors][8 pairs][512 threads]
    bool staged;
    DZO_DEVINL void fetch(const GridCtx& c, long long m2, const double* A, const double* B) {
        if constexpr (OWN == 1) {
            if (buf != nullptr && c.cta < 8 * c.nblocks) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long k = (long long)(c.cta >> 3) * kBlockPairs + (c.cta & 7) * 512 + threadIdx.x + (long long)DZO_TREE_WIDTH * i;
                    if (k < m2) {
                        cp_async16(&buf[(0 * 8 + i) * kClusterThreads + threadIdx.x], reinterpret_cast<const double2*>(A) + k);
                        cp_async16(&buf[(1 * 8 + i) * kClusterThreads + threadIdx.x], reinterpret_cast<const double2*>(B) + k);
                    }
                }
                cp_async_commit();
                staged = true;
            }
        }
    }
    DZO_DEVINL void wait() {
        if constexpr (OWN == 1) { if (staged) cp_async_wait_all(); }
    }
    DZO_DEVINL double2 get(int v, int i, const double* G, long long k) const {
        if constexpr (OWN == 1) { if (staged) return buf[(v * 8 + i) * kClusterThreads + threadIdx.x]; }
        return reinterpret_cast<const double2*>(G)[k];
    }
    DZO_DEVINL void done() { staged = false; }
};

template <int OWN>
static __global__ void __launch_bounds__(kClusterThreads, 1) grid_lbfgs_kernel(GridLbfgsArgs a) {
    extern __shared__ __align__(16) unsigned char grid_stage_raw[];
    constexpr int kGridOwn = OWN;          // the macros and arrays below size themselves by it
    __shared__ LbfgsCtrl sc;
    __shared__ double alpha[DZO_LBFGS_MAX_HISTORY];
    __shared__ double s_warp[kGridQ * 16];
    __shared__ unsigned s_wflag[16];
    __shared__ double s_out[kGridQ];
    __shared__ unsigned s_flags;
    GridCtx c{cg::this_grid(), (int)gridDim.x, (int)blockIdx.x, a.nblocks, 0, a.part, a.fpart, s_warp, s_wflag, s_out, &s_flags};
    grid_ctx_begin(c);
    __shared__ long long s_prof[4];
    long long t_kernel = 0;
    if ((a.stage & 2) && blockIdx.x == 0) {                    // measurement knob "grid_profile": cycle split of the leader CTA
        if (threadIdx.x == 0) { s_prof[0] = s_prof[1] = s_prof[2] = s_prof[3] = 0; t_kernel = clock64(); }
        c.prof = s_prof;
    }
    const long long n = a.n, m2 = n >> 1;
    const bool leader = (blockIdx.x == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0) sc = *a.ctrl;
    __syncthreads();
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
        grid_reduce<2, kGridOwn>(c, acc, fl, out, f);
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
        grid_ctx_end(c);
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
            DirRegs<OWN> D;
            PassStage<OWN> St{(OWN == 1 && (a.stage & 1)) ? reinterpret_cast<double2*>(grid_stage_raw) : nullptr, false};
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
            {
                const double* s0 = a.S + (long long)head * n;
                own_pairs_idx<OWN>(c, m2, [&](int j, int i, long long k) {                      // d = g, and s_0 . d  (:439)
                    const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                    const double2 ss = reinterpret_cast<const double2*>(s0)[k];
                    D.set(a.d, i, k, gg, false);
                    acc[0][j] += ss.x * gg.x;
                    acc[0][j] += ss.y * gg.y;
                });
            }
            // vectors of pass i of the first loop: y_p and (s_{p+1} | y_p of the last entry again for y_{cnt-1} . d)
            auto first_loop_vectors = [&](int i, const double*& y, const double*& nxt) {
                const int p = (head + i) % m;
                const bool last = (i == cnt - 1);
                const int pn = last ? p : (head + i + 1) % m;
                y = a.Y + (long long)p * n;
                nxt = last ? (a.Y + (long long)pn * n) : (a.S + (long long)pn * n);
            };
            // vectors of pass i of the second loop: s_p and (y_{p-1} | x for the fused first trial)
            auto second_loop_vectors = [&](int i, const double*& sp, const double*& other) {
                sp = a.S + (long long)((head + i) % m) * n;
                other = (i > 0) ? (a.Y + (long long)((head + i - 1) % m) * n) : a.x;
            };
            grid_reduce_then<1, kGridOwn>(c, acc, fl, out, f, [&] {
                const double *fy, *fn;
                first_loop_vectors(0, fy, fn);
                St.fetch(c, m2, fy, fn);
            });
            double al = out[0] / sc.rho[head];
            const double cc = -sc.rho[head] / sc.yy;                                            // :443 (dot(y_0, y_0) cached)
            double beta = 0.0;
            for (int i = 0; i < cnt; ++i) {
                const int p = (head + i) % m;
                if (threadIdx.x == 0) alpha[i] = al;
                const double na = -al;
                const double *y, *nxt;
                first_loop_vectors(i, y, nxt);
                const bool last = (i == cnt - 1);
                const int pn = last ? p : (head + i + 1) % m;
#pragma unroll
                for (int j = 0; j < kGridOwn; ++j) acc[0][j] = 0.0;
                St.wait();
                own_pairs_idx<OWN>(c, m2, [&](int j, int q, long long k) {
                    double2 dd = D.get(a.d, q, k);
                    const double2 yy = St.get(0, q, y, k);
                    const double2 nn = St.get(1, q, nxt, k);
                    dd.x += na * yy.x; dd.y += na * yy.y;                                       // :440
                    if (last) { dd.x *= cc; dd.y *= cc; }                                       // :443
                    D.set(a.d, q, k, dd, false);
                    acc[0][j] += nn.x * dd.x;                                                   // s_{i+1} . d (:439) or y_{cnt-1} . d (:446)
                    acc[0][j] += nn.y * dd.y;
                });
                St.done();
                grid_reduce_then<1, kGridOwn>(c, acc, fl, out, f, [&] {
                    const double *fa, *fb;
                    if (!last) first_loop_vectors(i + 1, fa, fb);
                    else second_loop_vectors(cnt - 1, fa, fb);
                    St.fetch(c, m2, fa, fb);
                });
                if (last) beta = out[0] / sc.rho[p];
                else al = out[0] / sc.rho[pn];
            }
            __syncthreads();   // alpha[] visible
            for (int i = cnt - 1; i >= 0; --i) {
                const double na = -(alpha[i] + beta);
                const double *sp, *other;
                second_loop_vectors(i, sp, other);
#pragma unroll
                for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; fl[j] = 0; }
                St.wait();
                if (i > 0) {
                    const int pn = (head + i - 1) % m;
                    own_pairs_idx<OWN>(c, m2, [&](int j, int q, long long k) {
                        double2 dd = D.get(a.d, q, k);
                        const double2 ss = St.get(0, q, sp, k);
                        const double2 yy = St.get(1, q, other, k);
                        dd.x += na * ss.x; dd.y += na * ss.y;                                   // :447
                        D.set(a.d, q, k, dd, false);
                        acc[0][j] += yy.x * dd.x;                                               // y_{i-1} . d  (:446)
                        acc[0][j] += yy.y * dd.y;
                    });
                    St.done();
                    grid_reduce_then<1, kGridOwn>(c, acc, fl, out, f, [&] {
                        const double *fa, *fb;
                        second_loop_vectors(i - 1, fa, fb);
                        St.fetch(c, m2, fa, fb);
                    });
                    beta = out[0] / sc.rho[pn];
                } else {
                    // last correction fused with the first trial of take_backtracking_step!(opt, 1, d)  (:124-138)
                    own_pairs_idx<OWN>(c, m2, [&](int j, int q, long long k) {
                        double2 dd = D.get(a.d, q, k);
                        const double2 ss = St.get(0, q, sp, k);
                        const double2 xx = St.get(1, q, other, k);
                        dd.x += na * ss.x; dd.y += na * ss.y;                                   // :447
                        D.set(a.d, q, k, dd, true);                                             // step_direction as the host reads it
                        const double w0 = xx.x + step * dd.x, w1 = xx.y + step * dd.y;         // :124 axpy!
                        if (!julia_isequal(w0, xx.x) || !julia_isequal(w1, xx.y)) fl[j] |= 1u;  // :128
                        acc[0][j] += RosenbrockVec::term(w0, w1);                               // :138
                    });
                    St.done();
                    grid_reduce<1, kGridOwn>(c, acc, fl, pr_out, pr_flags);
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
                grid_reduce<1, kGridOwn>(c, acc, fl, pr_out, pr_flags);
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
        grid_reduce<2, kGridOwn>(c, acc, fl, out, f);
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
    grid_ctx_end(c);
    if (c.prof != nullptr && threadIdx.x == 0)
        printf("grid_lbfgs profile (CTA 0, cycles): kernel %lld, %lld reductions: CTA-level %lld, waiting for the grid %lld, final combine %lld\n",
               clock64() - t_kernel, s_prof[3], s_prof[0], s_prof[1], s_prof[2]);
}

}  // namespace dzo

// ============================================================================= live AdGDOptimizer (:179-312), n > DZO_TREE_BLOCK
// The same control flow as cluster_adgd_kernel (lbfgs_kernels.cuh) on the cooperative grid / blocked tree.
namespace dzo {
struct GridAdgdArgs {
    AdgdArgs a;
    double* part;
    unsigned* fpart;
    int nblocks;
};
template <int INSTANCE>   // compiled in grid_lbfgs_tu.cu only
static __global__ void __launch_bounds__(kClusterThreads, 1) grid_adgd_kernel(GridAdgdArgs ga) {
    constexpr int kGridOwn = kGridOwnMax;
    const AdgdArgs& a = ga.a;
    __shared__ AdgdCtrl sc;
    __shared__ double s_warp[kGridQ * 16];
    __shared__ unsigned s_wflag[16];
    __shared__ double s_out[kGridQ];
    __shared__ unsigned s_flags;
    GridCtx c{cg::this_grid(), (int)gridDim.x, (int)blockIdx.x, ga.nblocks, 0, ga.part, ga.fpart, s_warp, s_wflag, s_out, &s_flags};
    grid_ctx_begin(c);
    const long long m2 = a.n >> 1;
    const bool leader = (blockIdx.x == 0 && threadIdx.x == 0);
    if (threadIdx.x == 0 && a.mode == 0) sc = *a.ctrl;
    __syncthreads();
    c.grid.sync();
    double acc[2][kGridOwn];
    unsigned fl[kGridOwn];
    double out[2];
    unsigned f;
    if (a.mode == 1) {                                                                          // :201-271
#pragma unroll
        for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; fl[j] = 0; }
        DZO_GRID_OWN_PAIRS(c, m2, j, k) {
            const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
            const double2 gg = RosenbrockVec::grad(xx.x, xx.y);                                 // :265
            reinterpret_cast<double2*>(a.g)[k] = gg;
            reinterpret_cast<double2*>(a.dx)[k] = make_double2(0.0, 0.0);
            reinterpret_cast<double2*>(a.dg)[k] = make_double2(0.0, 0.0);
            acc[0][j] += RosenbrockVec::term(xx.x, xx.y);                                       // :260
            acc[1][j] += gg.x * gg.x; acc[1][j] += gg.y * gg.y;
        }
        grid_reduce<2, kGridOwn>(c, acc, fl, out, f);
        const double gnorm = sqrt(out[1]);                                                      // :233
        if (leader) {
            AdgdCtrl t;
            t.f = out[0]; t.df = 0.0; t.iter = 0; t.pad = 0;
            t.stuck = (gnorm == 0.0);                                                           // :234
            t.cur = t.prev = t.stuck ? 0.0 : a.initial_step_length / gnorm;                     // :235-236
            *a.ctrl = t;
        }
        grid_ctx_end(c);
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
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; fl[j] = 0; }
            DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                const double2 dgv = reinterpret_cast<const double2*>(a.dg)[k];
                const double2 dxv = reinterpret_cast<const double2*>(a.dx)[k];
                acc[0][j] += dgv.x * dgv.x; acc[0][j] += dgv.y * dgv.y;
                acc[1][j] += dxv.x * dxv.x; acc[1][j] += dxv.y * dxv.y;
            }
            grid_reduce<2, kGridOwn>(c, acc, fl, out, f);
            const double dgn = sqrt(out[0]);
            if (dgn != 0.0) {
                const double inv_L = sqrt(out[1]) / dgn;
                next = julia_min(next, inv_sqrt_two * inv_L);
            }
        }
        // take_backtracking_step!(opt, -next, current_gradient)  :301, :107-154
        double step = -next, nxt = 0.0;
        bool accepted = false;
        for (;;) {
#pragma unroll
            for (int j = 0; j < kGridOwn; ++j) { acc[0][j] = 0.0; acc[1][j] = 0.0; fl[j] = 0; }
            DZO_GRID_OWN_PAIRS(c, m2, j, k) {
                const double2 xx = reinterpret_cast<const double2*>(a.x)[k];
                const double2 gg = reinterpret_cast<const double2*>(a.g)[k];
                const double w0 = xx.x + step * gg.x, w1 = xx.y + step * gg.y;
                if (!julia_isequal(w0, xx.x) || !julia_isequal(w1, xx.y)) fl[j] |= 1u;
                acc[0][j] += RosenbrockVec::term(w0, w1);
            }
            grid_reduce<2, kGridOwn>(c, acc, fl, out, f);
            if (!(f & 1u)) break;
            nxt = out[0];
            if (nxt < sc.f) { accepted = true; break; }
            step *= 0.5;
        }
        if (!accepted) {
            DZO_GRID_OWN_PAIRS(c, m2, j, k) reinterpret_cast<double2*>(a.dx)[k] = reinterpret_cast<const double2*>(a.x)[k];
            if (threadIdx.x == 0) { sc.stuck = 1; sc.prev = current; sc.cur = next; }
            __syncthreads();
            break;
        }
        DZO_GRID_OWN_PAIRS(c, m2, j, k) {
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
        __syncthreads();
        if (threadIdx.x == 0) {
            sc.df = nxt - sc.f; sc.f = nxt; sc.prev = current; sc.cur = next; sc.iter += 1;     // :298-299, :310
        }
        __syncthreads();
    }
    if (leader) *a.ctrl = sc;
    grid_ctx_end(c);
}
}  // namespace dzo
