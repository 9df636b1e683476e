// common.cuh -- shared device/host helpers of libdzopt_b200 (sm_100a only).
//
// Arithmetic policy: the legacy Julia on this path contains no muladd and no
// @simd reductions, so nothing may be contracted into FMA.  The whole library is compiled
// with -fmad=false; double-precision '/' and sqrt() are IEEE-correct in CUDA regardless of
// flags.  Never use rsqrt()/__drcp_* here; the only approximate instructions in the library are inside
// ieee_fast.cuh, which replicates the compiler's own IEEE sqrt / division fast paths and is self-tested on the device.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/dzopt.h"

#define DZO_DEVINL __device__ __forceinline__

namespace dzo {

// ----------------------------------------------------------------------------- control block
// Per-problem scalars of the large-n path that kernels hand to one another on the device
// (no host round trip inside step!).
struct LargeCtrl {
    double f;          // current_objective_value      legacy/DZOptimization.jl:740
    double L;          // last_step_length             :744
    long long iter;    // iteration_count              :737
    int type;          // last_step_type               :745
    int term;          // has_terminated               :738
    // hand-off between the kernels of one step!
    int kind;          // what THIS step did: DZO_STEP_NULL (nothing / terminated), _BFGS, _GRADIENT_DESCENT
    int pad;           // row-sharded mode: set to 1 when a wait for a peer's slab timed out
    double step_length;// -alpha handed to update_inverse_hessian! (:954)
    double overlap;    // :873
    double delta_norm; // :876
    // statistics
    long long evals;   // objective evaluations so far
    long long calls;   // step! calls so far on a non-terminated optimizer
    unsigned char kind_log[64];  // kind of call c at kind_log[c % 64] (lets a host time back-to-back steps
                                 // without a sync per step and still know which ones were BFGS-type)
};

// ----------------------------------------------------------------------------- warp butterflies
// Canonical tree (include/dzopt.h): lane bits are combined 16,8,4,2,1; every level
// above the warp is combined in ASCENDING bit order.
DZO_DEVINL double warp_butterfly_desc(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
DZO_DEVINL double warp_butterfly_asc(double v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce K independent quantities over the canonical tree inside ONE CTA of 1024 threads.
// Thread tid emulates the 4 virtual threads v = tid + 1024*q; p[k][q] is the partial of
// quantity k on virtual thread v.  `sm` needs K*132 doubles.  Result (identical on every
// thread) is written to out[k].  Contains 3 __syncthreads().
template <int K>
DZO_DEVINL void cta1024_tree_reduce(double (&p)[K][4], double* sm, double (&out)[K]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double s = warp_butterfly_desc(p[k][q]);           // bits 4,3,2,1,0
            if (lane == 0) sm[k * 132 + q * 32 + warp] = s;
        }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = warp_butterfly_asc(sm[k * 132 + q * 32 + lane]); // bits 5..9
            const double a = r[0] + r[1];                            // bit 10
            const double b = r[2] + r[3];
            if (lane == 0) sm[k * 132 + 128] = a + b;                // bit 11
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = sm[k * 132 + 128];
    __syncthreads();  // sm may be reused immediately by the caller
}

// The same tree inside ONE CTA of NT = 512 or 1024 threads: thread tid emulates the VT = 4096 / NT virtual threads
// v = tid + NT*q, so lane bits are tree bits 4..0, the warp index bits 5.. and q the top bits.  p[k][q] is the partial of
// quantity k on virtual thread v; `sm` needs K*132 doubles.  NT = 1024 is cta1024_tree_reduce, bit for bit.
template <int K, int NT>
DZO_DEVINL void cta_tree_reduce(double (&p)[K][4096 / NT], double* sm, double (&out)[K]) {
    constexpr int VT = 4096 / NT, W = NT / 32;
    static_assert(NT == 512 || NT == 1024, "canonical tree on one CTA: 512 or 1024 threads");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int q = 0; q < VT; ++q) {
            const double s = warp_butterfly_desc(p[k][q]);           // bits 4,3,2,1,0
            if (lane == 0) sm[k * 132 + q * W + warp] = s;
        }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double r[VT];
#pragma unroll
            for (int q = 0; q < VT; ++q) {
                double v = sm[k * 132 + q * W + (lane & (W - 1))];
#pragma unroll
                for (int o = 1; o < W; o <<= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);     // warp-index bits, ascending
                r[q] = v;
            }
#pragma unroll
            for (int w = 1; w < VT; w <<= 1)                         // q bits, ascending
#pragma unroll
                for (int i = 0; i < VT; i += 2 * w) r[i] = r[i] + r[i + w];
            if (lane == 0) sm[k * 132 + 128] = r[0];
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) out[k] = sm[k * 132 + 128];
    __syncthreads();  // sm may be reused immediately by the caller
}

// ----------------------------------------------------------------------------- PTX: mbarrier + bulk async copy (TMA 1-D)
DZO_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DZO_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DZO_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
DZO_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DZO_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DZO_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy, completion signalled on an mbarrier (SASS: UBLKCP)
DZO_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk-group completion)
DZO_DEVINL void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
DZO_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
DZO_DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
DZO_DEVINL void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
DZO_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 16-byte asynchronous global -> shared copy (SASS: LDGSTS), L2 only; both addresses 16-byte aligned
DZO_DEVINL void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
DZO_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DZO_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// streaming (evict-first) 128-bit global access for the n^2 sweeps: H has no reuse inside a sweep
DZO_DEVINL double2 ldg_stream2(const double* p) {
    double2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
DZO_DEVINL void stg_stream2(double* p, double2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

DZO_DEVINL bool finite_(double v) { return isfinite(v); }

// ----------------------------------------------------------------------------- peer memory (row-sharded mode)
// One process per GPU; every rank maps the result vectors and the flag words of all its peers
// (CUDA IPC over NVLink / NVSwitch).  A producer kernel stores its rows of the result straight into
// every peer's copy and then raises flag[producer rank] = sequence number in every peer; a consumer
// kernel spins on its LOCAL flag words.  This replaces "kernel -> ncclAllGather -> kernel".
constexpr int kMaxPeers = 8;
struct PeerSet {
    int nranks;                               // 1 = not sharded (or NCCL fallback): no peer traffic
    int rank;
    double* out[kMaxPeers];                   // out[p] = rank p's copy of the vector being produced
    unsigned long long* flags[kMaxPeers];     // flags[p] = rank p's flag array (kMaxPeers words)
    unsigned* done;                           // local: row blocks finished in this launch
};
DZO_DEVINL void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
DZO_DEVINL unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// consumer side: wait until every rank's slab of sequence `seq` has landed in local memory.  A peer that
// stops stepping (host error, killed process) must not hang this GPU for ever: after kPeerTimeoutNs the
// wait gives up and records the fact in *timeout_flag, which the host turns into DZO_ERR_NCCL at the next sync.
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;
DZO_DEVINL unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Lane p < nranks of warp 0 watches rank p's flag; the warp leaves the loop as a whole (vote), so it reaches the shuffles
// of the reductions that follow converged (grid_lbfgs.cuh: per-lane exits from a poll loop cost 3 us per reduction there).
DZO_DEVINL void peer_wait(const unsigned long long* local_flags, int nranks, unsigned long long seq, int* timeout_flag) {
    if (nranks > 1 && threadIdx.x < 32) {
        const unsigned long long t0 = global_timer_ns();
        for (;;) {
            const bool ok = ((int)threadIdx.x >= nranks) || (ld_acquire_sys(local_flags + threadIdx.x) >= seq);
            if (__all_sync(0xffffffffu, ok)) break;
            __nanosleep(64);
            const bool give_up = (global_timer_ns() - t0 > kPeerTimeoutNs);
            if (__any_sync(0xffffffffu, give_up)) {
                if (threadIdx.x == 0) atomicExch(timeout_flag, 1);
                break;
            }
        }
    }
    __syncthreads();
}

}  // namespace dzo
