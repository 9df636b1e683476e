// pairwise.cuh -- sm_100a versions of the accelerated pairwise radial kernels of the LIVE package
// (src/ExampleFunctions.jl:117-149 energy, :224-262 gradient, :367-424 Hessian-vector product) with the
// Lennard-Jones radial functions (:16-72).  Structure-of-arrays x, y, z [, u, v, w]; FP64.
//
// FMA policy: the reference writes muladd explicitly inside the LJ functions -> fma() here; everything
// else stays un-contracted (the library is compiled with -fmad=false).
//
// DZO_ORDER_SEQUENTIAL  one thread per particle i walking j = 1..n serially -- the reference kernel's own
//                       order -- with the j particles staged through shared memory in tiles of 256
//                       (the reference re-reads x, y, z from global memory for every i).
// DZO_ORDER_TREE        a CTA = 32 particles x 8 warps; warp w takes segments w, w+8, ... of 128 sources
//                       (sequential inside a segment), and after every round the 8 segment partials are
//                       added to the row accumulator in ascending segment order.  8x the parallelism of
//                       SEQUENTIAL with a fixed, launch-independent rounding; no workspace.
#pragma once
#include "common.cuh"

#include "ieee_fast.cuh"

namespace dzo {

struct LennardJones {
    // every radial function starts with inv_r2 = 1 / r2; the *_inv forms take it as given so that a batch of pair terms
    // can obtain its reciprocals through the interleavable IEEE fast path (ieee_fast.cuh) -- same operations, same bits
    static DZO_DEVINL double energy_inv(double inv_r2) {     // lj_energy  :16-27
        const double inv_r4 = inv_r2 * inv_r2;
        const double inv_r6 = inv_r4 * inv_r2;
        return 4.0 * fma(inv_r6, inv_r6, -inv_r6);
    }
    static DZO_DEVINL double first_inv(double inv_r2) {      // lj_first_derivative  :30-47
        const double inv_r4 = inv_r2 * inv_r2;
        const double inv_r6 = inv_r4 * inv_r2;
        const double inv_r8 = inv_r4 * inv_r4;
        return -12.0 * fma(inv_r8, inv_r6 + inv_r6, -inv_r8);
    }
    static DZO_DEVINL double second_inv(double inv_r2) {     // lj_second_derivative  :50-72
        const double inv_r4 = inv_r2 * inv_r2;
        const double inv_r8 = inv_r4 * inv_r4;
        const double inv_r10 = inv_r8 * inv_r2;
        return 48.0 * fma(3.5, inv_r8 * inv_r8, -inv_r10);
    }
    static DZO_DEVINL double energy(double r2) { return energy_inv(1.0 / r2); }
    static DZO_DEVINL double first(double r2) { return first_inv(1.0 / r2); }
    static DZO_DEVINL double second(double r2) { return second_inv(1.0 / r2); }
};

struct PairwiseArgs {
    long long n;
    const double *x, *y, *z, *u, *v, *w;
    double *o0, *o1, *o2;   // energy: o0 = point_energies; gradient: gx,gy,gz; hvp: px,py,pz
};

// WHAT: 0 energy, 1 gradient, 2 Hessian-vector product.  One interaction of particle i with source j.
template <int WHAT, class Pot>
DZO_DEVINL void pair_term(bool self, double xi, double yi, double zi, double ui, double vi, double wi, double xj,
                          double yj, double zj, double uj, double vj, double wj, double& ax, double& ay, double& az) {
    const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
    const double r2 = dx * dx + dy * dy + dz * dz;
    if (WHAT == 0) {
        ax += self ? 0.0 : Pot::energy(r2);                                  // :145
    } else if (WHAT == 1) {
        const double f = self ? 0.0 : Pot::first(r2);                        // :253
        ax += f * dx; ay += f * dy; az += f * dz;                            // :254-256
    } else {
        const double du = ui - uj, dv = vi - vj, dw = wi - wj;
        const double f = self ? 0.0 : Pot::first(r2);                        // :409
        const double s = self ? 0.0 : Pot::second(r2);                       // :410
        const double overlap = dx * du + dy * dv + dz * dw;                  // :411
        const double os = overlap * s;
        const double g = os + os;                                            // :413 twice(overlap * s)
        ax += f * du + g * dx; ay += f * dv + g * dy; az += f * dw + g * dz; // :416-418
    }
}

// Four consecutive sources at once: the four reciprocals run interleaved through ieee_fast_rcp (the compiler's own
// 1.0 / r2 puts a slow-path branch between neighbouring terms, which serialises their FP64 chains); the accumulation
// stays in source order.  A batch with an r2 outside [2^-500, 2^500) falls back to the operator.
template <int WHAT, class Pot>
DZO_DEVINL void pair_terms4(long long j0, long long i, double xi, double yi, double zi, double ui, double vi, double wi,
                            const double* sx, const double* sy, const double* sz, const double* su, const double* sv,
                            const double* sw, double& ax, double& ay, double& az) {
    double dx[4], dy[4], dz[4], rr[4], inv[4];
    bool safe = true;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        dx[u] = xi - sx[u]; dy[u] = yi - sy[u]; dz[u] = zi - sz[u];
        const double r2 = dx[u] * dx[u] + dy[u] * dy[u] + dz[u] * dz[u];
        rr[u] = (j0 + u == i) ? 1.0 : r2;                                     // the self term is skipped below
        safe &= ieee_fast_safe(rr[u]);
    }
    if (safe) {
#pragma unroll
        for (int u = 0; u < 4; ++u) inv[u] = ieee_fast_rcp(rr[u]);
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) inv[u] = 1.0 / rr[u];
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const bool self = (j0 + u == i);
        if (WHAT == 0) {
            ax += self ? 0.0 : Pot::energy_inv(inv[u]);                       // :145
        } else if (WHAT == 1) {
            const double f = self ? 0.0 : Pot::first_inv(inv[u]);             // :253
            ax += f * dx[u]; ay += f * dy[u]; az += f * dz[u];                // :254-256
        } else {
            const double du = ui - su[u], dv = vi - sv[u], dw = wi - sw[u];
            const double f = self ? 0.0 : Pot::first_inv(inv[u]);             // :409
            const double sd = self ? 0.0 : Pot::second_inv(inv[u]);           // :410
            const double overlap = dx[u] * du + dy[u] * dv + dz[u] * dw;      // :411
            const double os = overlap * sd;
            const double g = os + os;                                         // :413
            ax += f * du + g * dx[u]; ay += f * dv + g * dy[u]; az += f * dw + g * dz[u];   // :416-418
        }
    }
}

template <int WHAT>
DZO_DEVINL void pair_store(const PairwiseArgs& a, long long i, double ax, double ay, double az) {
    if (WHAT == 0) {
        a.o0[i] = 0.5 * ax;                                                  // :148
    } else {
        a.o0[i] = ax + ax; a.o1[i] = ay + ay; a.o2[i] = az + az;             // :258-260, :419-421
    }
}

constexpr int kPairSeqThreads = 256;   // the reference's default workgroupsize (:157, :273, :439)

template <int WHAT, class Pot>
static __global__ void __launch_bounds__(kPairSeqThreads) pairwise_seq_kernel(PairwiseArgs a) {
    __shared__ double sx[kPairSeqThreads], sy[kPairSeqThreads], sz[kPairSeqThreads];
    __shared__ double su[WHAT == 2 ? kPairSeqThreads : 1], sv[WHAT == 2 ? kPairSeqThreads : 1], sw[WHAT == 2 ? kPairSeqThreads : 1];
    const long long i = (long long)blockIdx.x * kPairSeqThreads + threadIdx.x;
    const bool valid = i < a.n;
    const double xi = valid ? a.x[i] : 0.0, yi = valid ? a.y[i] : 0.0, zi = valid ? a.z[i] : 0.0;
    double ui = 0.0, vi = 0.0, wi = 0.0;
    if (WHAT == 2 && valid) { ui = a.u[i]; vi = a.v[i]; wi = a.w[i]; }
    double ax = 0.0, ay = 0.0, az = 0.0;
    for (long long t0 = 0; t0 < a.n; t0 += kPairSeqThreads) {
        const long long j = t0 + threadIdx.x;
        if (j < a.n) {
            sx[threadIdx.x] = a.x[j]; sy[threadIdx.x] = a.y[j]; sz[threadIdx.x] = a.z[j];
            if (WHAT == 2) { su[threadIdx.x] = a.u[j]; sv[threadIdx.x] = a.v[j]; sw[threadIdx.x] = a.w[j]; }
        }
        __syncthreads();
        const int cnt = (int)((a.n - t0 < kPairSeqThreads) ? (a.n - t0) : kPairSeqThreads);
        if (valid) {
            int jj = 0;
            for (; jj + 4 <= cnt; jj += 4)
                pair_terms4<WHAT, Pot>(t0 + jj, i, xi, yi, zi, ui, vi, wi, sx + jj, sy + jj, sz + jj, su + (WHAT == 2 ? jj : 0),
                                       sv + (WHAT == 2 ? jj : 0), sw + (WHAT == 2 ? jj : 0), ax, ay, az);
            for (; jj < cnt; ++jj)
                pair_term<WHAT, Pot>(t0 + jj == i, xi, yi, zi, ui, vi, wi, sx[jj], sy[jj], sz[jj],
                                     WHAT == 2 ? su[jj] : 0.0, WHAT == 2 ? sv[jj] : 0.0, WHAT == 2 ? sw[jj] : 0.0, ax, ay, az);
        }
        __syncthreads();
    }
    if (valid) pair_store<WHAT>(a, i, ax, ay, az);
}

// warps per CTA = segments in flight per row block (8; 4 for the Hessian-vector product, whose staged
// segment is twice as large -- the static shared-memory limit is 48 KB)
template <int WHAT> struct PairTreeWarps { static constexpr int value = (WHAT == 2) ? 4 : 8; };

template <int WHAT, class Pot>
static __global__ void __launch_bounds__(32 * PairTreeWarps<WHAT>::value) pairwise_tree_kernel(PairwiseArgs a) {
    constexpr int kPairTreeWarps = PairTreeWarps<WHAT>::value;
    constexpr int NV = (WHAT == 2) ? 6 : 3;
    __shared__ double src[kPairTreeWarps][NV][DZO_RIESZ_SEG];   // per-warp staged segment
    __shared__ double part[kPairTreeWarps][3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * 32 + lane;
    const bool valid = i < a.n;
    const double xi = valid ? a.x[i] : 0.0, yi = valid ? a.y[i] : 0.0, zi = valid ? a.z[i] : 0.0;
    double ui = 0.0, vi = 0.0, wi = 0.0;
    if (WHAT == 2 && valid) { ui = a.u[i]; vi = a.v[i]; wi = a.w[i]; }
    const long long nseg = (a.n + DZO_RIESZ_SEG - 1) / DZO_RIESZ_SEG;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;   // row accumulators, kept by warp 0
    for (long long s_base = 0; s_base < nseg; s_base += kPairTreeWarps) {
        const long long s = s_base + warp;
        double ax = 0.0, ay = 0.0, az = 0.0;
        if (s < nseg) {
            const long long j0 = s * DZO_RIESZ_SEG;
            const int cnt = (int)((a.n - j0 < DZO_RIESZ_SEG) ? (a.n - j0) : DZO_RIESZ_SEG);
            for (int q = lane; q < cnt; q += 32) {
                src[warp][0][q] = a.x[j0 + q]; src[warp][1][q] = a.y[j0 + q]; src[warp][2][q] = a.z[j0 + q];
                if (WHAT == 2) { src[warp][3][q] = a.u[j0 + q]; src[warp][4][q] = a.v[j0 + q]; src[warp][5][q] = a.w[j0 + q]; }
            }
            __syncwarp();
            if (valid) {
                int jj = 0;
                for (; jj + 4 <= cnt; jj += 4)
                    pair_terms4<WHAT, Pot>(j0 + jj, i, xi, yi, zi, ui, vi, wi, &src[warp][0][jj], &src[warp][1][jj],
                                           &src[warp][2][jj], &src[warp][WHAT == 2 ? 3 : 0][jj], &src[warp][WHAT == 2 ? 4 : 0][jj],
                                           &src[warp][WHAT == 2 ? 5 : 0][jj], ax, ay, az);
                for (; jj < cnt; ++jj)
                    pair_term<WHAT, Pot>(j0 + jj == i, xi, yi, zi, ui, vi, wi, src[warp][0][jj], src[warp][1][jj],
                                         src[warp][2][jj], WHAT == 2 ? src[warp][3][jj] : 0.0,
                                         WHAT == 2 ? src[warp][4][jj] : 0.0, WHAT == 2 ? src[warp][5][jj] : 0.0, ax, ay, az);
            }
        }
        part[warp][0][lane] = ax; part[warp][1][lane] = ay; part[warp][2][lane] = az;
        __syncthreads();
        if (warp == 0) {
            // ascending segment order, starting from partial 0 (oracle pairwise_item)
#pragma unroll
            for (int k = 0; k < kPairTreeWarps; ++k) {
                if (s_base + k >= nseg) break;
                if (s_base + k == 0) { acc0 = part[k][0][lane]; acc1 = part[k][1][lane]; acc2 = part[k][2][lane]; }
                else { acc0 += part[k][0][lane]; acc1 += part[k][1][lane]; acc2 += part[k][2][lane]; }
            }
        }
        __syncthreads();
    }
    if (warp == 0 && valid) pair_store<WHAT>(a, i, acc0, acc1, acc2);
}

// sum(point_energies)  :172 through the canonical tree (point i -> virtual thread i mod 4096)
static __global__ void __launch_bounds__(1024, 1) pairwise_energy_sum_kernel(const double* e, long long n, double* out) {
    __shared__ double sm[132];
    double p[1][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double acc = 0.0;
        for (long long i = threadIdx.x + 1024 * q; i < n; i += DZO_TREE_WIDTH) acc += e[i];
        p[0][q] = acc;
    }
    double o[1];
    cta1024_tree_reduce<1>(p, sm, o);
    if (threadIdx.x == 0) *out = o[0];
}

}  // namespace dzo
