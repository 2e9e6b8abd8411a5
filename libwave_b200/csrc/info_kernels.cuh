// Lu & Milios information matrix on the device: ICPMatcher::estimateLUM / estimateLUMold
// (wave_matching/src/icp_pcl_functions.cpp:182-289 / 51-179).  Per pair the reference forms, in
// fp32, aver = 0.5f * (a + b) and diff = a - b from the aligned source point a and its target match
// b, then accumulates 12 entries of M'M and 6 of M'Z (double accumulators), solves D = MM^-1 MZ and
// sums the fp32 residual terms into `ss`.  Here the same fp32 terms are summed exactly in 128-bit
// fixed point (DESIGN.md "Estimator arithmetic"), so the result does not depend on summation order.
#pragma once
#include "common.cuh"
#include "icp_kernels.cuh"
#include "index.cuh"

namespace wavecu {

struct LumArgs {
    const float4 *cur;        // Morton-ordered working source (w = original index)
    const float4 *raw;        // source in original order
    int n_src;
    const float4 *tgt;        // Morton-ordered target
    const int *pos;           // per sorted source point: matched target position or -1
    const IcpState *st;       // final transform
    Acc128 *acc;              // kAccSlots x kMaxAcc, zeroed
    double scale;             // 2^k of the 15 M'M / M'Z sums
    double scale_ss;          // 2^k of the residual sum
    double D[6];
};

__device__ __forceinline__ void lum_pair(const LumArgs &a, int s, int pos, float av[3], float df[3]) {
    const int orig = __float_as_int(a.cur[s].w);
    const float4 p = a.raw[orig];
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = a.st->T_final[k];
    const float ax = xform_row(T + 0, p.x, p.y, p.z), ay = xform_row(T + 4, p.x, p.y, p.z),
                az = xform_row(T + 8, p.x, p.y, p.z);
    const float4 b = __ldg(a.tgt + pos);
    av[0] = __fmul_rn(0.5f, __fadd_rn(ax, b.x));
    av[1] = __fmul_rn(0.5f, __fadd_rn(ay, b.y));
    av[2] = __fmul_rn(0.5f, __fadd_rn(az, b.z));
    df[0] = __fsub_rn(ax, b.x);
    df[1] = __fsub_rn(ay, b.y);
    df[2] = __fsub_rn(az, b.z);
}

// PASS 0: the 15 sums + pair count.  PASS 1: ss given D.
template <int PASS>
__global__ void __launch_bounds__(kReduceThreads) lum_kernel(LumArgs a) {
    constexpr int NV = 16;
    __shared__ long long s_part[kReduceWarps][NV + 1];
    long long v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0;
    int count = 0;
    const int base = blockIdx.x * (kReduceThreads * kReducePerThread) + threadIdx.x;
    for (int r = 0; r < kReducePerThread; ++r) {
        const int s = base + r * kReduceThreads;
        if (s >= a.n_src) break;
        const int pos = a.pos[s];
        if (pos < 0) continue;
        float av[3], df[3];
        lum_pair(a, s, pos, av, df);
        ++count;
        if (PASS == 0) {
            const double sc = a.scale;
            v[0] += __double2ll_rn((double) av[0] * sc);
            v[1] += __double2ll_rn((double) av[1] * sc);
            v[2] += __double2ll_rn((double) av[2] * sc);
            v[3] += __double2ll_rn((double) __fmul_rn(av[0], av[2]) * sc);
            v[4] += __double2ll_rn((double) __fmul_rn(av[0], av[1]) * sc);
            v[5] += __double2ll_rn((double) __fmul_rn(av[1], av[2]) * sc);
            v[6] += __double2ll_rn((double) __fadd_rn(__fmul_rn(av[1], av[1]), __fmul_rn(av[2], av[2])) * sc);
            v[7] += __double2ll_rn((double) __fadd_rn(__fmul_rn(av[0], av[0]), __fmul_rn(av[1], av[1])) * sc);
            v[8] += __double2ll_rn((double) __fadd_rn(__fmul_rn(av[0], av[0]), __fmul_rn(av[2], av[2])) * sc);
            v[9] += __double2ll_rn((double) df[0] * sc);
            v[10] += __double2ll_rn((double) df[1] * sc);
            v[11] += __double2ll_rn((double) df[2] * sc);
            v[12] += __double2ll_rn((double) __fsub_rn(__fmul_rn(av[1], df[2]), __fmul_rn(av[2], df[1])) * sc);
            v[13] += __double2ll_rn((double) __fsub_rn(__fmul_rn(av[0], df[1]), __fmul_rn(av[1], df[0])) * sc);
            v[14] += __double2ll_rn((double) __fsub_rn(__fmul_rn(av[2], df[0]), __fmul_rn(av[0], df[2])) * sc);
        } else {
            const double *D = a.D;
            const double e0 = (double) df[0] - ((D[0] + (double) av[2] * D[5]) - (double) av[1] * D[4]);
            const double e1 = (double) df[1] - ((D[1] + (double) av[0] * D[4]) - (double) av[2] * D[3]);
            const double e2 = (double) df[2] - ((D[2] + (double) av[1] * D[3]) - (double) av[0] * D[5]);
            const float term = (float) ((e0 * e0 + e1 * e1) + e2 * e2);
            if (isfinite(term)) v[0] += __double2ll_rn((double) term * a.scale_ss);
            else v[1] += 1;  // poisons the sum: reported as non-finite
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    warp_reduce_transpose<NV>(v, lane);
    count = __reduce_add_sync(0xffffffffu, count);
    if ((lane & 1) == 0) s_part[warp][lane >> 1] = v[0];
    if (lane == 0) s_part[warp][NV] = count;
    __syncthreads();
    if (threadIdx.x <= NV) {
        __int128 tot = 0;
#pragma unroll
        for (int w = 0; w < kReduceWarps; ++w) tot += (__int128) s_part[w][threadIdx.x];
        if (tot != 0) {
            Acc128 *dst = a.acc + (blockIdx.x % kAccSlots) * kMaxAcc + threadIdx.x;
            atomic_add128(dst, (unsigned long long) tot, (long long) (tot >> 64));
        }
    }
}

// estimateLUMold's own correspondence pass: exact 1-NN of every aligned source point with
// d2 < max_corr^2 (strict), warm-started from the last ICP match.
struct LumOldArgs {
    const float4 *cur;
    const float4 *raw;
    int n_src;
    NnIndex ix;
    const int *warm;
    const IcpState *st;
    float thr_strict;
    int *pos_out;
};

__global__ void __launch_bounds__(kIterThreads, WCU_MINBLOCKS) lumold_corr_kernel(LumOldArgs a) {
    const int s = blockIdx.x * kIterThreads + threadIdx.x;
    if (s >= a.n_src) return;
    const float4 c = a.cur[s];
    const int orig = __float_as_int(c.w);
    if (orig == 0x7fffffff) {
        a.pos_out[s] = -1;
        return;
    }
    const float4 p = a.raw[orig];
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = a.st->T_final[k];
    const float x = xform_row(T + 0, p.x, p.y, p.z), y = xform_row(T + 4, p.x, p.y, p.z),
                z = xform_row(T + 8, p.x, p.y, p.z);
    float best = a.thr_strict;
    int best_idx = 0x7fffffff, best_pos = -1;
    const int warm = a.warm[s];
    if (warm >= 0) {
        const float4 q = __ldg(a.ix.pts + warm);
        const float d = l2_simple(x, y, z, q.x, q.y, q.z);
        if (d <= best) {
            best = d;
            best_idx = __float_as_int(q.w);
            best_pos = warm;
        }
    }
    nn_search(x, y, z, a.ix, best, best_idx, best_pos);
    a.pos_out[s] = best_pos;
}

__global__ void zero_acc_kernel(Acc128 *acc) {
    for (int i = threadIdx.x; i < kAccSlots * kMaxAcc; i += blockDim.x) {
        acc[i].lo = 0;
        acc[i].hi = 0;
    }
}

}  // namespace wavecu
