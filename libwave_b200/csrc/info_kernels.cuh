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
    nn_search_cells(x, y, z, a.ix, best, best_idx, best_pos);
    a.pos_out[s] = best_pos;
}

__global__ void zero_acc_kernel(Acc128 *acc) {
    for (int i = threadIdx.x; i < kAccSlots * kMaxAcc; i += blockDim.x) {
        acc[i].lo = 0;
        acc[i].hi = 0;
    }
}

// ---- Censi / Haralick closed-form covariance (ICPMatcher::estimateCensi, src/icp.cpp:167-397) ----
// Per correspondence (a = matched target point, b = reference point as handed to the last align())
// the second derivatives of J = | t + R a - b |^2, R = Rz(yaw) Ry(pitch) Rx(roll), at the match
// result: d2J/dX2 (15 non-constant upper entries) and D cov_Z D^T with D = d2J/dZdX (21 entries of
// a symmetric 6x6).  R and its first / second partial derivatives are per-match constants built on
// the host; the measurement covariance of each point follows the reference: fp32 range, bearing
// and elevation (sqrtf / atan2f / atanf), j diag(lin, ang, ang) j^T with the reference's j.
// Sums are fp64 per thread -> warp -> block, one partial row per block, added by the host in
// block order (deterministic; agreement with the CPU oracle is to ~1e-6 relative, limited by the
// ulp-level differences between the device's and glibc's atan2f / atanf).
constexpr int kCensiValues = 37;  // 15 of H, 21 of middle, pair count
constexpr int kCensiThreads = 128;

struct CensiConsts {
    double t[3];
    double R[9];
    double dR[3][9];     // d/droll, d/dpitch, d/dyaw
    double ddR[6][9];    // rr, rp, ry, pp, py, yy
    double lin, ang;
};

struct CensiArgs {
    const float4 *cur;   // Morton-ordered working source (w = original index)
    const float4 *raw;   // the source cloud of the last align(), original order
    int n_src;
    const float4 *tgt;   // Morton-ordered target
    const int *pos;      // per sorted source point: matched target position or -1
    double *partial;     // [gridDim.x][kCensiValues]
    CensiConsts c;
};

__device__ __forceinline__ void censi_point_cov(float x, float y, float z, double lin, double ang, double cov[9]) {
    const float rho2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const double rg = (double) sqrtf(__fadd_rn(rho2, __fmul_rn(z, z)));
    const double br = (double) atan2f(y, x);
    const double az = (double) atanf(__fdiv_rn(z, sqrtf(rho2)));
    const double cb = cos(br), sb = sin(br), ca = cos(az), sa = sin(az);
    const double j[9] = {cb * sa, -rg * sb * sa, rg * cb * ca, sb * sa, rg * cb * sa, rg * ca * sb, ca, 0.0, -rg * sa};
    const double w[3] = {lin, ang, ang};
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            cov[3 * r + c] = (j[3 * r] * w[0] * j[3 * c] + j[3 * r + 1] * w[1] * j[3 * c + 1]) + j[3 * r + 2] * w[2] * j[3 * c + 2];
}

__global__ void __launch_bounds__(kCensiThreads) censi_kernel(CensiArgs a) {
    double v[kCensiValues];
#pragma unroll
    for (int i = 0; i < kCensiValues; ++i) v[i] = 0.0;
    const CensiConsts &k = a.c;
    for (int s = blockIdx.x * kCensiThreads + threadIdx.x; s < a.n_src; s += gridDim.x * kCensiThreads) {
        const int pos = a.pos[s];
        if (pos < 0) continue;
        const float4 pa = __ldg(a.tgt + pos);
        const float4 pb = a.raw[__float_as_int(a.cur[s].w)];
        const double av[3] = {pa.x, pa.y, pa.z}, bv[3] = {pb.x, pb.y, pb.z};
        double e[3], g[3][3];
#pragma unroll
        for (int i = 0; i < 3; ++i)
            e[i] = k.t[i] + ((k.R[3 * i] * av[0] + k.R[3 * i + 1] * av[1]) + k.R[3 * i + 2] * av[2]) - bv[i];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                g[q][i] = (k.dR[q][3 * i] * av[0] + k.dR[q][3 * i + 1] * av[1]) + k.dR[q][3 * i + 2] * av[2];
        // H(t_i, k) -> v[3 i + k]; H(k, l), k <= l -> v[9 + ...]
        int u = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int q = 0; q < 3; ++q) v[u++] += 2.0 * g[q][i];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int l = q; l < 3; ++l) {
                const int sec = (q == 0) ? l : (q == 1 ? 2 + l : 5);  // rr rp ry | pp py | yy
                const double *S = k.ddR[sec];
                double ea = 0.0;
#pragma unroll
                for (int i = 0; i < 3; ++i) ea += e[i] * ((S[3 * i] * av[0] + S[3 * i + 1] * av[1]) + S[3 * i + 2] * av[2]);
                v[u++] += 2.0 * ((g[q][0] * g[l][0] + g[q][1] * g[l][1]) + g[q][2] * g[l][2]) + 2.0 * ea;
            }
        // D(z, x) = d2J / dZ_z dX_x
        double D[36];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                D[6 * i + j] = 2.0 * k.R[3 * j + i];
                D[6 * (3 + i) + j] = (i == j) ? -2.0 : 0.0;
            }
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double rg = (k.R[i] * g[q][0] + k.R[3 + i] * g[q][1]) + k.R[6 + i] * g[q][2];
                const double er = (e[0] * k.dR[q][i] + e[1] * k.dR[q][3 + i]) + e[2] * k.dR[q][6 + i];
                D[6 * i + 3 + q] = 2.0 * rg + 2.0 * er;
                D[6 * (3 + i) + 3 + q] = -2.0 * g[q][i];
            }
        double ca[9], cb[9];
        censi_point_cov(pa.x, pa.y, pa.z, k.lin, k.ang, ca);
        censi_point_cov(pb.x, pb.y, pb.z, k.lin, k.ang, cb);
        // middle += D cov_Z D^T (cov_Z = blockdiag(ca, cb) meets D's second index), upper triangle
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double Da[3], Db[3];  // row i of D times the two covariance blocks
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                Da[l] = (D[6 * i] * ca[l] + D[6 * i + 1] * ca[3 + l]) + D[6 * i + 2] * ca[6 + l];
                Db[l] = (D[6 * i + 3] * cb[l] + D[6 * i + 4] * cb[3 + l]) + D[6 * i + 5] * cb[6 + l];
            }
#pragma unroll
            for (int j = i; j < 6; ++j)
                v[u++] += ((Da[0] * D[6 * j] + Da[1] * D[6 * j + 1]) + Da[2] * D[6 * j + 2]) +
                          ((Db[0] * D[6 * j + 3] + Db[1] * D[6 * j + 4]) + Db[2] * D[6 * j + 5]);
        }
        v[36] += 1.0;
    }
    __shared__ double s_part[kCensiThreads / 32][kCensiValues];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kCensiValues; ++i) {
        double x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) s_part[warp][i] = x;
    }
    __syncthreads();
    if (threadIdx.x < kCensiValues) {
        double x = 0.0;
#pragma unroll
        for (int w = 0; w < kCensiThreads / 32; ++w) x += s_part[w][threadIdx.x];
        a.partial[(size_t) blockIdx.x * kCensiValues + threadIdx.x] = x;
    }
}

}  // namespace wavecu
