// Device side of one ICP iteration (pcl::IterativeClosestPoint::computeTransformation, SURVEY.md
// Appendix A.3-A.5; reference call sites wave_matching/src/icp.cpp:95,116,126).
//
//   correspond_kernel  per source point: the in-place incremental fp32 transform of the working
//                   cloud (A.3.2) fused with the exact 1-NN correspondence search and max-distance
//                   test (A.3.1).  Kept lean in registers (no accumulators live across the tree
//                   walk) so that 12 CTAs of 128 threads stay resident per SM - the walk is latency bound.
//   reduce_kernel   streaming pass over the correspondences: {n, sum p, sum q, sum q p^T, sum d2}
//                   for the SVD estimator (A.4) or the 21+6 entries of J^T J / J^T r for
//                   point-to-plane (8(a) A7) - as exact 128-bit fixed-point sums (DESIGN.md
//                   "Estimator arithmetic"): warp butterfly -> block -> a few 64-bit atomics.
//   solve_kernel    one thread: sums -> Umeyama rotation (one-sided Jacobi, fp64, fixed order) or
//                   6x6 solve, fp32 cast, final = T * final, DefaultConvergenceCriteria (A.5).
//   iterate_kernel  the default: the three of them in ONE launch per ICP iteration - every block reduces the
//                   pairs its own search just found, the last block to finish (a ticket) runs the solve.
//                   The separate kernels remain for the first iteration while the caller's normals are still
//                   being uploaded, for the tiled search (tile_nn.cuh) and for A/B runs (WAVECU_FUSED=0).
#pragma once
#include "common.cuh"
#include "index.cuh"

namespace wavecu {

#ifndef WCU_MINBLOCKS
// 12 blocks of 128 threads = 48 warps per SM at 40 registers (12 B of spills in the fused kernel).  Measured on the
// 1 M / 1 M match (ms per match, four matchers in flight / one at a time; 200 k batch scans/s): 8: 1.026 / - ; 10 (48
// registers): 0.958 / 1.222 / 3702; 12: 0.921 / 1.189 / 3802; 16 (32 registers, 76 B spills, and solve_block
// squeezed into 32 as well): 0.909 / 1.202 / 3172.
#define WCU_MINBLOCKS 12
#endif
#ifndef WCU_ITER_THREADS
#define WCU_ITER_THREADS 128
#endif
constexpr int kIterThreads = WCU_ITER_THREADS;  // correspondence kernel: one query per thread
#ifndef WCU_RED_THREADS
#define WCU_RED_THREADS 256
#endif
#ifndef WCU_RED_PER
#define WCU_RED_PER 8
#endif
constexpr int kReduceThreads = WCU_RED_THREADS;     // reduction kernel
constexpr int kReducePerThread = WCU_RED_PER;
constexpr int kReduceWarps = kReduceThreads / 32;

struct IterArgs {
    float4 *cur;                 // working source cloud, Morton order, w = original index
    int n_src;
    NnIndex ix;                  // target search structure
    const float4 *tgt;           // == ix.pts
    const float4 *nrm;           // sorted target normals (point-to-plane) or nullptr
    int *nn_pos;                 // sorted position of the match (-1: none) - warm start + gathers
    int *nn_idx;                 // original target index of the match (-1: none)
    float *nn_d2;
    IcpState *st;
    const MatchConsts *mc;
    Acc128 *acc;                 // [kAccSlots][kMaxAcc]
};

template <int EST>
struct EstTraits;
template <>
struct EstTraits<WAVECU_EST_SVD> {
    static constexpr int NV = 16;  // 3 + 3 + 9 + 1 (count travels separately)
};
template <>
struct EstTraits<WAVECU_EST_POINT_TO_PLANE> {
    static constexpr int NV = 32;  // 21 + 6 + 1, padded
};

__device__ __forceinline__ long long shfl_xor_ll(long long v, int mask) {
    int lo = (int) (unsigned long long) v, hi = (int) ((unsigned long long) v >> 32);
    lo = __shfl_xor_sync(0xffffffffu, lo, mask);
    hi = __shfl_xor_sync(0xffffffffu, hi, mask);
    return (long long) (((unsigned long long) (unsigned) hi << 32) | (unsigned) lo);
}

// Transposing warp reduction of NV (16 or 32) 64-bit values: ~NV shuffles instead of 5 NV.
// On return v[0] of lane L holds the warp total of value (L * NV) / 32.
template <int NV>
__device__ __forceinline__ void warp_reduce_transpose(long long (&v)[NV], int lane) {
    int bit = 16;
#pragma unroll
    for (int half = NV / 2; half >= 1; half >>= 1) {
        const bool upper = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const long long keep = upper ? v[half + i] : v[i];
            const long long send = upper ? v[i] : v[half + i];
            v[i] = keep + shfl_xor_ll(send, bit);
        }
        bit >>= 1;
    }
    for (; bit >= 1; bit >>= 1) v[0] += shfl_xor_ll(v[0], bit);
}

// Correspondence kernel (A.3.2 + A.3.1): cur <- T_inc (x) cur in place, then the exact 1-NN of the
// moved point within the max-correspondence distance, warm-started from the previous match.
__global__ void __launch_bounds__(kIterThreads, WCU_MINBLOCKS) correspond_kernel(IterArgs a) {
    if (a.st->done) return;
    __shared__ float sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = a.st->T_inc[threadIdx.x];
    __syncthreads();
    const int s = blockIdx.x * kIterThreads + threadIdx.x;
    if (s >= a.n_src) return;
    const float4 c = a.cur[s];
    if (!finite3(c.x, c.y, c.z)) {  // pads and non-finite source points take no part
        a.nn_pos[s] = -1;
        a.nn_idx[s] = -1;
        a.nn_d2[s] = INFINITY;
        return;
    }
    const float x = xform_row(sT + 0, c.x, c.y, c.z);
    const float y = xform_row(sT + 4, c.x, c.y, c.z);
    const float z = xform_row(sT + 8, c.x, c.y, c.z);
    a.cur[s] = make_float4(x, y, z, c.w);

    float best = a.mc->thr;
    int best_idx = 0x7fffffff, best_pos = -1;
    const int warm = a.st->iter > 0 ? a.nn_pos[s] : -1;   // the first iteration has no previous match
    if (warm >= 0) {
        const float4 p = __ldg(a.tgt + warm);
        const float d = l2_simple(x, y, z, p.x, p.y, p.z);
        if (d <= best) {
            best = d;
            best_idx = __float_as_int(p.w);
            best_pos = warm;
        }
    }
#if WCU_LEAF == 1
    nn_search_cells(x, y, z, a.ix, best, best_idx, best_pos);
#else
    nn_search(x, y, z, a.ix, best, best_idx, best_pos);
#endif
    a.nn_pos[s] = best_pos;
    a.nn_idx[s] = best_pos >= 0 ? best_idx : -1;
    a.nn_d2[s] = best;
}

// Estimator reduction over the correspondences (A.4 / 8(a) A7): a streaming pass - 16 B working
// point + 4 B match position + 4 B distance per source point, 16 B (+16 B normal) gathered per
// pair - into exact 128-bit fixed-point sums.  One pair per thread; the 16 / 28 terms of a pair
// are formed and warp-reduced eight at a time, so a thread never holds more than eight 64-bit
// partial sums (48 registers instead of 104: the pass is latency bound and needs the occupancy).
// Warp totals go to shared memory, block totals to a few 64-bit global atomics.
__device__ __forceinline__ long long fix_term(double v, double scale) { return __double2ll_rn(v * scale); }

// Block-level body shared by reduce_kernel and the fused iteration kernel: the calling thread's pair
// (if any) is its moved source point c, its match at sorted position pos and their fp32 distance.
template <int EST>
__device__ __forceinline__ void reduce_block(const IterArgs &a, bool pair, const float4 &c, int pos, float d2f,
                                             unsigned long long *s_acc /* shared, NV + 1 */) {
    constexpr int NV = EstTraits<EST>::NV;
    if (threadIdx.x <= NV) s_acc[threadIdx.x] = 0ull;
    __syncthreads();
    const double s_lin = a.mc->s_lin, s_quad = a.mc->s_quad, s_d2 = a.mc->s_d2, s_plane = a.mc->s_plane;
    const int lane = threadIdx.x & 31;
    double in[7] = {0, 0, 0, 0, 0, 0, 0};  // SVD: p, q.  PLANE: J[0..5], d
    double d2 = 0.0;
    bool plane_ok = false;
    if (pair) {
        d2 = (double) d2f;
        const float4 q = __ldg(a.tgt + pos);
        if (EST == WAVECU_EST_SVD) {
            in[0] = c.x; in[1] = c.y; in[2] = c.z;
            in[3] = q.x; in[4] = q.y; in[5] = q.z;
        } else {
            const float4 nn = __ldg(a.nrm + pos);
            plane_ok = finite3(nn.x, nn.y, nn.z);
            if (plane_ok) {
                // TransformationEstimationPointToPlaneLLS: a, b, c, d evaluated in fp32
                const float x = c.x, y = c.y, z = c.z;
                in[0] = __fsub_rn(__fmul_rn(nn.z, y), __fmul_rn(nn.y, z));
                in[1] = __fsub_rn(__fmul_rn(nn.x, z), __fmul_rn(nn.z, x));
                in[2] = __fsub_rn(__fmul_rn(nn.y, x), __fmul_rn(nn.x, y));
                in[3] = nn.x;
                in[4] = nn.y;
                in[5] = nn.z;
                float df = __fadd_rn(__fadd_rn(__fmul_rn(nn.x, q.x), __fmul_rn(nn.y, q.y)), __fmul_rn(nn.z, q.z));
                df = __fsub_rn(df, __fmul_rn(nn.x, x));
                df = __fsub_rn(df, __fmul_rn(nn.y, y));
                df = __fsub_rn(df, __fmul_rn(nn.z, z));
                in[6] = df;
            }
        }
    }
    // term t of the pair: SVD 0-2 p, 3-5 q, 6-14 q p^T (row major), 15 d2;
    // PLANE 0-20 J^T J (upper triangle, row major), 21-26 J^T d, 27 d2, 28-31 unused
    auto term = [&](int t) -> long long {
        if (EST == WAVECU_EST_SVD) {
            if (!pair) return 0;
            if (t < 6) return fix_term(in[t], s_lin);
            if (t < 15) return fix_term(in[3 + (t - 6) / 3] * in[(t - 6) % 3], s_quad);
            return fix_term(d2, s_d2);
        } else {
            if (t == 27) return pair ? fix_term(d2, s_d2) : 0;
            if (!plane_ok || t > 27) return 0;
            if (t >= 21) return fix_term(in[t - 21] * in[6], s_plane);
            int i = 0, base = 0;  // row i of the upper triangle starts at base
            while (t >= base + (6 - i)) {
                base += 6 - i;
                ++i;
            }
            return fix_term(in[i] * in[i + (t - base)], s_plane);
        }
    };
#pragma unroll
    for (int chunk = 0; chunk < NV / 8; ++chunk) {
        long long v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = term(chunk * 8 + k);  // indices are compile-time after unrolling
        warp_reduce_transpose<8>(v, lane);  // lane L now holds the warp total of term chunk*8 + L/4
        if ((lane & 3) == 0 && v[0] != 0) atomicAdd(&s_acc[chunk * 8 + (lane >> 2)], (unsigned long long) v[0]);
    }
    const int count = __reduce_add_sync(0xffffffffu, pair ? 1 : 0);
    if (lane == 0 && count) atomicAdd(&s_acc[NV], (unsigned long long) count);
    __syncthreads();
    if (threadIdx.x <= NV) {
        const long long tot = (long long) s_acc[threadIdx.x];
        if (tot != 0) {
            Acc128 *dst = a.acc + (blockIdx.x % kAccSlots) * kMaxAcc + threadIdx.x;
            atomic_add128(dst, (unsigned long long) tot, tot < 0 ? -1LL : 0LL);
        }
    }
}

template <int EST>
__global__ void __launch_bounds__(kReduceThreads) reduce_kernel(IterArgs a) {
    if (a.st->done) return;
    __shared__ unsigned long long s_acc[EstTraits<EST>::NV + 1];
    const int s = blockIdx.x * kReduceThreads + threadIdx.x;
    const int pos = (s < a.n_src) ? a.nn_pos[s] : -1;
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    float d2 = 0.0f;
    if (pos >= 0) {
        c = a.cur[s];
        d2 = a.nn_d2[s];
    }
    reduce_block<EST>(a, pos >= 0, c, pos, d2, s_acc);
}

struct SolveArgs {
    IcpState *st;
    const MatchConsts *mc;
    Acc128 *acc;
    TraceRow *trace;
    int max_iter;
    double t_eps, fit_eps;
    // mapped host word, (launch number << 1) | done: lets the host follow the iterations without
    // putting a copy or an event between the kernels
    volatile int *progress;
    int launch;
    int *fb_count;   // fallback list length of the tiled correspondence kernel: cleared for the next iteration
};

// (a posted store: nothing waits for it - it is visible to the host at the latest when the kernel ends)
__device__ __forceinline__ void publish_progress(const SolveArgs &a, int done) {
    *a.progress = (a.launch << 1) | (done ? 1 : 0);
}

// Rotation maximising trace(R^T S): one-sided Jacobi on S, fixed pair order, <= 12 sweeps (stops
// early only at an exact fixed point, which leaves the result unchanged).
__host__ __device__ inline void rotation_from_sigma(const double S[9], double R[9]) {
    double A[3][3], V[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            A[i][j] = S[3 * i + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 12; ++sweep) {
        bool changed = false;
#pragma unroll
        for (int pr = 0; pr < 3; ++pr) {
            const int p = (pr == 2) ? 1 : 0, q = (pr == 0) ? 1 : 2;
            const double alpha = (A[0][p] * A[0][p] + A[1][p] * A[1][p]) + A[2][p] * A[2][p];
            const double beta = (A[0][q] * A[0][q] + A[1][q] * A[1][q]) + A[2][q] * A[2][q];
            const double gamma = (A[0][p] * A[0][q] + A[1][p] * A[1][q]) + A[2][p] * A[2][q];
            if (gamma == 0.0) continue;
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t);
            const double s = c * t;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const double ap = A[i][p], aq = A[i][q];
                const double nap = c * ap - s * aq, naq = s * ap + c * aq;
                const double vp = V[i][p], vq = V[i][q];
                const double nvp = c * vp - s * vq, nvq = s * vp + c * vq;
                changed |= (nap != ap) | (naq != aq) | (nvp != vp) | (nvq != vq);
                A[i][p] = nap;
                A[i][q] = naq;
                V[i][p] = nvp;
                V[i][q] = nvq;
            }
        }
        if (!changed) break;
    }
    double nrm2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) nrm2[j] = (A[0][j] * A[0][j] + A[1][j] * A[1][j]) + A[2][j] * A[2][j];
    int i1 = 0;
    if (nrm2[1] > nrm2[i1]) i1 = 1;
    if (nrm2[2] > nrm2[i1]) i1 = 2;
    int i2 = -1;
    for (int j = 0; j < 3; ++j) {
        if (j == i1) continue;
        if (i2 < 0 || nrm2[j] > nrm2[i2]) i2 = j;
    }
    const double s1 = sqrt(nrm2[i1]), s2 = sqrt(nrm2[i2]);
    if (!(s1 > 0.0) || !(s2 > 0.0)) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double u1[3], u2[3], v1[3], v2[3], u3[3], v3[3];
    for (int i = 0; i < 3; ++i) {
        u1[i] = A[i][i1] / s1;
        u2[i] = A[i][i2] / s2;
        v1[i] = V[i][i1];
        v2[i] = V[i][i2];
    }
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
    u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
    u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
    v3[1] = v1[2] * v2[0] - v1[0] * v2[2];
    v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = (u1[i] * v1[j] + u2[i] * v2[j]) + u3[i] * v3[j];
}

// 6x6 Gaussian elimination with partial pivoting, fixed order; false if singular.
__device__ inline bool solve6(const double A_in[36], const double b_in[6], double x[6]) {
    double A[6][7];
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) A[i][j] = A_in[6 * i + j];
        A[i][6] = b_in[i];
    }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        double best = fabs(A[c][c]);
        for (int r = c + 1; r < 6; ++r)
            if (fabs(A[r][c]) > best) {
                best = fabs(A[r][c]);
                piv = r;
            }
        if (best == 0.0 || !isfinite(best)) return false;
        if (piv != c)
            for (int j = 0; j <= 6; ++j) {
                const double t = A[c][j];
                A[c][j] = A[piv][j];
                A[piv][j] = t;
            }
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j <= 6; ++j) A[r][j] = A[r][j] - f * A[c][j];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = A[r][6];
        for (int j = r + 1; j < 6; ++j) s = s - A[r][j] * x[j];
        x[r] = s / A[r][r];
    }
    return true;
}

// Estimator + convergence test for the calling block (>= 33 threads): used by solve_kernel, and by the
// last block of the fused iteration kernel, which reads the accumulators other blocks just added to
// (hence the L1-bypassing loads).
template <int EST>
__device__ __noinline__ void solve_block(const SolveArgs &a) {
    constexpr int NV = EstTraits<EST>::NV;
    // threads 0..NV: one accumulator each, summed over the slots (and cleared for the next
    // iteration); then thread 0 runs the estimator
    __shared__ unsigned long long s_lo[NV + 1];
    __shared__ long long s_hi[NV + 1];
    __shared__ double s_val[NV + 1];
    if (threadIdx.x <= NV) {
        __int128 t = 0;
#pragma unroll 4
        for (int sl = 0; sl < kAccSlots; ++sl) {
            Acc128 &c = a.acc[sl * kMaxAcc + threadIdx.x];
            const unsigned long long lo = __ldcg(&c.lo);
            const long long hi = __ldcg(&c.hi);
            t += ((__int128) hi << 64) + (__int128) lo;
            c.lo = 0;
            c.hi = 0;
        }
        s_lo[threadIdx.x] = (unsigned long long) t;
        s_hi[threadIdx.x] = (long long) (t >> 64);
        // each thread also converts its own sum (the fixed-point exponent depends only on the slot)
        const int i = threadIdx.x;
        int k;
        if (EST == WAVECU_EST_SVD) k = (i < 6) ? a.mc->k_lin : (i < 15 ? a.mc->k_quad : a.mc->k_d2);
        else k = (i == 27) ? a.mc->k_d2 : a.mc->k_quad - 3;
        s_val[i] = acc_to_double((unsigned long long) t, (long long) (t >> 64), k);
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    IcpState &st = *a.st;
    if (a.fb_count) {
        st.fb_total += *a.fb_count;
        *a.fb_count = 0;
    }
    const MatchConsts &mc = *a.mc;
    auto val = [&](int i, int /*k*/) { return s_val[i]; };
    const long long n = (long long) s_lo[NV];
    if (n < 3) {  // min_number_correspondences_
        st.n_corr = (int) n;
        st.converged = 0;
        st.state = WAVECU_CONV_NO_CORRESPONDENCES;
        st.done = 1;
        publish_progress(a, 1);
        return;
    }
    const double dn = (double) n;
    float T[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    double mse;
    if (EST == WAVECU_EST_SVD) {
        double sp[3], sq[3], mp[3], mq[3], S[9], R[9];
        for (int c = 0; c < 3; ++c) {
            sp[c] = val(c, mc.k_lin);
            sq[c] = val(3 + c, mc.k_lin);
            mp[c] = sp[c] / dn;
            mq[c] = sq[c] / dn;
        }
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) S[3 * r + c] = (val(6 + 3 * r + c, mc.k_quad) - (sq[r] * sp[c]) / dn) / dn;
        rotation_from_sigma(S, R);
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c) T[4 * r + c] = (float) R[3 * r + c];
            const double rr = (R[3 * r + 0] * mp[0] + R[3 * r + 1] * mp[1]) + R[3 * r + 2] * mp[2];
            T[4 * r + 3] = (float) (mq[r] - rr);
        }
        mse = val(15, mc.k_d2) / dn;
    } else {
        double ATA[36], ATb[6], x[6];
        const int k = mc.k_quad - 3;
        int u = 0;
        for (int r = 0; r < 6; ++r) {
            for (int c = r; c < 6; ++c) {
                ATA[6 * r + c] = ATA[6 * c + r] = val(u, k);
                ++u;
            }
            ATb[r] = val(21 + r, k);
        }
        mse = val(27, mc.k_d2) / dn;
        if (!solve6(ATA, ATb, x)) {
            st.n_corr = (int) n;
            st.converged = 0;
            st.state = WAVECU_CONV_NO_CORRESPONDENCES;
            st.done = 1;
            publish_progress(a, 1);
            return;
        }
        const double al = x[0], be = x[1], ga = x[2];
        T[0] = (float) (cos(ga) * cos(be));
        T[1] = (float) (-sin(ga) * cos(al) + cos(ga) * sin(be) * sin(al));
        T[2] = (float) (sin(ga) * sin(al) + cos(ga) * sin(be) * cos(al));
        T[4] = (float) (sin(ga) * cos(be));
        T[5] = (float) (cos(ga) * cos(al) + sin(ga) * sin(be) * sin(al));
        T[6] = (float) (-cos(ga) * sin(al) + sin(ga) * sin(be) * cos(al));
        T[8] = (float) (-sin(be));
        T[9] = (float) (cos(be) * sin(al));
        T[10] = (float) (cos(be) * cos(al));
        T[3] = (float) x[3];
        T[7] = (float) x[4];
        T[11] = (float) x[5];
    }

    // final_transformation_ = transformation_ * final_transformation_ (fp32, Eigen column order)
    float F[16];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            float acc = __fmul_rn(T[4 * r + 0], st.T_final[c]);
            acc = __fadd_rn(__fmul_rn(T[4 * r + 1], st.T_final[4 + c]), acc);
            acc = __fadd_rn(__fmul_rn(T[4 * r + 2], st.T_final[8 + c]), acc);
            acc = __fadd_rn(__fmul_rn(T[4 * r + 3], st.T_final[12 + c]), acc);
            F[4 * r + c] = acc;
        }
    for (int i = 0; i < 16; ++i) {
        st.T_final[i] = F[i];
        st.T_inc[i] = T[i];
    }
    const int it = ++st.iter;
    st.n_corr = (int) n;
    if (a.trace) {
        TraceRow &row = a.trace[it - 1];
        row.mse = mse;
        row.n_corr = (int) n;
        row.pad = 0;
        for (int i = 0; i < 16; ++i) row.T[i] = T[i];
    }

    // DefaultConvergenceCriteria::hasConverged; cos_angle / translation_sqr are fp32 expressions
    int conv = 0, state = WAVECU_CONV_NOT_CONVERGED;
    if (it >= a.max_iter) {
        conv = 1;
        state = WAVECU_CONV_ITERATIONS;
    } else {
        const float tr = __fsub_rn(__fadd_rn(__fadd_rn(T[0], T[5]), T[10]), 1.0f);
        const double cos_angle = 0.5 * (double) tr;
        const float t2 = __fadd_rn(__fadd_rn(__fmul_rn(T[3], T[3]), __fmul_rn(T[7], T[7])), __fmul_rn(T[11], T[11]));
        const double translation_sqr = (double) t2;
        if (cos_angle >= 1.0 - a.t_eps && translation_sqr <= a.t_eps) {
            conv = 1;
            state = WAVECU_CONV_TRANSFORM;
        } else if (fabs(mse - st.prev_mse) < 1e-12) {
            conv = 1;
            state = WAVECU_CONV_ABS_MSE;
        } else if (fabs(mse - st.prev_mse) / st.prev_mse < a.fit_eps) {
            conv = 1;
            state = WAVECU_CONV_REL_MSE;
        } else {
            st.prev_mse = mse;
        }
    }
    if (conv) {
        st.converged = 1;
        st.state = state;
        st.done = 1;
    }
    publish_progress(a, conv);
}

template <int EST>
__global__ void __launch_bounds__(64) solve_kernel(SolveArgs a) {
    if (a.st->done) {
        if (threadIdx.x == 0) publish_progress(a, 1);
        return;
    }
    solve_block<EST>(a);
}

// One ICP iteration in one launch: correspond_kernel's search, the estimator reduction of the pair it
// just found (no second pass over the cloud), and - in whichever block finishes last - the estimator
// and convergence test.  Blocks order their accumulator updates before their ticket with a fence; the
// last block resets the ticket for the next launch.
struct FusedArgs {
    IterArgs it;
    SolveArgs so;
    unsigned *ticket;
};

template <int EST>
__global__ void __launch_bounds__(kIterThreads, WCU_MINBLOCKS) iterate_kernel(const __grid_constant__ FusedArgs f) {
    const IterArgs &a = f.it;
    if (a.st->done) {
        if (blockIdx.x == 0 && threadIdx.x == 0) publish_progress(f.so, 1);
        return;
    }
    __shared__ float sT[12];
    __shared__ unsigned long long s_acc[EstTraits<EST>::NV + 1];
    __shared__ bool s_last;
    if (threadIdx.x < 12) sT[threadIdx.x] = a.st->T_inc[threadIdx.x];
    __syncthreads();
    const int s = blockIdx.x * kIterThreads + threadIdx.x;
    float4 moved = make_float4(0.f, 0.f, 0.f, 0.f);
    float best = a.mc->thr;
    int best_idx = 0x7fffffff, best_pos = -1;
    if (s < a.n_src) {
        const float4 c = a.cur[s];
        if (finite3(c.x, c.y, c.z)) {
            moved = make_float4(xform_row(sT + 0, c.x, c.y, c.z), xform_row(sT + 4, c.x, c.y, c.z),
                                xform_row(sT + 8, c.x, c.y, c.z), c.w);
            a.cur[s] = moved;
            const int warm = a.st->iter > 0 ? a.nn_pos[s] : -1;   // the first iteration has no previous match
            if (warm >= 0) {
                const float4 p = __ldg(a.tgt + warm);
                const float d = l2_simple(moved.x, moved.y, moved.z, p.x, p.y, p.z);
                if (d <= best) {
                    best = d;
                    best_idx = __float_as_int(p.w);
                    best_pos = warm;
                }
            }
#if WCU_LEAF == 1
            nn_search_cells(moved.x, moved.y, moved.z, a.ix, best, best_idx, best_pos);
#else
            nn_search(moved.x, moved.y, moved.z, a.ix, best, best_idx, best_pos);
#endif
            a.nn_pos[s] = best_pos;
            a.nn_idx[s] = best_pos >= 0 ? best_idx : -1;
            a.nn_d2[s] = best;
        } else {  // pads and non-finite source points take no part
            a.nn_pos[s] = -1;
            a.nn_idx[s] = -1;
            a.nn_d2[s] = INFINITY;
        }
    }
    reduce_block<EST>(a, best_pos >= 0, moved, best_pos, best, s_acc);
    __threadfence();   // this block's accumulator updates (and match arrays) before its ticket
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(f.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *f.ticket = 0u;
    solve_block<EST>(f.so);
}

}  // namespace wavecu
