// Shared device/host helpers of libwavecu (sm_100a only; the whole library is compiled with
// -fmad=false so that every fp32/fp64 multiply and add below rounds separately, as the
// arithmetic spec in DESIGN.md requires; FMA is used only where written explicitly).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace wavecu {

void set_last_error(const std::string &msg);

#define WCU_CHECK(expr)                                                                              \
    do {                                                                                             \
        cudaError_t err__ = (expr);                                                                  \
        if (err__ != cudaSuccess) {                                                                  \
            ::wavecu::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(err__) + " at " + \
                                     __FILE__ + ":" + std::to_string(__LINE__));                     \
            return WAVECU_ERR_CUDA;                                                                  \
        }                                                                                            \
    } while (0)

// A leaf of the target tree is a run of <= kLeaf consecutive Morton-sorted points.  kLeaf = 1: every
// leaf is one point, stored in its parent's record as a degenerate box, so the 1-NN walk is one
// uniform loop of node steps (no separate leaf-scan phase for the lanes of a warp to wait on).
#ifndef WCU_LEAF
#define WCU_LEAF 1
#endif
constexpr int kLeaf = WCU_LEAF;
constexpr int kAccSlots = 16;           // accumulator replicas (atomic contention spreading)
constexpr int kMaxAcc = 40;             // values per slot (p2p uses 17, point-to-plane 33)

// Per-match constants computed on the device by setup_kernel (no host round trip).
struct MatchConsts {
    float src_lo[3], src_hi[3];
    float tgt_lo[3], tgt_hi[3];
    float thr;          // largest fp32 d2 that passes the (double) max-correspondence test
    int k_lin, k_quad, k_d2;
    double s_lin, s_quad, s_d2, s_plane;   // 2^k scale factors
};

// Device-resident iteration state of one align().
struct IcpState {
    float T_inc[16];      // incremental transform to apply at the start of the next iteration
    float T_final[16];    // final_transformation_
    double prev_mse;
    int iter;
    int done;
    int converged;
    int state;
    int n_corr;
    int pad;
    long long fb_total;   // queries the tiled correspondence kernel handed to the LBVH walk, summed over the iterations
};

struct TraceRow {
    double mse;
    int n_corr;
    int pad;
    float T[16];
};

// ---- fp32 arithmetic with the reference's rounding -----------------------------------------
// flann::L2_Simple<float>: r = ((dx*dx) + dy*dy) + dz*dz, separately rounded
__device__ __forceinline__ float l2_simple(float qx, float qy, float qz, float px, float py, float pz) {
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Lower bound of l2_simple over every point inside [lo,hi]: the same expression evaluated on the
// per-axis gap; rounding is monotone, so the bound never exceeds the distance of a contained point.
__device__ __forceinline__ float aabb_dist(float qx, float qy, float qz, const float4 &lo, const float4 &hi) {
    const float ex = fmaxf(fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)), 0.0f);
    const float ey = fmaxf(fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)), 0.0f);
    const float ez = fmaxf(fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)), 0.0f);
    return __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
}

// pt' = T (x,y,z,1): t = m_i0*x; t = m_i1*y + t; t = m_i2*z + t; t = m_i3 + t (Eigen 4x4 * 4x1)
__device__ __forceinline__ float xform_row(const float *m, float x, float y, float z) {
    float t = __fmul_rn(m[0], x);
    t = __fadd_rn(__fmul_rn(m[1], y), t);
    t = __fadd_rn(__fmul_rn(m[2], z), t);
    t = __fadd_rn(m[3], t);
    return t;
}

__device__ __forceinline__ bool finite3(float x, float y, float z) {
    return isfinite(x) && isfinite(y) && isfinite(z);
}

// ---- 128-bit accumulation --------------------------------------------------------------------
struct Acc128 {
    unsigned long long lo;
    long long hi;
};

// acc += x (sign-extended), order independent: the carry of each addition is derived from the
// value its own atomicAdd observed.
__device__ __forceinline__ void atomic_add128(Acc128 *acc, unsigned long long x_lo, long long x_hi) {
    const unsigned long long old = atomicAdd(&acc->lo, x_lo);
    const unsigned long long sum = old + x_lo;
    const long long carry = (sum < old) ? 1 : 0;
    const long long add_hi = x_hi + carry;
    if (add_hi != 0) atomicAdd(reinterpret_cast<unsigned long long *>(&acc->hi), (unsigned long long) add_hi);
}

// 128-bit fixed point -> fp64, sign-magnitude: |v| = hi * 2^64 + lo is converted (each half rounds
// once, the sum once more) and the sign re-applied.  Converting a negative total's two's-complement
// halves directly would round lo ~ 2^64 - |v| to 53 bits before the cancelling add and lose the
// low bits of small negative sums.  Same formula in oracle/smallmat.hpp and the host code.
__host__ __device__ __forceinline__ double acc_to_double(unsigned long long lo, long long hi, int k) {
    const bool neg = hi < 0;
    if (neg) {  // two's-complement negate of the 128-bit value
        lo = ~lo + 1ull;
        hi = ~hi + (lo == 0ull ? 1 : 0);
    }
    const double mag = ldexp((double) hi, 64 - k) + ldexp((double) lo, -k);
    return neg ? -mag : mag;
}

}  // namespace wavecu
