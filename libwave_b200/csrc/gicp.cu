// GICPMatcher on the device (wavecu_gicp_*): replaces pcl::GeneralizedIterativeClosestPoint as the
// reference drives it (wave_matching/src/gicp.cpp:20-64; PCL 1.8 registration/impl/gicp.hpp and
// registration/bfgs.h).
//
//   covariance kernel   computeCovariances: exact k-NN over the cloud's own LBVH, covariance of the
//                       neighbours (fp32 products, fp64 sums, as PCL writes them), 3x3 Jacobi,
//                       singular values replaced by (1, 1, 1e-3); once per cloud.
//   correspondence      per outer iteration: q = transformation_ * p (fp32), exact 1-NN with
//   kernel              d2 < 25 (strict, PCL's corr_dist_threshold_ = 5), and the Mahalanobis matrix
//                       M = (R C1 R^T + C2)^-1 in fp64, stored per source point.
//   cost kernel         one pass over the pairs per function / gradient evaluation of the BFGS line
//                       search: f = sum r^T M r, g_t = sum M r, Racc = sum p (M r)^T - 13 fp64 sums,
//                       block partials combined in a fixed order (deterministic).
//   host                pcl::BFGS (a port of GSL's vector_bfgs2: Fletcher bracketing / sectioning with
//                       cubic and quadratic interpolation), <= 20 inner iterations per outer one, and
//                       GICP's delta test (entries of the fp32 transform scaled by 1/rotation_epsilon
//                       or 1/transformation_epsilon).
#include <sched.h>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/wavecu.h"
#include "index.cuh"
#include "linalg.cuh"
#include "voxel.cuh"

namespace wavecu {

namespace {

struct GicpPose {
    float T[16];
    double scale;   // 2^k of the fixed-point cost sums
    int k;
};

constexpr int kGicpThreads = 128;
#ifndef WCU_GICP_U
#define WCU_GICP_U 1          // pairs per thread and trip of the cost kernel (measured: 1 and 4 equal, 2 slower)
#endif
#ifndef WCU_GICP_PPT
#define WCU_GICP_PPT 4        // pairs per thread of the cost kernel (sets its grid; measured 1 / 2 / 4 / 8: 29.8 / 24.7 / 22.2 / 24.1 us)
#endif
constexpr int kCostVals = 14;  // f, g_t(3), Racc(9), pair count
constexpr int kCostSlots = 16; // accumulator replicas of the cost kernel (atomic contention spreading)

// one thread per Morton-sorted point; covs indexed by sorted position (9 doubles, row major)
__global__ void __launch_bounds__(kGicpThreads) gicp_cov_kernel(NnIndex ix, int n, int k, double eps, double *covs) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 q = ix.pts[s];
    double *out = covs + 9 * (size_t) s;
    if (__float_as_int(q.w) == 0x7fffffff) {  // pad (non-finite point)
        for (int i = 0; i < 9; ++i) out[i] = 0.0;
        return;
    }
    KnnList nb;
    nb.init(k);
    knn_search(q.x, q.y, q.z, ix, nb);
    double mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int j = 0; j < k; ++j) {
        if (nb.pos[j] < 0) break;
        const float4 p = __ldg(ix.pts + nb.pos[j]);
        mean[0] += p.x;
        mean[1] += p.y;
        mean[2] += p.z;
        cov[0] += __fmul_rn(p.x, p.x);  // fp32 products, as gicp.hpp writes them
        cov[3] += __fmul_rn(p.y, p.x);
        cov[4] += __fmul_rn(p.y, p.y);
        cov[6] += __fmul_rn(p.z, p.x);
        cov[7] += __fmul_rn(p.z, p.y);
        cov[8] += __fmul_rn(p.z, p.z);
    }
    for (int d = 0; d < 3; ++d) mean[d] /= (double) k;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c <= r; ++c) {
            cov[3 * r + c] /= (double) k;
            cov[3 * r + c] -= mean[r] * mean[c];
            cov[3 * c + r] = cov[3 * r + c];
        }
    double U[9];
    eig_sym3_desc(cov, U);
    for (int i = 0; i < 9; ++i) out[i] = 0.0;
    for (int kk = 0; kk < 3; ++kk) {
        const double v = (kk == 2) ? eps : 1.0;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) out[3 * r + c] += v * U[3 * r + kk] * U[3 * c + kk];
    }
}

struct GicpIterConsts {
    float T[12];   // transformation_ (fp32), applied as Eigen's 4x4 * 4x1
    double R[9];   // its 3x3 block widened to fp64
    float thr;     // largest fp32 d2 with (double) d2 < corr_dist_threshold^2
};

// per sorted source point: match position (-1: none) and Mahalanobis matrix
__global__ void __launch_bounds__(kGicpThreads) gicp_corr_kernel(const float4 *__restrict__ src_sorted, int n_src,
                                                                 NnIndex tgt, const double *__restrict__ cov_src,
                                                                 const double *__restrict__ cov_tgt,
                                                                 const GicpIterConsts *__restrict__ kc, int *pos,
                                                                 double *mahal, size_t ld) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_src) return;
    const float4 p = src_sorted[s];
    if (!finite3(p.x, p.y, p.z)) {
        pos[s] = -1;
        return;
    }
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = kc->T[i];
    const float x = xform_row(T + 0, p.x, p.y, p.z), y = xform_row(T + 4, p.x, p.y, p.z), z = xform_row(T + 8, p.x, p.y, p.z);
    float best = kc->thr;
    int best_idx = 0x7fffffff, best_pos = -1;
    const int warm = pos[s];
    if (warm >= 0) {
        const float4 w = __ldg(tgt.pts + warm);
        const float d = l2_simple(x, y, z, w.x, w.y, w.z);
        if (d <= best) {
            best = d;
            best_idx = __float_as_int(w.w);
            best_pos = warm;
        }
    }
    nn_search(x, y, z, tgt, best, best_idx, best_pos);
    pos[s] = best_pos;
    if (best_pos < 0) return;
    const double *C1 = cov_src + 9 * (size_t) s, *C2 = cov_tgt + 9 * (size_t) best_pos;
    double R[9], M[9], t[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = kc->R[i];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) M[3 * r + c] = (R[3 * r] * C1[c] + R[3 * r + 1] * C1[3 + c]) + R[3 * r + 2] * C1[6 + c];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            t[3 * r + c] = ((M[3 * r] * R[3 * c] + M[3 * r + 1] * R[3 * c + 1]) + M[3 * r + 2] * R[3 * c + 2]) + C2[3 * r + c];
    // M = temp^-1 (cofactors)
    const double a = t[0], b = t[1], c = t[2], d = t[3], e = t[4], f = t[5], g = t[6], h = t[7], i = t[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    // stored entry-major (nine arrays of `ld` doubles): the cost kernel's loads of one entry are then
    // contiguous across the lanes of a warp instead of 72 bytes apart
    double *out = mahal + (size_t) s;
    out[0 * ld] = (e * i - f * h) * id;
    out[1 * ld] = (c * h - b * i) * id;
    out[2 * ld] = (b * f - c * e) * id;
    out[3 * ld] = (f * g - d * i) * id;
    out[4 * ld] = (a * i - c * g) * id;
    out[5 * ld] = (c * d - a * f) * id;
    out[6 * ld] = (d * h - e * g) * id;
    out[7 * ld] = (b * g - a * h) * id;
    out[8 * ld] = (a * e - b * d) * id;
}

// f, g_t, Racc of OptimizationFunctorWithIndices::fdf for the transform T (fp32)
__global__ void __launch_bounds__(kGicpThreads) gicp_cost_kernel(const float4 *__restrict__ src_sorted, int n_src,
                                                                 const float4 *__restrict__ tgt_sorted,
                                                                 const int *__restrict__ pos,
                                                                 const double *__restrict__ mahal, size_t ld, GicpPose pose,
                                                                 Acc128 *acc128, unsigned *ticket,
                                                                 volatile double *host_sums, volatile int *host_seq,
                                                                 int seq) {
    __shared__ float T[12];
    if (threadIdx.x < 12) T[threadIdx.x] = pose.T[threadIdx.x];
    __syncthreads();
    // Exact sums (same spec as the oracle's fdf and as the ICP estimator): every term is
    // rounded once to a multiple of 2^-k, clamped to +-2^52 units, and added as an integer - per thread
    // (<= 16 terms: the grid grows with the cloud), per warp, per block in 64 bits, over the blocks in
    // 128 bits.  Integer addition is associative, so the result is independent of the Morton order and
    // of which block finishes last, and equal to the oracle's bit for bit.
    long long acc[kCostVals];
#pragma unroll
    for (int i = 0; i < kCostVals; ++i) acc[i] = 0;
    const double scale = pose.scale;
    auto fix = [&](double v) -> long long {
        v = v * scale;
        v = fmin(fmax(v, -4503599627370496.0), 4503599627370496.0);
        return __double2ll_rn(v);
    };
    // Four pairs per thread and trip: the match positions first, then every load of the four pairs
    // (source point, gathered target point, nine Mahalanobis entries) before any arithmetic - the pass is
    // latency bound (the operands sit in L2 across the evaluations of a match), so what counts is how many
    // independent loads a warp has in flight.
    constexpr int kU = WCU_GICP_U;
    const int stride = gridDim.x * blockDim.x;
    for (int s0 = blockIdx.x * blockDim.x + threadIdx.x; s0 < n_src; s0 += kU * stride) {
        int j[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int s = s0 + u * stride;
            j[u] = s < n_src ? __ldg(pos + s) : -1;
        }
        float4 p[kU], q[kU];
        double M[kU][9];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const int s = s0 + u * stride;
            if (j[u] >= 0) {
                p[u] = __ldg(src_sorted + s);
                q[u] = __ldg(tgt_sorted + j[u]);
                const double *Mp = mahal + (size_t) s;
#pragma unroll
                for (int k = 0; k < 9; ++k) M[u][k] = __ldg(Mp + (size_t) k * ld);
            }
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (j[u] < 0) continue;
            const float px = xform_row(T + 0, p[u].x, p[u].y, p[u].z), py = xform_row(T + 4, p[u].x, p[u].y, p[u].z),
                        pz = xform_row(T + 8, p[u].x, p[u].y, p[u].z);
            const double r0 = (double) __fsub_rn(px, q[u].x), r1 = (double) __fsub_rn(py, q[u].y),
                         r2 = (double) __fsub_rn(pz, q[u].z);
            const double t0 = (M[u][0] * r0 + M[u][1] * r1) + M[u][2] * r2, t1 = (M[u][3] * r0 + M[u][4] * r1) + M[u][5] * r2,
                         t2 = (M[u][6] * r0 + M[u][7] * r1) + M[u][8] * r2;
            acc[0] += fix((r0 * t0 + r1 * t1) + r2 * t2);
            acc[1] += fix(t0);
            acc[2] += fix(t1);
            acc[3] += fix(t2);
            const double b0 = p[u].x, b1 = p[u].y, b2 = p[u].z;  // base_transformation_ (identity) * p_src
            acc[4] += fix(b0 * t0); acc[5] += fix(b0 * t1); acc[6] += fix(b0 * t2);
            acc[7] += fix(b1 * t0); acc[8] += fix(b1 * t1); acc[9] += fix(b1 * t2);
            acc[10] += fix(b2 * t0); acc[11] += fix(b2 * t1); acc[12] += fix(b2 * t2);
            acc[13] += 1;
        }
    }
    __shared__ long long s_red[kGicpThreads / 32][kCostVals];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kCostVals; ++i) {
        long long v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    // block totals go to one of kCostSlots replicas of 128-bit accumulators (64-bit atomics, the carry
    // derived from what each add observed - exact and order independent); the last block to arrive adds the
    // replicas, clears them for the next evaluation and hands the 14 sums to the host through mapped memory:
    // one launch and no copy per BFGS evaluation, of which a match makes several hundred.
    if (threadIdx.x < kCostVals) {
        long long v = 0;
        for (int w = 0; w < kGicpThreads / 32; ++w) v += s_red[w][threadIdx.x];
        if (v != 0) atomic_add128(acc128 + (blockIdx.x % kCostSlots) * 16 + threadIdx.x, (unsigned long long) v, v < 0 ? -1LL : 0LL);
    }
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < kCostVals) {
        __int128 v = 0;
        for (int sl = 0; sl < kCostSlots; ++sl) {
            Acc128 &c = acc128[sl * 16 + threadIdx.x];
            const unsigned long long lo = __ldcg(&c.lo);
            const long long hi = __ldcg(&c.hi);
            v += ((__int128) hi << 64) + (__int128) lo;
            c.lo = 0;
            c.hi = 0;
        }
        host_sums[threadIdx.x] = threadIdx.x == 13 ? (double) (long long) v
                                                   : acc_to_double((unsigned long long) v, (long long) (v >> 64), pose.k);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0u;
        *host_seq = seq;
    }
}

// covariances from sorted order back to original point order (test hook)
__global__ void gicp_unsort_cov_kernel(const float4 *__restrict__ sorted, const double *__restrict__ covs, int n,
                                       double *out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int orig = __float_as_int(sorted[s].w);
    if (orig == 0x7fffffff) return;
    for (int i = 0; i < 9; ++i) out[9 * (size_t) orig + i] = covs[9 * (size_t) s + i];
}

// ---- host: the state parametrisation and the BFGS minimiser -------------------------------------------
// applyState: rotation block <- Rz(x5) Ry(x4) Rx(x3) * rotation block (fp32), translation += x0..2
void apply_state(float *T, const double x[6]) {
    const float rx = static_cast<float>(x[3]), ry = static_cast<float>(x[4]), rz = static_cast<float>(x[5]);
    const float cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz),
                sz = std::sin(rz);
    const float R[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
                        sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
                        -sy,     cy * sx,                cy * cx};
    float out[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            float acc = R[3 * r + 0] * T[0 * 4 + c];
            acc = R[3 * r + 1] * T[1 * 4 + c] + acc;
            acc = R[3 * r + 2] * T[2 * 4 + c] + acc;
            out[3 * r + c] = acc;
        }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[4 * r + c] = out[3 * r + c];
    T[3] += static_cast<float>(x[0]);
    T[7] += static_cast<float>(x[1]);
    T[11] += static_cast<float>(x[2]);
}

void identity4(float *T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}

// computeRDerivative with matricesInnerProd(m1, m2) = sum_ij m1(j,i) m2(i,j), as written in gicp.hpp
void rotation_gradient(const double x[6], const double R[9], double g[6]) {
    const double cphi = std::cos(x[3]), sphi = std::sin(x[3]), ctheta = std::cos(x[4]), stheta = std::sin(x[4]),
                 cpsi = std::cos(x[5]), spsi = std::sin(x[5]);
    const double dPhi[9] = {0., sphi * spsi + cphi * cpsi * stheta, cphi * spsi - cpsi * sphi * stheta,
                            0., -cpsi * sphi + cphi * spsi * stheta, -cphi * cpsi - sphi * spsi * stheta,
                            0., cphi * ctheta, -ctheta * sphi};
    const double dTheta[9] = {-cpsi * stheta, cpsi * ctheta * sphi, cphi * cpsi * ctheta,
                              -spsi * stheta, ctheta * sphi * spsi, cphi * ctheta * spsi,
                              -ctheta, -sphi * stheta, -cphi * stheta};
    const double dPsi[9] = {-ctheta * spsi, -cphi * cpsi - sphi * spsi * stheta, cpsi * sphi - cphi * spsi * stheta,
                            cpsi * ctheta, -cphi * spsi + cpsi * sphi * stheta, sphi * spsi + cphi * cpsi * stheta,
                            0., 0., 0.};
    auto inner = [](const double *m1, const double *m2) {
        double r = 0.;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r += m1[3 * j + i] * m2[3 * i + j];
        return r;
    };
    g[3] = inner(dPhi, R);
    g[4] = inner(dTheta, R);
    g[5] = inner(dPsi, R);
}

double eval_poly(const double *c, int n, double x) {  // Eigen::poly_eval
    if (x * x <= 1.0) {
        double val = c[n - 1];
        for (int i = n - 2; i >= 0; --i) val = val * x + c[i];
        return val;
    }
    double val = c[0];
    const double inv_x = 1.0 / x;
    for (int i = 1; i < n; ++i) val = val * inv_x + c[i];
    return std::pow(x, (double) (n - 1)) * val;
}

}  // namespace

struct GicpHandle {
    wavecu_gicp_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    TargetIndex src, tgt;       // both clouds carry an LBVH: the covariances need k-NN in each
    VoxelWork vox;
    float4 *d_stage = nullptr;  // unfiltered upload when res > 0
    size_t stage_cap = 0;
    double *d_cov_src = nullptr, *d_cov_tgt = nullptr, *d_mahal = nullptr, *d_cov_out = nullptr;
    int *d_pos = nullptr;
    size_t src_cap = 0, tgt_cap = 0;
    bool cov_src_ok = false, cov_tgt_ok = false;
    GicpIterConsts *d_consts = nullptr;
    Acc128 *d_partial = nullptr;      // kCostSlots x 16 fixed-point accumulators of the cost kernel
    double *h_sums = nullptr;         // mapped host memory
    int sum_k = 28;                   // fixed-point exponent of the cost sums of the current match
    unsigned *d_ticket = nullptr;
    int *h_seq = nullptr;   // mapped: number of the last evaluation whose sums are in h_sums
    int cost_seq = 0;
    int n_blocks = 148 * 4;
    long long launches = 0, evaluations = 0, inner_iterations = 0;
    size_t n_corr = 0;
    // optional per-kernel timing (CUDA events on the handle's stream around every cost-kernel launch)
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double cost_ms = 0;
    long long cost_launches = 0;

    // ---- functor state (the pairs of the current outer iteration) ----
    long long m_pairs = 0;

    int init() {
        WCU_CHECK(cudaSetDevice(device));
        if (!stream) {
            WCU_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            own_stream = true;
        }
        src.cloud.device = tgt.cloud.device = vox.device = device;
        src.cloud.stream = tgt.cloud.stream = vox.stream = stream;
        WCU_CHECK(cudaMalloc((void **) &d_consts, sizeof(GicpIterConsts)));
        WCU_CHECK(cudaHostAlloc((void **) &h_sums, sizeof(double) * kCostVals, cudaHostAllocMapped));
        WCU_CHECK(cudaHostAlloc((void **) &h_seq, sizeof(int), cudaHostAllocMapped));
        *h_seq = 0;
        WCU_CHECK(cudaMalloc((void **) &d_ticket, sizeof(unsigned)));
        WCU_CHECK(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned), stream));
        return WAVECU_OK;
    }

    // setRef / setTarget: voxel filter first when res > 0 (src/gicp.cpp:37-55)
    int set_cloud(TargetIndex &dst, bool &cov_ok, const float *xyzw, size_t n, bool from_device) {
        WCU_CHECK(cudaSetDevice(device));
        cov_ok = false;
        if (!(prm.res > 0) || n == 0) return dst.set_points(xyzw, n, from_device);
        if (n > stage_cap) {
            if (d_stage) WCU_CHECK(cudaFree(d_stage));
            d_stage = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_stage, (n + 64) * sizeof(float4)));
            stage_cap = n + 64;
        }
        WCU_CHECK(cudaMemcpyAsync(d_stage, xyzw, n * sizeof(float4),
                                  from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
        int rc = dst.cloud.reserve(n, 0);
        if (rc) return rc;
        size_t n_out = 0;
        rc = vox.filter(d_stage, n, prm.res, dst.cloud.d_raw, &n_out, nullptr);
        if (rc) return rc;
        dst.cloud.n = n_out;
        dst.nrm_n = 0;
        dst.dirty = true;
        return WAVECU_OK;
    }

    int ensure_buffers() {
        const size_t ns = std::max<size_t>(src.cloud.n, 1), nt = std::max<size_t>(tgt.cloud.n, 1);
        if (ns > src_cap) {
            for (void *p : {(void *) d_cov_src, (void *) d_mahal, (void *) d_pos})
                if (p) WCU_CHECK(cudaFree(p));
            d_cov_src = d_mahal = nullptr;
            d_pos = nullptr;
            const size_t a = ns + ns / 8 + 64;
            WCU_CHECK(cudaMalloc((void **) &d_cov_src, a * 9 * sizeof(double)));
            WCU_CHECK(cudaMalloc((void **) &d_mahal, a * 9 * sizeof(double)));
            WCU_CHECK(cudaMalloc((void **) &d_pos, a * sizeof(int)));
            src_cap = a;
            cov_src_ok = false;
        }
        if (nt > tgt_cap) {
            if (d_cov_tgt) WCU_CHECK(cudaFree(d_cov_tgt));
            d_cov_tgt = nullptr;
            const size_t a = nt + nt / 8 + 64;
            WCU_CHECK(cudaMalloc((void **) &d_cov_tgt, a * 9 * sizeof(double)));
            tgt_cap = a;
            cov_tgt_ok = false;
        }
        return WAVECU_OK;
    }

    int prepare() {  // trees + covariances of whatever changed
        int rc = ensure_buffers();
        if (rc) return rc;
        const int k = std::max(1, std::min(prm.corr_rand, kMaxKnn));
        if (tgt.dirty) {
            rc = tgt.build();
            if (rc) return rc;
            cov_tgt_ok = false;
        }
        if (src.dirty) {
            rc = src.build();
            if (rc) return rc;
            cov_src_ok = false;
        }
        if (!cov_tgt_ok && tgt.cloud.n) {
            const int n = (int) tgt.cloud.n;
            if ((size_t) k > tgt.cloud.n) WCU_CHECK(cudaMemsetAsync(d_cov_tgt, 0, (size_t) n * 9 * sizeof(double), stream));
            else gicp_cov_kernel<<<(n + kGicpThreads - 1) / kGicpThreads, kGicpThreads, 0, stream>>>(tgt.index(), n, k, 1e-3, d_cov_tgt);
            ++launches;
            cov_tgt_ok = true;
        }
        if (!cov_src_ok && src.cloud.n) {
            const int n = (int) src.cloud.n;
            if ((size_t) k > src.cloud.n) WCU_CHECK(cudaMemsetAsync(d_cov_src, 0, (size_t) n * 9 * sizeof(double), stream));
            else gicp_cov_kernel<<<(n + kGicpThreads - 1) / kGicpThreads, kGicpThreads, 0, stream>>>(src.index(), n, k, 1e-3, d_cov_src);
            ++launches;
            cov_src_ok = true;
        }
        WCU_CHECK(cudaGetLastError());
        return WAVECU_OK;
    }

    // OptimizationFunctorWithIndices::fdf at state x (base_transformation_ = identity)
    int cost(const double x[6], double *f, double *g) {
        float T[16];
        identity4(T);
        apply_state(T, x);
        GicpPose pose;
        std::memcpy(pose.T, T, sizeof T);
        pose.k = sum_k;
        pose.scale = std::ldexp(1.0, sum_k);
        ++evaluations;
        const int seq = ++cost_seq;
        if (profiling) {
            if (!ev0) {
                WCU_CHECK(cudaEventCreate(&ev0));
                WCU_CHECK(cudaEventCreate(&ev1));
            }
            WCU_CHECK(cudaEventRecord(ev0, stream));
        }
        gicp_cost_kernel<<<n_blocks, kGicpThreads, 0, stream>>>(src.cloud.d_sorted, (int) src.cloud.n, tgt.cloud.d_sorted,
                                                                d_pos, d_mahal, src_cap, pose, d_partial, d_ticket, h_sums,
                                                                h_seq, seq);
        if (profiling) WCU_CHECK(cudaEventRecord(ev1, stream));
        ++launches;
        WCU_CHECK(cudaGetLastError());
        // wait for the kernel's own hand-over instead of synchronising the stream
        for (int spins = 0; *(volatile int *) h_seq != seq;) {
            if (++spins % 4096 == 0) {
                const cudaError_t q = cudaStreamQuery(stream);
                if (q != cudaErrorNotReady) {
                    if (q != cudaSuccess) WCU_CHECK(q);
                    if (*(volatile int *) h_seq != seq) {
                        set_last_error("gicp_cost_kernel finished without publishing its sums");
                        return WAVECU_ERR_CUDA;
                    }
                }
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if (spins > 2048 && (spins & 255) == 255) sched_yield();   // see icp.cu: fewer cores than matcher threads
        }
        if (profiling) {
            float ms = 0;
            WCU_CHECK(cudaEventSynchronize(ev1));
            WCU_CHECK(cudaEventElapsedTime(&ms, ev0, ev1));
            cost_ms += ms;
            ++cost_launches;
        }
        const double m = (double) m_pairs;
        if (f) *f = h_sums[0] / m;
        if (g) {
            double R[9];
            for (int d = 0; d < 3; ++d) g[d] = h_sums[1 + d] * (2.0 / m);
            for (int q = 0; q < 9; ++q) R[q] = h_sums[4 + q] * (2.0 / m);
            rotation_gradient(x, R, g);
        }
        return WAVECU_OK;
    }

    int match(double *T_out, int *converged_out, int *iterations_out);

    void release() {
        cudaSetDevice(device);
        src.release();
        tgt.release();
        vox.release();
        for (void *p : {(void *) d_stage, (void *) d_cov_src, (void *) d_cov_tgt, (void *) d_mahal, (void *) d_cov_out,
                        (void *) d_pos, (void *) d_consts, (void *) d_partial})
            if (p) cudaFree(p);
        if (h_sums) cudaFreeHost(h_sums);
        if (h_seq) cudaFreeHost(h_seq);
        if (d_ticket) cudaFree(d_ticket);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace {

// pcl::BFGS over GicpHandle::cost - same control flow as pcl/registration/bfgs.h
class BfgsMinimizer {
 public:
    enum Status { NegativeGradientEpsilon = -3, NotStarted = -2, Running = -1, Success = 0, NoProgress = 1, Failed = 99 };
    explicit BfgsMinimizer(GicpHandle &h) : h_(h) {}
    int error = WAVECU_OK;

    void init(const double x[6]) {
        delta_f_ = 0;
        eval(x, &f_, grad_);
        copy(x0_, x);
        copy(g0_, grad_);
        g0norm_ = norm(g0_);
        for (int i = 0; i < 6; ++i) p_[i] = grad_[i] * -1 / g0norm_;
        pnorm_ = norm(p_);
        fp0_ = -g0norm_;
        copy(x_alpha_, x0_);
        copy(g_alpha_, g0_);
        f_alpha_ = f_;
        x_key_ = f_key_ = g_key_ = df_key_ = 0;
        df_alpha_ = dot(g_alpha_, p_);
    }

    Status step(double x[6]) {
        double alpha = 0.0, alpha1;
        const double f0 = f_;
        if (pnorm_ == 0.0 || g0norm_ == 0.0 || fp0_ == 0) return NoProgress;
        if (delta_f_ < 0) {
            const double del = std::max(-delta_f_, 10 * std::numeric_limits<double>::epsilon() * std::fabs(f0));
            alpha1 = std::min(1.0, 2.0 * del / (-fp0_));
        } else {
            alpha1 = 1.0;  // |parameters.step_size|
        }
        const Status st = line_search(alpha1, alpha);
        if (error) return Failed;
        if (st != Success) return st;
        double fa, dfa;
        apply_fdf(alpha, fa, dfa);  // updatePosition
        f_ = f_alpha_;
        copy(x, x_alpha_);
        copy(grad_, g_alpha_);
        delta_f_ = f_ - f0;
        double dx0[6], dg0[6];
        for (int i = 0; i < 6; ++i) {
            dx0[i] = x[i] - x0_[i];
            dg0[i] = grad_[i] - g0_[i];
        }
        const double dxg = dot(dx0, grad_), dgg = dot(dg0, grad_), dxdg = dot(dx0, dg0), dgnorm = norm(dg0);
        double A = 0, B = 0;
        if (dxdg != 0) {
            B = dxg / dxdg;
            A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
        }
        for (int i = 0; i < 6; ++i) {
            p_[i] = -A * dx0[i];
            p_[i] += grad_[i];
            p_[i] += -B * dg0[i];
        }
        copy(g0_, grad_);
        copy(x0_, x);
        g0norm_ = norm(g0_);
        pnorm_ = norm(p_);
        const double dir = (dot(p_, grad_) > 0) ? -1.0 : 1.0;
        for (int i = 0; i < 6; ++i) p_[i] *= dir / pnorm_;
        pnorm_ = norm(p_);
        fp0_ = dot(p_, g0_);
        copy(x_alpha_, x0_);  // changeDirection
        copy(g_alpha_, g0_);
        x_key_ = f_key_ = g_key_ = df_key_ = 0.0;
        df_alpha_ = dot(g_alpha_, p_);
        return Success;
    }

    Status test_gradient(double eps) const {
        if (eps < 0) return NegativeGradientEpsilon;
        return norm(grad_) < eps ? Success : Running;
    }

 private:
    static void copy(double *d, const double *s) { std::memcpy(d, s, 6 * sizeof(double)); }
    static double dot(const double *a, const double *b) {
        double s = 0;
        for (int i = 0; i < 6; ++i) s += a[i] * b[i];
        return s;
    }
    static double norm(const double *a) { return std::sqrt(dot(a, a)); }
    void eval(const double *x, double *f, double *g) {
        const int rc = h_.cost(x, f, g);
        if (rc) error = rc;
    }
    void move_to(double alpha) {
        for (int i = 0; i < 6; ++i) x_alpha_[i] = x0_[i] + alpha * p_[i];
        x_key_ = alpha;
    }
    double apply_f(double alpha) {
        if (alpha == f_key_) return f_alpha_;
        move_to(alpha);
        eval(x_alpha_, &f_alpha_, nullptr);
        f_key_ = alpha;
        return f_alpha_;
    }
    double apply_df(double alpha) {
        if (alpha == df_key_) return df_alpha_;
        move_to(alpha);
        if (alpha != g_key_) {
            eval(x_alpha_, nullptr, g_alpha_);
            g_key_ = alpha;
        }
        df_alpha_ = dot(g_alpha_, p_);
        df_key_ = alpha;
        return df_alpha_;
    }
    void apply_fdf(double alpha, double &f, double &df) {
        if (alpha == f_key_ && alpha == df_key_) {
            f = f_alpha_;
            df = df_alpha_;
            return;
        }
        if (alpha == f_key_ || alpha == df_key_) {
            f = apply_f(alpha);
            df = apply_df(alpha);
            return;
        }
        move_to(alpha);
        eval(x_alpha_, &f_alpha_, g_alpha_);
        f_key_ = g_key_ = alpha;
        df_alpha_ = dot(g_alpha_, p_);
        df_key_ = alpha;
        f = f_alpha_;
        df = df_alpha_;
    }

    // minimiser of the interpolating cubic (order 3, both slopes known) or quadratic on [xmin, xmax]
    static double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin,
                              double xmax, int order) {
        double y, ymin = (xmin - a) / (b - a), ymax = (xmax - a) / (b - a), fmin;
        if (ymin > ymax) std::swap(ymin, ymax);
        if (order > 2 && !(fpb != fpb) && fpb != std::numeric_limits<double>::infinity()) {
            fpa = fpa * (b - a);
            fpb = fpb * (b - a);
            const double eta = 3 * (fb - fa) - 2 * fpa - fpb, xi = fpa + fpb - 2 * (fb - fa);
            const double c[4] = {fa, fpa, eta, xi};
            y = ymin;
            fmin = eval_poly(c, 4, ymin);
            auto consider = [&](double t) {
                const double v = eval_poly(c, 4, t);
                if (v < fmin) {
                    y = t;
                    fmin = v;
                }
            };
            consider(ymax);
            // derivative c1 + 2 c2 t + 3 c3 t^2: closed-form roots (PolynomialSolver<Scalar, 2> in bfgs.h)
            const double q0 = c[1], q1 = 2 * c[2], q2 = 3 * c[3], a2 = 2 * q2, disc = q1 * q1 - 4 * q0 * q2;
            if (0 < disc) {
                const double sq = std::sqrt(disc);
                double y0 = (-q1 - sq) / a2, y1 = (-q1 + sq) / a2;
                if (y0 > y1) std::swap(y0, y1);
                if (y0 > ymin && y0 < ymax) consider(y0);
                if (y1 > ymin && y1 < ymax) consider(y1);
            } else if (0 == disc) {
                const double y0 = -q1 / a2;
                if (y0 > ymin && y0 < ymax) consider(y0);
            }
        } else {
            fpa = fpa * (b - a);
            const double fl = fa + ymin * (fpa + ymin * (fb - fa - fpa));
            const double fh = fa + ymax * (fpa + ymax * (fb - fa - fpa));
            const double curv = 2 * (fb - fa - fpa);
            y = ymin;
            fmin = fl;
            if (fh < fmin) {
                y = ymax;
                fmin = fh;
            }
            if (curv > a) {  // sic: bfgs.h compares the curvature with a
                const double z = -fpa / curv;
                if (z > ymin && z < ymax) {
                    const double fz = fa + z * (fpa + z * (fb - fa - fpa));
                    if (fz < fmin) {
                        y = z;
                        fmin = fz;
                    }
                }
            }
        }
        return a + y * (b - a);
    }

    Status line_search(double alpha1, double &alpha_new) {
        const double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5;
        const int order = 3, bracket_iters = 100, section_iters = 100;
        double f0, fp0, falpha, falpha_prev, fpalpha, fpalpha_prev, delta, alpha_next;
        double alpha = alpha1, alpha_prev = 0.0, a = 0.0, b = alpha, fa, fb = 0.0, fpa, fpb = 0.0;
        int i = 0;
        apply_fdf(0.0, f0, fp0);
        falpha_prev = f0;
        fpalpha_prev = fp0;
        fa = f0;
        fpa = fp0;
        while (i++ < bracket_iters) {  // bracketing
            falpha = apply_f(alpha);
            if (error) return Failed;
            if (falpha > f0 + alpha * rho * fp0 || falpha >= falpha_prev) {  // Fletcher's rho test
                a = alpha_prev; fa = falpha_prev; fpa = fpalpha_prev;
                b = alpha; fb = falpha; fpb = std::numeric_limits<double>::quiet_NaN();
                break;
            }
            fpalpha = apply_df(alpha);
            if (error) return Failed;
            if (std::fabs(fpalpha) <= -sigma * fp0) {  // Fletcher's sigma test
                alpha_new = alpha;
                return Success;
            }
            if (fpalpha >= 0) {
                a = alpha; fa = falpha; fpa = fpalpha;
                b = alpha_prev; fb = falpha_prev; fpb = fpalpha_prev;
                break;
            }
            delta = alpha - alpha_prev;
            alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, alpha + delta,
                                     alpha + tau1 * delta, order);
            alpha_prev = alpha;
            falpha_prev = falpha;
            fpalpha_prev = fpalpha;
            alpha = alpha_next;
        }
        while (i++ < section_iters) {  // sectioning of [a, b]
            delta = b - a;
            alpha = interpolate(a, fa, fpa, b, fb, fpb, a + tau2 * delta, b - tau3 * delta, order);
            falpha = apply_f(alpha);
            if (error) return Failed;
            if ((a - alpha) * fpa <= std::numeric_limits<double>::epsilon()) return NoProgress;  // roundoff
            if (falpha > f0 + rho * alpha * fp0 || falpha >= fa) {
                b = alpha; fb = falpha; fpb = std::numeric_limits<double>::quiet_NaN();
            } else {
                fpalpha = apply_df(alpha);
                if (error) return Failed;
                if (std::fabs(fpalpha) <= -sigma * fp0) {
                    alpha_new = alpha;
                    return Success;
                }
                if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
                    b = a; fb = fa; fpb = fpa;
                }
                a = alpha; fa = falpha; fpa = fpalpha;
            }
        }
        return Success;
    }

    GicpHandle &h_;
    double f_ = 0, delta_f_ = 0, fp0_ = 0, pnorm_ = 0, g0norm_ = 0;
    double x_key_ = 0, f_key_ = 0, g_key_ = 0, df_key_ = 0, f_alpha_ = 0, df_alpha_ = 0;
    double grad_[6], x0_[6], g0_[6], p_[6], x_alpha_[6], g_alpha_[6];
};

}  // namespace

int GicpHandle::match(double *T_out, int *converged_out, int *iterations_out) {
    WCU_CHECK(cudaSetDevice(device));
    launches = evaluations = inner_iterations = 0;
    n_corr = 0;
    cost_ms = 0;
    cost_launches = 0;
    float transformation[16], previous[16];
    identity4(transformation);
    identity4(previous);
    bool converged = false;
    int nr_iterations = 0;
    const size_t n_src = src.cloud.n, n_tgt = tgt.cloud.n;
    if (n_src && n_tgt) {
        int rc = prepare();
        if (rc) return rc;
        // cost-kernel grid: WCU_GICP_PPT pairs per thread
        n_blocks = (int) std::max<size_t>(1, (n_src + (size_t) kGicpThreads * WCU_GICP_PPT - 1) / ((size_t) kGicpThreads * WCU_GICP_PPT));
        if (!d_partial) {
            WCU_CHECK(cudaMalloc((void **) &d_partial, sizeof(Acc128) * kCostSlots * 16));
            WCU_CHECK(cudaMemsetAsync(d_partial, 0, sizeof(Acc128) * kCostSlots * 16, stream));
        }
        {   // fixed-point exponent of the cost sums from the source extent (oracle: gicp_sum_exponent)
            unsigned bb[6];
            WCU_CHECK(cudaMemcpyAsync(bb, src.cloud.d_bbox, sizeof bb, cudaMemcpyDeviceToHost, stream));
            WCU_CHECK(cudaStreamSynchronize(stream));
            double bmax = 0;
            for (int d = 0; d < 6; ++d) {
                const unsigned u = bb[d];
                const unsigned bits = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;   // ordered_to_float
                float v;
                std::memcpy(&v, &bits, sizeof v);
                if (std::isfinite(v)) bmax = std::max(bmax, (double) std::fabs(v));
            }
            int e;
            std::frexp(std::max(bmax, 64.0) * 32768.0, &e);
            sum_k = 50 - e;
        }
        WCU_CHECK(cudaMemsetAsync(d_pos, 0xff, n_src * sizeof(int), stream));
        const double dist_threshold = 5.0 * 5.0;  // corr_dist_threshold_ (libwave never sets it)
        float thr = (float) dist_threshold;
        if (!((double) thr < dist_threshold)) thr = std::nextafterf(thr, -INFINITY);
        const double rotation_epsilon = prm.r_eps, transformation_epsilon = 5e-4;
        const int max_inner = 20;
        int *h_count = nullptr;
        WCU_CHECK(cudaHostAlloc((void **) &h_count, sizeof(int), cudaHostAllocDefault));
        while (!converged) {
            GicpIterConsts c;
            std::memcpy(c.T, transformation, sizeof(float) * 12);
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q) c.R[3 * r + q] = (double) transformation[4 * r + q];
            c.thr = thr;
            WCU_CHECK(cudaMemcpyAsync(d_consts, &c, sizeof c, cudaMemcpyHostToDevice, stream));
            gicp_corr_kernel<<<(unsigned) ((n_src + kGicpThreads - 1) / kGicpThreads), kGicpThreads, 0, stream>>>(
                src.cloud.d_sorted, (int) n_src, tgt.index(), d_cov_src, d_cov_tgt, d_consts, d_pos, d_mahal, src_cap);
            ++launches;
            std::memcpy(previous, transformation, sizeof previous);
            // the pair count comes with the first cost evaluation (slot 13)
            double x[6];
            x[0] = transformation[3];
            x[1] = transformation[7];
            x[2] = transformation[11];
            x[3] = std::atan2(transformation[9], transformation[10]);
            x[4] = std::asin(-transformation[8]);
            x[5] = std::atan2(transformation[4], transformation[0]);
            m_pairs = 1;
            rc = cost(x, nullptr, nullptr);
            if (rc) break;
            --evaluations;
            m_pairs = (long long) (h_sums[13] + 0.5);
            n_corr = (size_t) m_pairs;
            if (m_pairs < 4) break;  // NotEnoughPointsException -> caught: converged_ stays false
            BfgsMinimizer bfgs(*this);
            bfgs.init(x);
            int inner = 0;
            BfgsMinimizer::Status result = BfgsMinimizer::Running;
            do {
                ++inner;
                result = bfgs.step(x);
                if (result) break;
                result = bfgs.test_gradient(1e-2);
            } while (result == BfgsMinimizer::Running && inner < max_inner);
            inner_iterations += inner;
            if (bfgs.error) {
                rc = bfgs.error;
                break;
            }
            if (!(result == BfgsMinimizer::NoProgress || result == BfgsMinimizer::Success || inner == max_inner)) break;
            identity4(transformation);
            apply_state(transformation, x);
            double delta = 0.;
            for (int k = 0; k < 4; k++)
                for (int l = 0; l < 4; l++) {
                    const double ratio = (k < 3 && l < 3) ? 1. / rotation_epsilon : 1. / transformation_epsilon;
                    const double c_delta = ratio * std::fabs(previous[4 * k + l] - transformation[4 * k + l]);
                    if (c_delta > delta) delta = c_delta;
                }
            nr_iterations++;
            if (nr_iterations >= prm.max_iter || delta < 1) {
                converged = true;
                std::memcpy(previous, transformation, sizeof previous);
            }
        }
        cudaFreeHost(h_count);
        if (rc) return rc;
    }
    if (T_out)
        for (int i = 0; i < 16; ++i) T_out[i] = (double) previous[i];
    if (converged_out) *converged_out = converged ? 1 : 0;
    if (iterations_out) *iterations_out = nr_iterations;
    return WAVECU_OK;
}

}  // namespace wavecu

using namespace wavecu;

struct wavecu_gicp {
    GicpHandle h;
};

extern "C" {

void wavecu_gicp_default_params(wavecu_gicp_params *p) {
    if (!p) return;
    p->corr_rand = 10;
    p->max_iter = 100;
    p->r_eps = 1e-8;
    p->fit_eps = 1e-2;
    p->res = 0.1f;
}

int wavecu_gicp_create(const wavecu_gicp_params *params, int device, void *stream, wavecu_gicp **out) {
    if (!out) return WAVECU_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_last_error("no such CUDA device (libwavecu has no CPU fallback)");
        return WAVECU_ERR_CUDA;
    }
    wavecu_gicp *w = new wavecu_gicp();
    if (params) w->h.prm = *params;
    else wavecu_gicp_default_params(&w->h.prm);
    w->h.device = device;
    w->h.stream = (cudaStream_t) stream;
    const int rc = w->h.init();
    if (rc) {
        w->h.release();
        delete w;
        return rc;
    }
    *out = w;
    return WAVECU_OK;
}

int wavecu_gicp_destroy(wavecu_gicp *w) {
    if (!w) return WAVECU_OK;
    w->h.release();
    delete w;
    return WAVECU_OK;
}

int wavecu_gicp_set_params(wavecu_gicp *w, const wavecu_gicp_params *params) {
    if (!w || !params) return WAVECU_ERR_ARG;
    w->h.prm = *params;
    return WAVECU_OK;
}

int wavecu_gicp_set_source(wavecu_gicp *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    return w->h.set_cloud(w->h.src, w->h.cov_src_ok, xyzw, n, false);
}
int wavecu_gicp_set_target(wavecu_gicp *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    return w->h.set_cloud(w->h.tgt, w->h.cov_tgt_ok, xyzw, n, false);
}
int wavecu_gicp_set_source_device(wavecu_gicp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    return w->h.set_cloud(w->h.src, w->h.cov_src_ok, (const float *) d, n, true);
}
int wavecu_gicp_set_target_device(wavecu_gicp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    return w->h.set_cloud(w->h.tgt, w->h.cov_tgt_ok, (const float *) d, n, true);
}

int wavecu_gicp_match(wavecu_gicp *w, double T_out[16], int *converged, int *iterations) {
    if (!w) return WAVECU_ERR_ARG;
    return w->h.match(T_out, converged, iterations);
}

int wavecu_gicp_covariances(wavecu_gicp *w, int which, double *covs9, size_t *n) {
    if (!w || !n) return WAVECU_ERR_ARG;
    GicpHandle &h = w->h;
    WCU_CHECK(cudaSetDevice(h.device));
    int rc = h.prepare();
    if (rc) return rc;
    TargetIndex &ti = which ? h.tgt : h.src;
    const double *d_cov = which ? h.d_cov_tgt : h.d_cov_src;
    *n = ti.cloud.n;
    if (!covs9 || ti.cloud.n == 0) return WAVECU_OK;
    const size_t cnt = ti.cloud.n;
    double *d_out = nullptr;
    WCU_CHECK(cudaMalloc((void **) &d_out, cnt * 9 * sizeof(double)));
    cudaMemsetAsync(d_out, 0, cnt * 9 * sizeof(double), h.stream);
    gicp_unsort_cov_kernel<<<(unsigned) ((cnt + 255) / 256), 256, 0, h.stream>>>(ti.cloud.d_sorted, d_cov, (int) cnt, d_out);
    cudaMemcpyAsync(covs9, d_out, cnt * 9 * sizeof(double), cudaMemcpyDeviceToHost, h.stream);
    const cudaError_t e = cudaStreamSynchronize(h.stream);
    cudaFree(d_out);
    if (e != cudaSuccess) {
        set_last_error(std::string("wavecu_gicp_covariances: ") + cudaGetErrorString(e));
        return WAVECU_ERR_CUDA;
    }
    return WAVECU_OK;
}

int wavecu_gicp_cloud(wavecu_gicp *w, int which, float *xyzw, size_t *n) {
    if (!w || !n) return WAVECU_ERR_ARG;
    GicpHandle &h = w->h;
    WCU_CHECK(cudaSetDevice(h.device));
    TargetIndex &ti = which ? h.tgt : h.src;
    *n = ti.cloud.n;
    if (!xyzw || ti.cloud.n == 0) return WAVECU_OK;
    WCU_CHECK(cudaMemcpyAsync(xyzw, ti.cloud.d_raw, ti.cloud.n * sizeof(float4), cudaMemcpyDeviceToHost, h.stream));
    WCU_CHECK(cudaStreamSynchronize(h.stream));
    return WAVECU_OK;
}

int wavecu_gicp_stats(wavecu_gicp *w, long long *kernel_launches, long long *evaluations, long long *inner_iterations,
                      size_t *n_corr) {
    if (!w) return WAVECU_ERR_ARG;
    if (kernel_launches) *kernel_launches = w->h.launches + w->h.src.cloud.launches + w->h.tgt.cloud.launches;
    if (evaluations) *evaluations = w->h.evaluations;
    if (inner_iterations) *inner_iterations = w->h.inner_iterations;
    if (n_corr) *n_corr = w->h.n_corr;
    return WAVECU_OK;
}

int wavecu_gicp_set_profiling(wavecu_gicp *w, int enabled) {
    if (!w) return WAVECU_ERR_ARG;
    w->h.profiling = enabled != 0;
    return WAVECU_OK;
}

int wavecu_gicp_timing(wavecu_gicp *w, double *cost_kernel_ms, long long *cost_kernel_launches) {
    if (!w) return WAVECU_ERR_ARG;
    if (cost_kernel_ms) *cost_kernel_ms = w->h.cost_ms;
    if (cost_kernel_launches) *cost_kernel_launches = w->h.cost_launches;
    return WAVECU_OK;
}

}  // extern "C"
