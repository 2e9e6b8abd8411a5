// NDTMatcher on the device (wavecu_ndt_*): replaces pcl::NormalDistributionsTransform as the
// reference drives it (wave_matching/src/ndt.cpp:18-65; PCL 1.8 registration/impl/ndt.hpp and
// filters/impl/voxel_grid_covariance.hpp).
//
//   grid build   VoxelGridCovariance: voxel keys + stable sort (shared with VoxelGrid), then one
//                thread per occupied voxel accumulates, in cloud order, the fp32 centroid (what PCL's
//                radius search runs on) and the fp64 sum / sum of outer products, forms the
//                covariance, clamps its small eigenvalues to 0.01 * largest (3x3 Jacobi in fp64)
//                and inverts it; voxels with >= 6 points go into an open-addressing hash table
//                keyed by voxel index.
//   derivatives  computeDerivatives: one thread per source point transforms it (fp32 4x4), probes
//                the 27 voxels around it - a centroid within `res` of the point can only live there,
//                which replaces PCL's kd-tree radius search over the centroids - and accumulates
//                score, gradient (6) and Hessian (21 unique) in fp64; per-block partial sums are
//                combined in a fixed order by a second kernel (deterministic).
//   host         Newton step through a 6x6 SVD solve and PCL's computeStepLengthMT, including its
//                PCL 1.8 behaviour that the More-Thuente loop is skipped whenever step_max > step_min.
#include <sched.h>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/wavecu.h"
#include "index.cuh"
#include "voxel.cuh"

namespace wavecu {

namespace {

struct NdtLeafDev {
    double mean[3];
    double icov[6];  // xx xy xz yy yz zz
    float centroid[3];
    int voxel;
    int n;
    int valid;
};

constexpr int kNdtThreads = 128;
constexpr int kNdtVals = 28;  // score, gradient 6, Hessian upper triangle 21

// symmetric 3x3 eigen-decomposition (cyclic Jacobi, fp64); eigenvalues ascending
__device__ void eig_sym3(const double A_in[9], double evals[3], double V[9]) {
    double A[3][3], Q[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = A_in[3 * i + j];
            Q[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double qkp = Q[k][p], qkq = Q[k][q];
                    Q[k][p] = c * qkp - s * qkq;
                    Q[k][q] = s * qkp + c * qkq;
                }
            }
    }
    int o0 = 0, o1 = 1, o2 = 2;
    if (A[o1][o1] < A[o0][o0]) { const int t = o0; o0 = o1; o1 = t; }
    if (A[o2][o2] < A[o1][o1]) { const int t = o1; o1 = o2; o2 = t; }
    if (A[o1][o1] < A[o0][o0]) { const int t = o0; o0 = o1; o1 = t; }
    const int order[3] = {o0, o1, o2};
    for (int j = 0; j < 3; ++j) {
        evals[j] = A[order[j]][order[j]];
        for (int i = 0; i < 3; ++i) V[3 * i + j] = Q[i][order[j]];
    }
}

// voxel slot -> first element of its run in the sorted arrays
__global__ void __launch_bounds__(256) ndt_starts_kernel(const unsigned *__restrict__ keys, const int *__restrict__ pos,
                                                         size_t n, int *starts) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = keys[i];
    if ((k != 0xffffffffu) && (i == 0 || keys[i - 1] != k)) starts[pos[i]] = (int) i;
}

// points in voxel-sorted order (pcl's index_vector after its sort): the leaf kernel below then streams
// contiguous memory instead of chasing vals[] -> in[] for every point
__global__ void __launch_bounds__(256) ndt_gather_kernel(const float4 *__restrict__ in, const unsigned *__restrict__ keys,
                                                         const unsigned *__restrict__ vals, size_t n, float4 *out) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (keys[i] != 0xffffffffu) out[i] = in[vals[i]];
}

// VoxelGridCovariance leaf statistics, one WARP per occupied voxel.  A map voxel near the sensor holds
// tens of thousands of points: with one thread per voxel (round 1) that single serial loop of dependent
// gathers took 8.3 ms of a 16 ms match on the 5 M-point map.  The lanes read 4 x 32 consecutive points of
// the run per trip; the fp64 sums (mean, covariance) are per-lane partials combined by a fixed shuffle
// tree; the fp32 centroid - which only decides which cells a point can reach, and which the parity tests
// compare bit for bit - keeps PCL's strictly sequential summation order: one chain of adds over the lanes'
// values in run order, fed by shuffles, so the only serial cost left is the add latency itself.
__global__ void __launch_bounds__(256) ndt_leaf_kernel(const float4 *__restrict__ pts, const unsigned *__restrict__ keys,
                                                       const int *__restrict__ starts, int n_voxels, size_t n,
                                                       NdtLeafDev *leaves) {
    const int slot = (int) ((blockIdx.x * (size_t) blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (slot >= n_voxels) return;
    const size_t start = (size_t) starts[slot];
    const unsigned k = keys[start];
    // end of the run: the start of the next voxel's, or the first element that is not this voxel's
    size_t end;
    if (slot + 1 < n_voxels) end = (size_t) starts[slot + 1];
    else {
        end = start;
        while (end < n && keys[end] == k) ++end;   // last voxel only; lanes agree
    }
    float cx = 0.f, cy = 0.f, cz = 0.f;
    double s[3] = {0, 0, 0}, ss[6] = {0, 0, 0, 0, 0, 0};  // xx xy xz yy yz zz
    const size_t cnt_total = end - start;
    constexpr int kTrip = 4;
    for (size_t base = start; base < end; base += 32 * kTrip) {
        float4 p[kTrip];
#pragma unroll
        for (int u = 0; u < kTrip; ++u) {
            const size_t j = base + (size_t) u * 32 + lane;
            p[u] = j < end ? pts[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kTrip; ++u) {
            const size_t j0 = base + (size_t) u * 32;
            if (j0 >= end) break;
            const int cnt = (int) (end - j0 < 32 ? end - j0 : 32);   // members are lanes 0 .. cnt-1
            if (lane < cnt) {
                const double q0 = p[u].x, q1 = p[u].y, q2 = p[u].z;
                s[0] += q0; s[1] += q1; s[2] += q2;
                ss[0] += q0 * q0; ss[1] += q0 * q1; ss[2] += q0 * q2;
                ss[3] += q1 * q1; ss[4] += q1 * q2; ss[5] += q2 * q2;
            }
#pragma unroll
            for (int l = 0; l < 32; ++l) {
                const float vx = __shfl_sync(0xffffffffu, p[u].x, l), vy = __shfl_sync(0xffffffffu, p[u].y, l),
                            vz = __shfl_sync(0xffffffffu, p[u].z, l);
                if (l < cnt) {
                    cx = __fadd_rn(cx, vx);
                    cy = __fadd_rn(cy, vy);
                    cz = __fadd_rn(cz, vz);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int d = 0; d < 3; ++d) s[d] += __shfl_xor_sync(0xffffffffu, s[d], o);
#pragma unroll
        for (int d = 0; d < 6; ++d) ss[d] += __shfl_xor_sync(0xffffffffu, ss[d], o);
    }
    if (lane != 0) return;
    // raw sums; ndt_leaf_finish_kernel (one thread per voxel, all lanes busy) turns them into the statistics
    NdtLeafDev leaf;
    leaf.voxel = (int) k;
    leaf.n = (int) cnt_total;
    leaf.valid = 0;
    leaf.centroid[0] = cx; leaf.centroid[1] = cy; leaf.centroid[2] = cz;
    for (int d = 0; d < 3; ++d) leaf.mean[d] = s[d];
    for (int d = 0; d < 6; ++d) leaf.icov[d] = ss[d];
    leaves[slot] = leaf;
}

// mean, regularised covariance and its inverse of every voxel from the sums above (VoxelGridCovariance:
// >= 6 points, eigenvalue floor 0.01 * largest)
__global__ void __launch_bounds__(128) ndt_leaf_finish_kernel(NdtLeafDev *leaves, int n_voxels) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_voxels) return;
    NdtLeafDev leaf = leaves[slot];
    const int cnt = leaf.n;
    const double s[3] = {leaf.mean[0], leaf.mean[1], leaf.mean[2]};
    const double ss[6] = {leaf.icov[0], leaf.icov[1], leaf.icov[2], leaf.icov[3], leaf.icov[4], leaf.icov[5]};
    const float fn = (float) cnt;
    leaf.centroid[0] = __fdiv_rn(leaf.centroid[0], fn);
    leaf.centroid[1] = __fdiv_rn(leaf.centroid[1], fn);
    leaf.centroid[2] = __fdiv_rn(leaf.centroid[2], fn);
    for (int d = 0; d < 3; ++d) leaf.mean[d] = s[d] / cnt;
    for (int d = 0; d < 6; ++d) leaf.icov[d] = 0;
    if (cnt >= 6) {  // min_points_per_voxel_
        const double full[9] = {ss[0], ss[1], ss[2], ss[1], ss[3], ss[4], ss[2], ss[4], ss[5]};
        double cov[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                cov[3 * r + c] = (full[3 * r + c] - 2 * (s[r] * leaf.mean[c])) / cnt + leaf.mean[r] * leaf.mean[c];
        for (int q = 0; q < 9; ++q) cov[q] *= (cnt - 1.0) / cnt;
        double ev[3], V[9];
        eig_sym3(cov, ev, V);
        if (!(ev[0] < 0 || ev[1] < 0 || ev[2] <= 0)) {
            const double min_ev = 0.01 * ev[2];  // min_covar_eigvalue_mult_
            if (ev[0] < min_ev) {
                ev[0] = min_ev;
                if (ev[1] < min_ev) ev[1] = min_ev;
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) {
                        double acc = 0;
                        for (int q = 0; q < 3; ++q) acc += V[3 * r + q] * ev[q] * V[3 * c + q];
                        cov[3 * r + c] = acc;
                    }
            }
            const double a = cov[0], b = cov[1], c = cov[2], d = cov[4], e = cov[5], f = cov[8];
            const double det = a * (d * f - e * e) - b * (b * f - e * c) + c * (b * e - d * c);
            if (det != 0.0 && isfinite(det)) {
                const double id = 1.0 / det;
                leaf.icov[0] = (d * f - e * e) * id;
                leaf.icov[1] = (c * e - b * f) * id;
                leaf.icov[2] = (b * e - c * d) * id;
                leaf.icov[3] = (a * f - c * c) * id;
                leaf.icov[4] = (b * c - a * e) * id;
                leaf.icov[5] = (a * d - b * b) * id;
                bool ok = true;
                for (int q = 0; q < 6; ++q) ok = ok && isfinite(leaf.icov[q]);
                leaf.valid = ok ? 1 : 0;
            }
        }
    }
    leaves[slot] = leaf;
}

__device__ __forceinline__ unsigned hash_voxel(int v, unsigned mask) {
    unsigned h = (unsigned) v * 2654435761u;
    h ^= h >> 15;
    return h & mask;
}

__global__ void ndt_hash_insert_kernel(const NdtLeafDev *__restrict__ leaves, int n_leaves, int *table_key,
                                       int *table_slot, unsigned mask, int *n_valid, int *dense) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n_leaves && leaves[i].valid;
    const unsigned ballot = __ballot_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(n_valid, __popc(ballot));
    if (!valid) return;
    const int v = leaves[i].voxel;
    if (dense) {  // small enough grids: voxel index -> cell, one load per probe instead of a hash walk
        dense[v] = i;
        return;
    }
    unsigned h = hash_voxel(v, mask);
    for (;;) {
        const int prev = atomicCAS(&table_key[h], -1, v);
        if (prev == -1 || prev == v) {
            table_slot[h] = i;
            return;
        }
        h = (h + 1) & mask;
    }
}

struct NdtConsts {
    float T[12];       // fp32 pose matrix applied to the source (transformPointCloud with a Matrix4f)
    double ja[3], jb[3], jc[3], jd[3], je[3], jf[3], jg[3], jh[3];
    double a2[3], a3[3], b2[3], b3[3], c2[3], c3[3], d1[3], d2[3], d3[3], e1[3], e2[3], e3[3], f1[3], f2[3], f3[3];
    double gauss_d1, gauss_d2;
    float r2;          // (float)(res * res): FLANN keeps centroids with d2 < r2
    int with_hessian;
    GridDesc grid;
};

__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// four blocks per SM = the whole 592-block grid resident in one wave (128 registers, a few spills):
// measured 10 % faster than the 239-register build on the 1M-vs-5M configuration
#ifndef WCU_NDT_MINBLOCKS
#define WCU_NDT_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(kNdtThreads, WCU_NDT_MINBLOCKS) ndt_derivative_kernel(const float4 *__restrict__ src, int n_src,
                                                                     const NdtLeafDev *__restrict__ leaves,
                                                                     const int *__restrict__ table_key,
                                                                     const int *__restrict__ table_slot, unsigned mask,
                                                                     const int *__restrict__ dense,
                                                                     const __grid_constant__ NdtConsts kc,
                                                                     double *partial, unsigned *ticket,
                                                                     volatile double *host_sums, volatile int *host_seq,
                                                                     int seq) {
    // the per-pass constants travel as a kernel argument (no upload in front of every pass)
    __shared__ NdtConsts c;
    for (int w = threadIdx.x; w < (int) (sizeof(NdtConsts) / 4); w += blockDim.x)
        reinterpret_cast<int *>(&c)[w] = reinterpret_cast<const int *>(&kc)[w];
    __syncthreads();
    double acc[kNdtVals];
#pragma unroll
    for (int i = 0; i < kNdtVals; ++i) acc[i] = 0.0;

    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_src; idx += gridDim.x * blockDim.x) {
        const float4 p = src[idx];
        if (!finite3(p.x, p.y, p.z)) continue;
        // ((m0*x + m1*y) + m2*z) + m3, fp32, separately rounded
        const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.T[0], p.x), __fmul_rn(c.T[1], p.y)), __fmul_rn(c.T[2], p.z)), c.T[3]);
        const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.T[4], p.x), __fmul_rn(c.T[5], p.y)), __fmul_rn(c.T[6], p.z)), c.T[7]);
        const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(c.T[8], p.x), __fmul_rn(c.T[9], p.y)), __fmul_rn(c.T[10], p.z)), c.T[11]);
        const int i0 = (int) __fsub_rn(floorf(__fmul_rn(tx, c.grid.inv)), (float) c.grid.min_b[0]);
        const int i1 = (int) __fsub_rn(floorf(__fmul_rn(ty, c.grid.inv)), (float) c.grid.min_b[1]);
        const int i2 = (int) __fsub_rn(floorf(__fmul_rn(tz, c.grid.inv)), (float) c.grid.min_b[2]);
        const double x[3] = {p.x, p.y, p.z};
        // position inside the own voxel, in voxel units: a neighbour voxel at offset o can only hold a
        // centroid within `res` of the point if the point is closer than res to that voxel's box (the
        // centroid of a voxel's points lies inside the voxel) - on average half of the 27 are ruled out
        const float sx = __fmul_rn(tx, c.grid.inv), sy = __fmul_rn(ty, c.grid.inv), sz = __fmul_rn(tz, c.grid.inv);
        const float fx = sx - floorf(sx), fy = sy - floorf(sy), fz = sz - floorf(sz);
        const float gap[3][3] = {{fx, 0.0f, 1.0f - fx}, {fy, 0.0f, 1.0f - fy}, {fz, 0.0f, 1.0f - fz}};
        // Phase 1: the 27 hash probes, independent of each other so that their loads overlap (the
        // heavy per-hit arithmetic below would otherwise sit between them at 8 warps per SM).
        // A centroid within `res` of the point can only live in these 27 voxels; on lidar maps
        // 0-4 of them hold a cell.
        int hits[27];
        int n_hits = 0;
        if (dense) {
            // dense voxel -> cell table: every reachable neighbour is one independent load; the cells found
            // are then kept only if their centroid is within `res` (FLANN keeps d2 < r2), so that the
            // arithmetic loop below runs exactly once per contributing cell
            int slot_k[27];
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                const int v0 = i0 + (k % 3) - 1, v1 = i1 + ((k / 3) % 3) - 1, v2 = i2 + (k / 9) - 1;
                const float gx = gap[0][k % 3], gy = gap[1][(k / 3) % 3], gz = gap[2][k / 9];
                const bool reachable = gx * gx + gy * gy + gz * gz < 1.0002f;
                const bool inside = reachable && !(v0 < 0 || v1 < 0 || v2 < 0 || v0 >= c.grid.div_b[0] ||
                                                   v1 >= c.grid.div_b[1] || v2 >= c.grid.div_b[2]);
                slot_k[k] = inside ? __ldg(dense + (v0 * c.grid.mul[0] + v1 * c.grid.mul[1] + v2 * c.grid.mul[2])) : -1;
            }
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                if (slot_k[k] < 0) continue;
                const float *cen = leaves[slot_k[k]].centroid;
                const float dc = l2_simple(tx, ty, tz, __ldg(cen), __ldg(cen + 1), __ldg(cen + 2));
                if (dc < c.r2) hits[n_hits++] = slot_k[k];
            }
        } else {
            int first_key[27], vox_id[27];
            unsigned first_h[27];
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                const int v0 = i0 + (k % 3) - 1, v1 = i1 + ((k / 3) % 3) - 1, v2 = i2 + (k / 9) - 1;
                const float gx = gap[0][k % 3], gy = gap[1][(k / 3) % 3], gz = gap[2][k / 9];
                const bool reachable = gx * gx + gy * gy + gz * gz < 1.0002f;
                const bool inside = reachable && !(v0 < 0 || v1 < 0 || v2 < 0 || v0 >= c.grid.div_b[0] ||
                                                   v1 >= c.grid.div_b[1] || v2 >= c.grid.div_b[2]);
                vox_id[k] = inside ? v0 * c.grid.mul[0] + v1 * c.grid.mul[1] + v2 * c.grid.mul[2] : -2;
                first_h[k] = hash_voxel(vox_id[k], mask);
                first_key[k] = inside ? __ldg(table_key + first_h[k]) : -1;
            }
#pragma unroll
            for (int k = 0; k < 27; ++k) {
                if (vox_id[k] < 0) continue;
                unsigned h = first_h[k];
                int key = first_key[k], slot = -1;
                for (;;) {  // linear probing; almost always decided by the first key
                    if (key == vox_id[k]) {
                        slot = __ldg(table_slot + h);
                        break;
                    }
                    if (key == -1) break;
                    h = (h + 1) & mask;
                    key = __ldg(table_key + h);
                }
                if (slot < 0) continue;
                const float *cen = leaves[slot].centroid;
                const float dc = l2_simple(tx, ty, tz, __ldg(cen), __ldg(cen + 1), __ldg(cen + 2));
                if (dc < c.r2) hits[n_hits++] = slot;
            }
        }
        // Per-hit arithmetic, written out for the structure of the problem: the translation columns of the
        // point Jacobian are the identity (C J_q is a column of C, x'^T C J_q a component of C x'), only the
        // three rotation columns and the six rotation second derivatives are data; x'^T C J_q is taken as
        // (C x') . J_q (C is symmetric).  ~160 fused multiply-adds per hit instead of ~500 for the generic
        // 6x6 loops; the sums agree with the oracle's to rounding (1e-9 relative is the parity bar).
        bool have_point_terms = false;
        double J3[3], J4[3], J5[3], Hp[6][3];  // rotation columns of J; a b c d e f of eq. 6.21
        for (int hit = 0; hit < n_hits; ++hit) {
            const int slot = hits[hit];
            const NdtLeafDev &cell = leaves[slot];
            if (!have_point_terms) {  // computePointDerivatives, once per point
                have_point_terms = true;
                J3[0] = 0.0;            J3[1] = dot3(x, c.ja); J3[2] = dot3(x, c.jb);
                J4[0] = dot3(x, c.jc);  J4[1] = dot3(x, c.jd); J4[2] = dot3(x, c.je);
                J5[0] = dot3(x, c.jf);  J5[1] = dot3(x, c.jg); J5[2] = dot3(x, c.jh);
                if (c.with_hessian) {
                    Hp[0][0] = 0; Hp[0][1] = dot3(x, c.a2); Hp[0][2] = dot3(x, c.a3);
                    Hp[1][0] = 0; Hp[1][1] = dot3(x, c.b2); Hp[1][2] = dot3(x, c.b3);
                    Hp[2][0] = 0; Hp[2][1] = dot3(x, c.c2); Hp[2][2] = dot3(x, c.c3);
                    Hp[3][0] = dot3(x, c.d1); Hp[3][1] = dot3(x, c.d2); Hp[3][2] = dot3(x, c.d3);
                    Hp[4][0] = dot3(x, c.e1); Hp[4][1] = dot3(x, c.e2); Hp[4][2] = dot3(x, c.e3);
                    Hp[5][0] = dot3(x, c.f1); Hp[5][1] = dot3(x, c.f2); Hp[5][2] = dot3(x, c.f3);
                }
            }
            const double xt[3] = {(double) tx - cell.mean[0], (double) ty - cell.mean[1], (double) tz - cell.mean[2]};
            const double C00 = cell.icov[0], C01 = cell.icov[1], C02 = cell.icov[2], C11 = cell.icov[3],
                         C12 = cell.icov[4], C22 = cell.icov[5];
            const double Cx[3] = {C00 * xt[0] + C01 * xt[1] + C02 * xt[2], C01 * xt[0] + C11 * xt[1] + C12 * xt[2],
                                  C02 * xt[0] + C12 * xt[1] + C22 * xt[2]};
            double e = exp(-c.gauss_d2 * dot3(xt, Cx) / 2);
            const double score_inc = -c.gauss_d1 * e;
            e = c.gauss_d2 * e;
            if (e > 1 || e < 0 || e != e) continue;  // the score increment is dropped with it
            e *= c.gauss_d1;
            // x'^T C J_q for the six parameters
            const double s3 = dot3(Cx, J3), s4 = dot3(Cx, J4), s5 = dot3(Cx, J5);
            acc[0] += score_inc;
            acc[1] += Cx[0] * e;
            acc[2] += Cx[1] * e;
            acc[3] += Cx[2] * e;
            acc[4] += s3 * e;
            acc[5] += s4 * e;
            acc[6] += s5 * e;
            if (c.with_hessian) {
                const double md2 = -c.gauss_d2;
                // C J_q for the rotation columns
                const double c3[3] = {C00 * J3[0] + C01 * J3[1] + C02 * J3[2], C01 * J3[0] + C11 * J3[1] + C12 * J3[2],
                                      C02 * J3[0] + C12 * J3[1] + C22 * J3[2]};
                const double c4[3] = {C00 * J4[0] + C01 * J4[1] + C02 * J4[2], C01 * J4[0] + C11 * J4[1] + C12 * J4[2],
                                      C02 * J4[0] + C12 * J4[1] + C22 * J4[2]};
                const double c5[3] = {C00 * J5[0] + C01 * J5[1] + C02 * J5[2], C01 * J5[0] + C11 * J5[1] + C12 * J5[2],
                                      C02 * J5[0] + C12 * J5[1] + C22 * J5[2]};
                // rows in the order (a, b >= a) of the 21 unique entries, acc[7..27]
                // translation x translation: e (-d2 Cx_a Cx_b + C_ab)
                acc[7] += e * (md2 * Cx[0] * Cx[0] + C00);
                acc[8] += e * (md2 * Cx[0] * Cx[1] + C01);
                acc[9] += e * (md2 * Cx[0] * Cx[2] + C02);
                // translation a x rotation b: e (-d2 Cx_a s_b + (C J_b)_a)
                acc[10] += e * (md2 * Cx[0] * s3 + c3[0]);
                acc[11] += e * (md2 * Cx[0] * s4 + c4[0]);
                acc[12] += e * (md2 * Cx[0] * s5 + c5[0]);
                acc[13] += e * (md2 * Cx[1] * Cx[1] + C11);
                acc[14] += e * (md2 * Cx[1] * Cx[2] + C12);
                acc[15] += e * (md2 * Cx[1] * s3 + c3[1]);
                acc[16] += e * (md2 * Cx[1] * s4 + c4[1]);
                acc[17] += e * (md2 * Cx[1] * s5 + c5[1]);
                acc[18] += e * (md2 * Cx[2] * Cx[2] + C22);
                acc[19] += e * (md2 * Cx[2] * s3 + c3[2]);
                acc[20] += e * (md2 * Cx[2] * s4 + c4[2]);
                acc[21] += e * (md2 * Cx[2] * s5 + c5[2]);
                // rotation x rotation: e (-d2 s_a s_b + Cx . Hp_ab + J_b . C J_a); (3,3)=a (3,4)=b (3,5)=c (4,4)=d (4,5)=e (5,5)=f
                acc[22] += e * (md2 * s3 * s3 + dot3(Cx, Hp[0]) + dot3(J3, c3));
                acc[23] += e * (md2 * s3 * s4 + dot3(Cx, Hp[1]) + dot3(J4, c3));
                acc[24] += e * (md2 * s3 * s5 + dot3(Cx, Hp[2]) + dot3(J5, c3));
                acc[25] += e * (md2 * s4 * s4 + dot3(Cx, Hp[3]) + dot3(J4, c4));
                acc[26] += e * (md2 * s4 * s5 + dot3(Cx, Hp[4]) + dot3(J5, c4));
                acc[27] += e * (md2 * s5 * s5 + dot3(Cx, Hp[5]) + dot3(J5, c5));
            }
        }
    }
    // block reduction: warp shuffles, then one partial row per block
    __shared__ double s_red[kNdtThreads / 32][kNdtVals];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kNdtVals; ++i) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNdtVals) {
        double v = 0;
        for (int w = 0; w < kNdtThreads / 32; ++w) v += s_red[w][threadIdx.x];
        partial[(size_t) blockIdx.x * kNdtVals + threadIdx.x] = v;
    }
    // the last block to arrive adds the block rows in block order and hands the 28 sums to the host
    // through mapped memory: one launch, no copy and no stream synchronisation per derivative pass
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    // all 128 threads add the block rows (value = thread % 32, four interleaved chunks of blocks per value),
    // the chunk sums are then added in a fixed order: still independent of which block came last
    __shared__ double s_fin[4][32];
    {
        const int v = threadIdx.x & 31, chunk = threadIdx.x >> 5;
        double t = 0;
        if (v < kNdtVals)
            for (unsigned b = chunk; b < gridDim.x; b += 4) t += __ldcg(partial + (size_t) b * kNdtVals + v);
        s_fin[chunk][v] = t;
    }
    __syncthreads();
    if (threadIdx.x < kNdtVals)
        host_sums[threadIdx.x] = (s_fin[0][threadIdx.x] + s_fin[1][threadIdx.x]) + (s_fin[2][threadIdx.x] + s_fin[3][threadIdx.x]);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0u;
        *host_seq = seq;
    }
}

// ---- host-side numerics of the optimiser (fp64) ------------------------------------------------------
void svd_solve6(const double H[36], const double b[6], double x[6]) {
    // x = pinv(H) b by one-sided Jacobi; singular values <= eps * 6 * sigma_max are dropped
    // (Eigen::JacobiSVD::solve with its default threshold)
    double A[6][6], V[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            A[i][j] = H[6 * i + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < 6; ++k) {
                    alpha += A[k][p] * A[k][p];
                    beta += A[k][q] * A[k][q];
                    gamma += A[k][p] * A[k][q];
                }
                if (gamma == 0.0 || std::fabs(gamma) <= 1e-300) continue;
                if (std::fabs(gamma) <= 2.220446049250313e-16 * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 6; ++k) {
                    const double ap = A[k][p], aq = A[k][q];
                    A[k][p] = c * ap - s * aq;
                    A[k][q] = s * ap + c * aq;
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - s * vq;
                    V[k][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double sig[6], smax = 0;
    for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += A[k][j] * A[k][j];
        sig[j] = std::sqrt(s);
        smax = std::max(smax, sig[j]);
    }
    const double thr = 2.220446049250313e-16 * 6 * smax;
    for (int i = 0; i < 6; ++i) x[i] = 0;
    for (int j = 0; j < 6; ++j) {
        if (!(sig[j] > thr)) continue;
        double ub = 0;
        for (int k = 0; k < 6; ++k) ub += A[k][j] * b[k];
        const double coef = ub / (sig[j] * sig[j]);
        for (int i = 0; i < 6; ++i) x[i] += V[i][j] * coef;
    }
}

// (Translation * AngleAxis(rx, X) * AngleAxis(ry, Y) * AngleAxis(rz, Z)).matrix(), Scalar = float
void pose_to_matrix4f(const double p[6], float T[16]) {
    const float rx = static_cast<float>(p[3]), ry = static_cast<float>(p[4]), rz = static_cast<float>(p[5]);
    const float cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz),
                sz = std::sin(rz);
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
    T[0] = cy * cz;
    T[1] = -cy * sz;
    T[2] = sy;
    T[4] = sx * sy * cz + cx * sz;
    T[5] = -sx * sy * sz + cx * cz;
    T[6] = -sx * cy;
    T[8] = -cx * sy * cz + sx * sz;
    T[9] = cx * sy * sz + sx * cz;
    T[10] = cx * cy;
    T[3] = static_cast<float>(p[0]);
    T[7] = static_cast<float>(p[1]);
    T[11] = static_cast<float>(p[2]);
}

void set3(double *v, double a, double b, double c) {
    v[0] = a;
    v[1] = b;
    v[2] = c;
}

// computeAngleDerivatives (eq. 6.19 / 6.21) with PCL's |angle| < 10e-5 shortcut
void angle_terms(const double p[6], NdtConsts &t) {
    double cx, cy, cz, sx, sy, sz;
    if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
    if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
    if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
    set3(t.ja, (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy));
    set3(t.jb, (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy));
    set3(t.jc, (-sy * cz), sy * sz, cy);
    set3(t.jd, sx * cy * cz, (-sx * cy * sz), sx * sy);
    set3(t.je, (-cx * cy * cz), cx * cy * sz, (-cx * sy));
    set3(t.jf, (-cy * sz), (-cy * cz), 0);
    set3(t.jg, (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0);
    set3(t.jh, (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0);
    set3(t.a2, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy);
    set3(t.a3, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy));
    set3(t.b2, (cx * cy * cz), (-cx * cy * sz), (cx * sy));
    set3(t.b3, (sx * cy * cz), (-sx * cy * sz), (sx * sy));
    set3(t.c2, (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0);
    set3(t.c3, (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0);
    set3(t.d1, (-cy * cz), (cy * sz), (sy));
    set3(t.d2, (-sx * sy * cz), (sx * sy * sz), (sx * cy));
    set3(t.d3, (cx * sy * cz), (-cx * sy * sz), (-cx * cy));
    set3(t.e1, (sy * sz), (sy * cz), 0);
    set3(t.e2, (-sx * cy * sz), (-sx * cy * cz), 0);
    set3(t.e3, (cx * cy * sz), (cx * cy * cz), 0);
    set3(t.f1, (-cy * cz), (cy * sz), 0);
    set3(t.f2, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0);
    set3(t.f3, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0);
}

double psi_mt(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
double dpsi_mt(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

bool update_interval_mt(double &a_l, double &f_l, double &g_l, double &a_u, double &f_u, double &g_u, double a_t,
                        double f_t, double g_t) {
    if (f_t > f_l) {  // case U1 / a
        a_u = a_t; f_u = f_t; g_u = g_t;
        return false;
    }
    if (g_t * (a_l - a_t) > 0) {  // case U2 / b
        a_l = a_t; f_l = f_t; g_l = g_t;
        return false;
    }
    if (g_t * (a_l - a_t) < 0) {  // case U3 / c
        a_u = a_l; f_u = f_l; g_u = g_l;
        a_l = a_t; f_l = f_t; g_l = g_t;
        return false;
    }
    return true;
}

double cubic_min(double a_1, double f_1, double g_1, double a_2, double f_2, double g_2) {
    // minimiser of the cubic through (a_1, f_1, g_1), (a_2, f_2, g_2) [Sun & Yuan 2006, eq. 2.4.52 / 2.4.56]
    const double z = 3 * (f_2 - f_1) / (a_2 - a_1) - g_2 - g_1;
    const double w = std::sqrt(z * z - g_2 * g_1);
    return a_1 + (a_2 - a_1) * (w - g_1 - z) / (g_2 - g_1 + 2 * w);
}

double trial_value_mt(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t,
                      double g_t) {
    if (f_t > f_l) {  // case 1
        const double a_c = cubic_min(a_l, f_l, g_l, a_t, f_t, g_t);
        const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
        return (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) ? a_c : 0.5 * (a_q + a_c);
    }
    if (g_t * g_l < 0) {  // case 2
        const double a_c = cubic_min(a_l, f_l, g_l, a_t, f_t, g_t);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        return (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) ? a_c : a_s;
    }
    if (std::fabs(g_t) <= std::fabs(g_l)) {  // case 3
        const double a_c = cubic_min(a_l, f_l, g_l, a_t, f_t, g_t);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        const double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
        return (a_t > a_l) ? std::min(a_t + 0.66 * (a_u - a_t), a_t_next) : std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
    }
    return cubic_min(a_u, f_u, g_u, a_t, f_t, g_t);  // case 4
}

}  // namespace

struct NdtHandle {
    wavecu_ndt_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    float4 *d_src = nullptr, *d_tgt = nullptr;
    size_t n_src = 0, n_tgt = 0, src_cap = 0, tgt_cap = 0;
    // The derivative pass reads the source in Morton order: the 32 lanes of a warp then fall into the same
    // few voxels (their cell records become broadcast loads, their hit counts agree) instead of 32
    // different ones, as in a lidar driver's ring-major point order.  Sorted once per source cloud.
    MortonCloud src_sorted;
    bool src_dirty = true;
    bool grid_dirty = true;
    VoxelWork vox;
    NdtLeafDev *d_leaves = nullptr;
    size_t leaf_cap = 0;
    int n_leaves = 0;      // occupied voxels
    int n_cells = -1;      // of which normal-distribution cells (>= 6 points, usable covariance); -1: not read yet
    int *d_n_valid = nullptr;
    int *d_table_key = nullptr, *d_table_slot = nullptr;
    float4 *d_gathered = nullptr;    // target points in voxel-sorted order
    size_t gathered_cap = 0;
    int *d_starts = nullptr;         // voxel slot -> first element of its run
    size_t starts_cap = 0;
    int *d_dense = nullptr;          // voxel index -> cell for grids of up to kDenseMax voxels, else nullptr (hash)
    size_t dense_cap = 0;
    bool use_dense = false;
    static constexpr long long kDenseMax = 64ll << 20;   // 256 MB of ints
    size_t table_cap = 0;
    unsigned table_mask = 0;
    float resolution = 1.f;
    GridDesc grid{};
    bool grid_ok = false;
    double *d_partial = nullptr, *h_sums = nullptr;  // h_sums: mapped host memory
    unsigned *d_ticket = nullptr;
    int *h_seq = nullptr;  // mapped: number of the last pass whose sums are in h_sums
    int pass_seq = 0;
    int n_blocks = 0;
    long long launches = 0;
    long long derivative_passes = 0;
    float final_T[16];
    // optional per-kernel timing (CUDA events on the handle's stream around every derivative pass)
    bool profiling = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double derivative_ms = 0;
    int last_iterations = 0;

    int init() {
        WCU_CHECK(cudaSetDevice(device));
        if (!stream) {
            WCU_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            own_stream = true;
        }
        vox.device = device;
        vox.stream = stream;
        src_sorted.device = device;
        src_sorted.stream = stream;
        src_sorted.key_bits = 10;   // 30 sorted bits: four radix passes; finer keys buy no more coherence
        n_blocks = 148 * WCU_NDT_MINBLOCKS;   // the whole grid resident in one wave
        WCU_CHECK(cudaMalloc((void **) &d_partial, sizeof(double) * kNdtVals * (size_t) n_blocks));
        WCU_CHECK(cudaHostAlloc((void **) &h_sums, sizeof(double) * kNdtVals, cudaHostAllocMapped));
        WCU_CHECK(cudaHostAlloc((void **) &h_seq, sizeof(int), cudaHostAllocMapped));
        *h_seq = 0;
        WCU_CHECK(cudaMalloc((void **) &d_ticket, sizeof(unsigned)));
        WCU_CHECK(cudaMalloc((void **) &d_n_valid, sizeof(int)));
        WCU_CHECK(cudaMemsetAsync(d_ticket, 0, sizeof(unsigned), stream));
        return WAVECU_OK;
    }

    int upload(float4 *&dst, size_t &cap, size_t &n_dst, const float *xyzw, size_t n, bool from_device) {
        WCU_CHECK(cudaSetDevice(device));
        if (n > cap) {
            if (dst) WCU_CHECK(cudaFree(dst));
            dst = nullptr;
            WCU_CHECK(cudaMalloc((void **) &dst, (n + 64) * sizeof(float4)));
            cap = n + 64;
        }
        n_dst = n;
        if (n)
            WCU_CHECK(cudaMemcpyAsync(dst, xyzw, n * sizeof(float4),
                                      from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
        return WAVECU_OK;
    }

    float clamped_res() const { return prm.res < 0.05f ? 0.05f : prm.res; }  // src/ndt.cpp:23-26

    int build_grid() {
        resolution = clamped_res();
        grid_ok = false;
        n_leaves = 0;
        n_cells = 0;
        grid_dirty = false;
        if (n_tgt == 0) return WAVECU_OK;
        int status = 0;
        int rc = vox.prepare(d_tgt, n_tgt, resolution, &status);
        if (rc) return rc;
        if (status != 1) return WAVECU_OK;  // overflow: VoxelGridCovariance clears its output -> no cells
        grid = vox.grid;
        n_leaves = vox.n_voxels;
        if ((size_t) n_leaves > leaf_cap) {
            if (d_leaves) WCU_CHECK(cudaFree(d_leaves));
            d_leaves = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_leaves, ((size_t) n_leaves + 64) * sizeof(NdtLeafDev)));
            leaf_cap = (size_t) n_leaves + 64;
        }
        size_t slots = 64;
        while (slots < 2 * (size_t) n_leaves) slots <<= 1;
        if (slots > table_cap) {
            if (d_table_key) WCU_CHECK(cudaFree(d_table_key));
            if (d_table_slot) WCU_CHECK(cudaFree(d_table_slot));
            d_table_key = d_table_slot = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_table_key, slots * sizeof(int)));
            WCU_CHECK(cudaMalloc((void **) &d_table_slot, slots * sizeof(int)));
            table_cap = slots;
        }
        table_mask = (unsigned) (slots - 1);
        WCU_CHECK(cudaMemsetAsync(d_table_key, 0xff, slots * sizeof(int), stream));
        const long long n_vox = (long long) grid.div_b[0] * grid.div_b[1] * grid.div_b[2];
        use_dense = n_vox > 0 && n_vox <= kDenseMax;
        if (use_dense) {
            if ((size_t) n_vox > dense_cap) {
                if (d_dense) WCU_CHECK(cudaFree(d_dense));
                d_dense = nullptr;
                WCU_CHECK(cudaMalloc((void **) &d_dense, (size_t) n_vox * sizeof(int)));
                dense_cap = (size_t) n_vox;
            }
            WCU_CHECK(cudaMemsetAsync(d_dense, 0xff, (size_t) n_vox * sizeof(int), stream));
        }
        if ((size_t) n_leaves > starts_cap) {
            if (d_starts) WCU_CHECK(cudaFree(d_starts));
            d_starts = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_starts, ((size_t) n_leaves + 64) * sizeof(int)));
            starts_cap = (size_t) n_leaves + 64;
        }
        ndt_starts_kernel<<<(unsigned) ((n_tgt + 255) / 256), 256, 0, stream>>>(vox.d_keys, vox.d_pos, n_tgt, d_starts);
        if (n_tgt > gathered_cap) {
            if (d_gathered) WCU_CHECK(cudaFree(d_gathered));
            d_gathered = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_gathered, (n_tgt + 64) * sizeof(float4)));
            gathered_cap = n_tgt + 64;
        }
        ndt_gather_kernel<<<(unsigned) ((n_tgt + 255) / 256), 256, 0, stream>>>(d_tgt, vox.d_keys, vox.d_vals, n_tgt,
                                                                                d_gathered);
        ++launches;
        if (n_leaves)
            ndt_leaf_kernel<<<(unsigned) (((size_t) n_leaves * 32 + 255) / 256), 256, 0, stream>>>(
                d_gathered, vox.d_keys, d_starts, n_leaves, n_tgt, d_leaves);
        if (n_leaves) ndt_leaf_finish_kernel<<<(unsigned) ((n_leaves + 127) / 128), 128, 0, stream>>>(d_leaves, n_leaves);
        launches += 2;
        WCU_CHECK(cudaMemsetAsync(d_n_valid, 0, sizeof(int), stream));
        n_cells = -1;
        if (n_leaves)
            ndt_hash_insert_kernel<<<(n_leaves + 255) / 256, 256, 0, stream>>>(d_leaves, n_leaves, d_table_key,
                                                                               d_table_slot, table_mask, d_n_valid,
                                                                               use_dense ? d_dense : nullptr);
        launches += 2;
        WCU_CHECK(cudaGetLastError());
        grid_ok = true;
        return WAVECU_OK;
    }

    int sort_source() {
        if (!src_dirty) return WAVECU_OK;
        src_dirty = false;
        if (n_src == 0) return WAVECU_OK;
        int rc = src_sorted.upload((const float *) d_src, n_src, true);
        if (rc) return rc;
        rc = src_sorted.sort(n_src);
        launches += src_sorted.launches;
        src_sorted.launches = 0;
        return rc;
    }

    // computeDerivatives at pose p (the source is transformed by the fp32 matrix of p)
    int derivatives(const double p[6], const float T[16], double gd1, double gd2, bool with_hessian, double *score,
                    double g[6], double H[36]) {
        NdtConsts c;
        std::memcpy(c.T, T, sizeof(float) * 12);
        angle_terms(p, c);
        c.gauss_d1 = gd1;
        c.gauss_d2 = gd2;
        c.r2 = static_cast<float>((double) resolution * (double) resolution);
        c.with_hessian = with_hessian ? 1 : 0;
        c.grid = grid;
        const int seq = ++pass_seq;
        if (profiling) {
            if (!ev0) {
                WCU_CHECK(cudaEventCreate(&ev0));
                WCU_CHECK(cudaEventCreate(&ev1));
            }
            WCU_CHECK(cudaEventRecord(ev0, stream));
        }
        ndt_derivative_kernel<<<n_blocks, kNdtThreads, 0, stream>>>(src_sorted.d_sorted, (int) n_src, d_leaves, d_table_key,
                                                                     d_table_slot, table_mask, use_dense ? d_dense : nullptr, c,
                                                                     d_partial, d_ticket,
                                                                     h_sums, h_seq, seq);
        if (profiling) WCU_CHECK(cudaEventRecord(ev1, stream));
        ++launches;
        ++derivative_passes;
        WCU_CHECK(cudaGetLastError());
        for (int spins = 0; *(volatile int *) h_seq != seq;) {  // the kernel's own hand-over
            if (++spins % 4096 == 0) {
                const cudaError_t q = cudaStreamQuery(stream);
                if (q != cudaErrorNotReady) {
                    if (q != cudaSuccess) WCU_CHECK(q);
                    if (*(volatile int *) h_seq != seq) {
                        set_last_error("ndt_derivative_kernel finished without publishing its sums");
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
            derivative_ms += ms;
        }
        *score = h_sums[0];
        for (int i = 0; i < 6; ++i) g[i] = h_sums[1 + i];
        if (with_hessian) {
            int u = 7;
            for (int a = 0; a < 6; ++a)
                for (int b = a; b < 6; ++b) {
                    H[6 * a + b] = H[6 * b + a] = h_sums[u];
                    ++u;
                }
        }
        return WAVECU_OK;
    }

    int match(double *T_out, int *converged_out, int *iterations_out) {
        WCU_CHECK(cudaSetDevice(device));
        for (int i = 0; i < 16; ++i) final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
        launches = 0;
        derivative_passes = 0;
        derivative_ms = 0;
        bool converged = false;
        int nr_iterations = 0;
        static const bool trace_steps = getenv("WAVECU_NDT_TRACE") != nullptr;   // debugging aid
        if (grid_dirty || clamped_res() != resolution) {
            const int rc = build_grid();
            if (rc) return rc;
        }
        // a target without a single cell (no finite point, or VoxelGridCovariance's index overflow, which clears
        // its output): every derivative is zero, so the first Newton step has norm 0 and PCL reports convergence
        if (n_src && n_tgt && !grid_ok) converged = true;
        if (n_src && n_tgt && grid_ok) {
            {
                const int rc = sort_source();
                if (rc) return rc;
            }
            const double outlier_ratio = 0.55;
            const double c1 = 10 * (1 - outlier_ratio);
            const double c2 = outlier_ratio / std::pow((double) resolution, 3);
            const double d3 = -std::log(c2);
            const double gd1 = -std::log(c1 + c2) - d3;
            const double gd2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / gd1);
            const double step_max = (double) prm.step_size, step_min = prm.t_eps / 2;
            double p[6] = {0, 0, 0, 0, 0, 0}, g[6], H[36], score = 0;
            int rc = derivatives(p, final_T, gd1, gd2, true, &score, g, H);
            if (rc) return rc;
            while (!converged) {
                double neg_g[6], dir[6];
                for (int i = 0; i < 6; ++i) neg_g[i] = -g[i];
                svd_solve6(H, neg_g, dir);
                double nrm = 0;
                for (int i = 0; i < 6; ++i) nrm += dir[i] * dir[i];
                double delta_p_norm = std::sqrt(nrm);
                if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
                    converged = delta_p_norm == delta_p_norm;
                    break;
                }
                for (int i = 0; i < 6; ++i) dir[i] /= delta_p_norm;
                // ---- computeStepLengthMT ----
                double a_t = 0;
                const double phi_0 = -score;
                double d_phi_0 = 0;
                for (int i = 0; i < 6; ++i) d_phi_0 -= g[i] * dir[i];
                bool have_step = true;
                if (d_phi_0 >= 0) {
                    if (d_phi_0 == 0) have_step = false;  // return 0
                    else {
                        d_phi_0 *= -1;
                        for (int i = 0; i < 6; ++i) dir[i] *= -1;
                    }
                }
                if (have_step) {
                    const double mu = 1.e-4, nu = 0.9;
                    double a_l = 0, a_u = 0;
                    double f_l = psi_mt(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsi_mt(d_phi_0, d_phi_0, mu);
                    double f_u = psi_mt(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsi_mt(d_phi_0, d_phi_0, mu);
                    // PCL 1.8: `> 0`, true whenever step_max > step_min, which skips the loop below;
                    // PCL >= 1.9: `< 0`, the search runs (wavecu.h, wavecu_ndt_params::line_search)
                    bool interval_converged = prm.line_search == WAVECU_NDT_LS_PCL18 ? (step_max - step_min) > 0
                                                                                      : (step_max - step_min) < 0,
                         open_interval = true;
                    a_t = std::max(std::min(delta_p_norm, step_max), step_min);
                    double x_t[6];
                    auto evaluate = [&](bool hess) -> int {
                        for (int i = 0; i < 6; ++i) x_t[i] = p[i] + dir[i] * a_t;
                        pose_to_matrix4f(x_t, final_T);
                        return derivatives(x_t, final_T, gd1, gd2, hess, &score, g, H);
                    };
                    rc = evaluate(true);
                    if (rc) return rc;
                    double phi_t = -score, d_phi_t = 0;
                    for (int i = 0; i < 6; ++i) d_phi_t -= g[i] * dir[i];
                    double psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
                    int step_iterations = 0;
                    while (!interval_converged && step_iterations < 10 && !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
                        a_t = open_interval ? trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t)
                                            : trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
                        a_t = std::max(std::min(a_t, step_max), step_min);
                        rc = evaluate(false);
                        if (rc) return rc;
                        phi_t = -score;
                        d_phi_t = 0;
                        for (int i = 0; i < 6; ++i) d_phi_t -= g[i] * dir[i];
                        psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu);
                        d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
                        if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
                            open_interval = false;
                            f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
                            g_l = g_l + mu * d_phi_0;
                            f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
                            g_u = g_u + mu * d_phi_0;
                        }
                        interval_converged = open_interval
                                               ? update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t)
                                               : update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
                        ++step_iterations;
                    }
                    if (step_iterations) {  // computeHessian at the accepted point
                        double g_keep[6], s_keep = score;
                        std::memcpy(g_keep, g, sizeof g);
                        rc = derivatives(x_t, final_T, gd1, gd2, true, &score, g, H);
                        if (rc) return rc;
                        std::memcpy(g, g_keep, sizeof g);
                        score = s_keep;
                    }
                }
                delta_p_norm = a_t;
                if (trace_steps) fprintf(stderr, "[wavecu ndt] iteration %d step %.17g score %.17g\n", nr_iterations, a_t, score);
                for (int i = 0; i < 6; ++i) p[i] = p[i] + dir[i] * delta_p_norm;
                if (nr_iterations > prm.max_iter || (nr_iterations && (std::fabs(delta_p_norm) < prm.t_eps)))
                    converged = true;
                nr_iterations++;
            }
        }
        if (T_out)
            for (int i = 0; i < 16; ++i) T_out[i] = (double) final_T[i];
        if (converged_out) *converged_out = converged ? 1 : 0;
        if (iterations_out) *iterations_out = nr_iterations;
        return WAVECU_OK;
    }

    void release() {
        cudaSetDevice(device);
        vox.release();
        src_sorted.release();
        if (d_dense) cudaFree(d_dense);
        if (d_starts) cudaFree(d_starts);
        if (d_gathered) cudaFree(d_gathered);
        for (void *p : {(void *) d_src, (void *) d_tgt, (void *) d_leaves, (void *) d_table_key, (void *) d_table_slot,
                        (void *) d_partial})
            if (p) cudaFree(p);
        if (h_sums) cudaFreeHost(h_sums);
        if (h_seq) cudaFreeHost(h_seq);
        if (d_ticket) cudaFree(d_ticket);
        if (d_n_valid) cudaFree(d_n_valid);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

}  // namespace wavecu

using namespace wavecu;

struct wavecu_ndt {
    NdtHandle h;
};

extern "C" {

void wavecu_ndt_default_params(wavecu_ndt_params *p) {
    if (!p) return;
    p->step_size = 3;
    p->max_iter = 100;
    p->t_eps = 1e-8;
    p->res = 5.f;
    p->line_search = WAVECU_NDT_LS_MORE_THUENTE;
}

int wavecu_ndt_create(const wavecu_ndt_params *params, int device, void *stream, wavecu_ndt **out) {
    if (!out) return WAVECU_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_last_error("no such CUDA device (libwavecu has no CPU fallback)");
        return WAVECU_ERR_CUDA;
    }
    wavecu_ndt *w = new wavecu_ndt();
    if (params) w->h.prm = *params;
    else wavecu_ndt_default_params(&w->h.prm);
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

int wavecu_ndt_destroy(wavecu_ndt *w) {
    if (!w) return WAVECU_OK;
    w->h.release();
    delete w;
    return WAVECU_OK;
}

int wavecu_ndt_set_params(wavecu_ndt *w, const wavecu_ndt_params *params) {
    if (!w || !params) return WAVECU_ERR_ARG;
    w->h.prm = *params;
    return WAVECU_OK;
}

int wavecu_ndt_set_source(wavecu_ndt *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    w->h.src_dirty = true;
    return w->h.upload(w->h.d_src, w->h.src_cap, w->h.n_src, xyzw, n, false);
}
int wavecu_ndt_set_source_device(wavecu_ndt *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    w->h.src_dirty = true;
    return w->h.upload(w->h.d_src, w->h.src_cap, w->h.n_src, (const float *) d, n, true);
}
int wavecu_ndt_set_target(wavecu_ndt *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    w->h.grid_dirty = true;
    return w->h.upload(w->h.d_tgt, w->h.tgt_cap, w->h.n_tgt, xyzw, n, false);
}
int wavecu_ndt_set_target_device(wavecu_ndt *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    w->h.grid_dirty = true;
    return w->h.upload(w->h.d_tgt, w->h.tgt_cap, w->h.n_tgt, (const float *) d, n, true);
}

int wavecu_ndt_match(wavecu_ndt *w, double T_out[16], int *converged, int *iterations) {
    if (!w) return WAVECU_ERR_ARG;
    return w->h.match(T_out, converged, iterations);
}

int wavecu_ndt_grid(wavecu_ndt *w, int *n_cells, int *voxel, int *count, float *centroid3, double *mean3, double *icov9,
                    int capacity) {
    if (!w || !n_cells) return WAVECU_ERR_ARG;
    NdtHandle &h = w->h;
    WCU_CHECK(cudaSetDevice(h.device));
    if (h.grid_dirty || h.clamped_res() != h.resolution) {
        const int rc = h.build_grid();
        if (rc) return rc;
    }
    std::vector<NdtLeafDev> all((size_t) h.n_leaves);
    if (h.n_leaves) {
        WCU_CHECK(cudaMemcpyAsync(all.data(), h.d_leaves, sizeof(NdtLeafDev) * (size_t) h.n_leaves, cudaMemcpyDeviceToHost,
                                  h.stream));
        WCU_CHECK(cudaStreamSynchronize(h.stream));
    }
    int m = 0;
    for (const NdtLeafDev &l : all) {
        if (!l.valid) continue;
        if (m < capacity) {
            if (voxel) voxel[m] = l.voxel;
            if (count) count[m] = l.n;
            for (int d = 0; d < 3; ++d) {
                if (centroid3) centroid3[3 * m + d] = l.centroid[d];
                if (mean3) mean3[3 * m + d] = l.mean[d];
            }
            if (icov9) {
                const double *c = l.icov;
                const double full[9] = {c[0], c[1], c[2], c[1], c[3], c[4], c[2], c[4], c[5]};
                std::memcpy(icov9 + 9 * m, full, sizeof full);
            }
        }
        ++m;
    }
    *n_cells = m;
    return WAVECU_OK;
}

int wavecu_ndt_derivatives(wavecu_ndt *w, const double pose6[6], const float T16[16], double *score, double g6[6],
                           double H36[36]) {
    if (!w || !pose6 || !T16 || !score || !g6 || !H36) return WAVECU_ERR_ARG;
    NdtHandle &h = w->h;
    WCU_CHECK(cudaSetDevice(h.device));
    if (h.grid_dirty || h.clamped_res() != h.resolution) {
        const int rc = h.build_grid();
        if (rc) return rc;
    }
    *score = 0;
    for (int i = 0; i < 6; ++i) g6[i] = 0;
    for (int i = 0; i < 36; ++i) H36[i] = 0;
    if (!h.grid_ok || !h.n_src) return WAVECU_OK;
    {
        const int rc = h.sort_source();
        if (rc) return rc;
    }
    const double o = 0.55, c1 = 10 * (1 - o), c2 = o / std::pow((double) h.resolution, 3), d3 = -std::log(c2);
    const double gd1 = -std::log(c1 + c2) - d3;
    const double gd2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / gd1);
    return h.derivatives(pose6, T16, gd1, gd2, true, score, g6, H36);
}

int wavecu_ndt_stats(wavecu_ndt *w, long long *kernel_launches, long long *derivative_passes, int *n_cells) {
    if (!w) return WAVECU_ERR_ARG;
    if (kernel_launches) *kernel_launches = w->h.launches + w->h.vox.launches;
    if (derivative_passes) *derivative_passes = w->h.derivative_passes;
    if (n_cells) {
        NdtHandle &h = w->h;
        if (h.n_cells < 0) {  // counted by the hash-insert kernel of the last grid build
            WCU_CHECK(cudaSetDevice(h.device));
            WCU_CHECK(cudaMemcpyAsync(&h.n_cells, h.d_n_valid, sizeof(int), cudaMemcpyDeviceToHost, h.stream));
            WCU_CHECK(cudaStreamSynchronize(h.stream));
        }
        *n_cells = h.n_cells;
    }
    return WAVECU_OK;
}

int wavecu_ndt_set_profiling(wavecu_ndt *w, int enabled) {
    if (!w) return WAVECU_ERR_ARG;
    w->h.profiling = enabled != 0;
    return WAVECU_OK;
}

int wavecu_ndt_timing(wavecu_ndt *w, double *derivative_kernel_ms) {
    if (!w) return WAVECU_ERR_ARG;
    if (derivative_kernel_ms) *derivative_kernel_ms = w->h.derivative_ms;
    return WAVECU_OK;
}

}  // extern "C"
