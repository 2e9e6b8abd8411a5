// Morton-ordered clouds and the implicit bounding-volume tree over the target.
//
// Replaces pcl::KdTreeFLANN (FLANN KDTreeSingleIndex) that pcl::Registration::initCompute builds
// for every align() target (reference call sites: wave_matching/src/icp.cpp:124-126,
// src/icp_pcl_functions.cpp:67-68).  B200 design instead of a pointer kd-tree:
//   * points are sorted once by a 63-bit Morton key and stored as float4 (w carries the original
//     index), so a leaf of 8 consecutive points is exactly one aligned 128-byte line;
//   * the tree over the leaves is a complete binary heap of AABBs (children of i are 2i, 2i+1,
//     a child pair is one aligned 64-byte segment) - no pointers, built bottom-up in one launch;
//   * queries are Morton-sorted as well, so the 32 lanes of a warp walk the same few nodes and
//     leaves and their loads coalesce in L1/L2 (the whole structure for 1 M points is ~24 MB and
//     stays resident in the 126 MB L2).
#pragma once
#include "common.cuh"

namespace wavecu {

struct MortonCloud {
    int device = 0;
    cudaStream_t stream = nullptr;
    size_t n = 0;            // points in d_raw
    size_t cap = 0;          // allocated points
    float4 *d_raw = nullptr;     // as given (original order)
    float4 *d_sorted = nullptr;  // Morton order, w = original index bits
    unsigned *d_bbox = nullptr;  // 6 order-preserving uints: lo xyz, hi xyz
    unsigned long long *d_keys = nullptr, *d_keys_alt = nullptr;
    unsigned *d_vals = nullptr, *d_vals_alt = nullptr;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    size_t sorted_cap = 0;
    long long launches = 0;

    int reserve(size_t n_points, size_t sorted_points);
    int upload(const float *xyzw, size_t n_points, bool from_device);
    // bbox + Morton keys + radix sort + gather into d_sorted[0..n_sorted_pad) (pads = +inf)
    int sort(size_t n_sorted_pad, const float4 *d_extra_in = nullptr, float4 *d_extra_out = nullptr);
    void release();
};

struct TargetIndex {
    MortonCloud cloud;
    int P = 0;               // leaf slots, power of two; nodes are 1..2P-1, leaf j is node P+j
    Node *d_nodes = nullptr;
    int *d_flags = nullptr;
    float4 *d_nrm_raw = nullptr, *d_nrm_sorted = nullptr;  // optional normals
    size_t nrm_n = 0;
    size_t node_cap = 0, nrm_cap = 0;
    bool dirty = true;

    int set_points(const float *xyzw, size_t n, bool from_device);
    int set_normals(const float *nxyzw, size_t n, bool from_device);
    int build();             // sort + tree; clears dirty
    void release();
};

__device__ __forceinline__ unsigned float_to_ordered(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Exact 1-NN of (qx,qy,qz) under l2_simple with the lowest original index among exact ties.
// best / best_idx come in as the current bound (e.g. the max-correspondence threshold with
// best_idx = INT_MAX, or a warm start) and leave as the result.  A subtree is skipped only when
// its bound is strictly greater than best, so equal-distance candidates are always examined.
__device__ __forceinline__ void nn_search(float qx, float qy, float qz, const Node *__restrict__ nodes,
                                          const float4 *__restrict__ pts, int P, float &best, int &best_idx,
                                          int &best_pos) {
    unsigned node = 1u, pending = 0u;
    {
        const Node root = nodes[1];
        if (aabb_dist(qx, qy, qz, root.lo, root.hi) > best) return;
    }
    for (;;) {
        bool up = false;
        if (node >= (unsigned) P) {
            const float4 *leaf = pts + (size_t)(node - P) * kLeaf;
#pragma unroll
            for (int k = 0; k < kLeaf; ++k) {
                const float4 p = __ldg(leaf + k);
                const float d = l2_simple(qx, qy, qz, p.x, p.y, p.z);
                const int idx = __float_as_int(p.w);
                if (d < best || (d == best && idx < best_idx)) {
                    best = d;
                    best_idx = idx;
                    best_pos = (int) (node - P) * kLeaf + k;
                }
            }
            up = true;
        } else {
            const Node c0 = nodes[2 * node], c1 = nodes[2 * node + 1];
            const float d0 = aabb_dist(qx, qy, qz, c0.lo, c0.hi);
            const float d1 = aabb_dist(qx, qy, qz, c1.lo, c1.hi);
            const bool right_first = d1 < d0;
            const float dn = right_first ? d1 : d0, df = right_first ? d0 : d1;
            if (dn > best) {
                up = true;
            } else {
                pending = (pending << 1) | (df <= best ? 1u : 0u);
                node = 2 * node + (right_first ? 1u : 0u);
            }
        }
        if (up) {
            for (;;) {
                if (node == 1u) return;
                if (pending & 1u) {
                    pending &= ~1u;
                    node ^= 1u;
                    const Node s = nodes[node];
                    if (aabb_dist(qx, qy, qz, s.lo, s.hi) <= best) break;
                } else {
                    node >>= 1;
                    pending >>= 1;
                }
            }
        }
    }
}

}  // namespace wavecu
