// Morton-ordered clouds and the implicit bounding-volume tree over the target.
//
// Replaces pcl::KdTreeFLANN (FLANN KDTreeSingleIndex) that pcl::Registration::initCompute builds
// for every align() target (reference call sites: wave_matching/src/icp.cpp:124-126,
// src/icp_pcl_functions.cpp:67-68).  B200 design instead of a pointer kd-tree:
//   * points are sorted once by a 63-bit Morton key and stored as float4 (w carries the original
//     index); a leaf is a run of <= 8 consecutive sorted points;
//   * the hierarchy is the radix tree of the sorted keys (split at the highest differing bit),
//     built bottom-up in ONE launch (one thread per point, rendezvous by atomicExch - the
//     agglomerative LBVH construction), emitting 64-byte nodes that hold both children's boxes;
//     measured on the 1 M-point lidar workload this visits ~3 leaves / ~30 nodes per query where
//     a count-balanced split of the same Morton order visits 22 / 220 (tools/probe_*.cpp);
//   * queries are Morton-sorted as well, so the 32 lanes of a warp walk the same few nodes and
//     leaves and their loads coalesce in L1/L2 (the whole structure for 1 M points is ~24 MB and
//     stays resident in the 126 MB L2).
#pragma once
#include "common.cuh"

namespace wavecu {

// A captured launch sequence, replayed while its key (sizes and buffer addresses) stays the same:
// the build of one cloud is ~25 short kernels, and enqueueing them one by one costs the host more
// time (~0.1 ms) than a small cloud's build takes on the device.
struct GraphCache {
    cudaGraphExec_t exec = nullptr;
    unsigned long long key[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long launches = 0;
    // A sequence is captured only when the same key comes twice in a row: clouds whose size changes on
    // every call (real scans, voxel-filter outputs, MultiMatcher jobs) would otherwise pay a capture and
    // an instantiation - more than the launches they save - each time, and are enqueued directly instead.
    unsigned long long pending[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool has_pending = false;
    bool seen_before(const unsigned long long (&k)[8]) {
        bool same = has_pending;
        for (int i = 0; i < 8; ++i) {
            same = same && pending[i] == k[i];
            pending[i] = k[i];
        }
        has_pending = true;
        return same;
    }
    bool matches(const unsigned long long (&k)[8]) const {
        if (!exec) return false;
        for (int i = 0; i < 8; ++i)
            if (key[i] != k[i]) return false;
        return true;
    }
    void release() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr;
    }
};
bool graphs_enabled();  // false with WAVECU_NO_GRAPH=1 (debugging / profiling aid)

struct MortonCloud {
    int device = 0;
    cudaStream_t stream = nullptr;        // sort / gather kernels
    // Optional second stream for upload(): the copy then overlaps kernels queued on `stream` (another
    // cloud's sort, the tree build).  ev_up orders sort() behind the copy, ev_used orders the next
    // copy behind the last sort that read d_raw.  nullptr: everything runs on `stream`.
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_up = nullptr, ev_used = nullptr;
    cudaEvent_t ev_tl[3] = {nullptr, nullptr, nullptr};   // WAVECU_TIMELINE (debugging aid): keys / sorted / gathered
    bool up_pending = false, used_pending = false;
    size_t n = 0;            // points in d_raw
    size_t cap = 0;          // allocated points
    float4 *d_raw = nullptr;     // as given (original order)
    float4 *d_sorted = nullptr;  // Morton order, w = original index bits
    unsigned *d_bbox = nullptr;  // 6 order-preserving uints: lo xyz, hi xyz
    unsigned long long *d_keys = nullptr, *d_keys_alt = nullptr;   // unsorted keys / radix-sort double buffer
    unsigned *d_vals = nullptr, *d_vals_alt = nullptr;
    unsigned long long *d_keys_sorted = nullptr;   // whichever of the two buffers the sort left its result in
    unsigned *d_vals_sorted = nullptr;
    GraphCache sort_graph;
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    size_t sorted_cap = 0;
    long long launches = 0;
    int key_bits = 15;           // Morton bits per axis; sorted bits = 3 * key_bits + 1 (validity bit)

    int reserve(size_t n_points, size_t sorted_points);
    int upload(const float *xyzw, size_t n_points, bool from_device);
    // bbox + Morton keys + radix sort + gather into d_sorted[0..n_sorted_pad) (pads = +inf)
    int sort(size_t n_sorted_pad, const float4 *d_extra_in = nullptr, float4 *d_extra_out = nullptr);
    // the bare launch sequence of sort() on `stream` (buffers reserved by the caller; capturable)
    int enqueue_sort(size_t n_sorted_pad, const float4 *d_extra_in, float4 *d_extra_out);
    int pre_sort(size_t n_sorted_pad);   // reserve + order behind the pending upload
    int post_sort();                     // mark d_raw as consumed
    void release();
};

// Traversal node of the radix-tree LBVH: both children's boxes and links in one aligned 64-byte
// record, so one node visit is one segment load.  link >= 0: index of the child's own node;
// link < 0: the child is a leaf, a run of <= kLeaf consecutive sorted points (make_leaf_link).
// A single-point child is the degenerate box lo = hi = the point, with the point's original index
// in hi.w: aabb_dist of such a box is bit-for-bit l2_simple of the point.
// Halves are addressed as 2*node + side.
struct __align__(64) TNode {
    float4 lo0;  // xyz, w = link0 (int bits)
    float4 hi0;  // xyz
    float4 lo1;  // xyz, w = link1
    float4 hi1;  // xyz
};

// Quantisation of the target frame behind the Morton keys (monotone in every coordinate).
struct QuantParams {
    float lx, ly, lz, scale, qmax;
    int bits;
};

// Entry table of the tree: for every cell of the coarsest 2^kCellBits^3 Morton grid, the link of
// the subtree (or single point) that holds exactly the target points of that cell; kCellEmpty
// (memset 0x80) for a cell without points.  A search whose ball lies inside a few cells starts at
// their entries instead of walking the ~10 levels above them, most of which only split empty space
// around the dense part of a lidar scan.
#ifndef WCU_CELL_BITS
#define WCU_CELL_BITS 7
#endif
constexpr int kCellBits = WCU_CELL_BITS;  // 0: no table
constexpr size_t kCellCount = kCellBits > 0 ? (size_t) 1 << (3 * kCellBits) : 1;
typedef int CellEntry;
constexpr int kCellEmpty = (int) 0x80808080;  // < every leaf link (those are >= -2^27)

struct TreeRoot {
    float4 lo, hi;   // box of all finite points; lo.w = link (int bits), hi.w = count (int bits)
};

// Box pyramid over the Morton-sorted target for the tiled correspondence kernel (tile_nn.cuh): level 0
// holds the bounding box of every run of 8 consecutive sorted points, level k the box of 8 level k-1
// boxes (level 1 = 64 points: the unit that is staged into shared memory).  Box i of level k is the pair
// (lv[k][2 i], lv[k][2 i + 1]) = (lo, hi), xyz used.  Runs are contiguous in the sorted array, so a
// level-1 box names one 1 KB slice of points and one 256 B slice of level-0 boxes - bulk-copy units.
// Pads (+inf points behind the last finite one) give +inf box corners, which no query reaches.
constexpr int kBoxLevelsMax = 10;
constexpr int kBoxTop = 128;      // the top level has at most this many boxes
struct BoxLevels {
    const float4 *lv[kBoxLevelsMax];
    int cnt[kBoxLevelsMax];
    int n_levels;                 // >= 2
};

struct TargetIndex {
    MortonCloud cloud;
    TNode *d_nodes = nullptr;    // n-1 records, indexed by split position
    int *d_other = nullptr;      // n-1 rendezvous slots of the bottom-up build
    TreeRoot *d_root = nullptr;
    CellEntry *d_cells = nullptr;   // kCellCount entries, rebuilt with the tree
    bool want_boxes = false;        // also build the box pyramid (ICP: the tiled correspondence kernel)
    float4 *d_boxes = nullptr;      // all levels, level 0 first
    size_t boxes_cap = 0;           // in float4
    BoxLevels boxes{};              // device pointers into d_boxes; valid when want_boxes and !dirty
    size_t sorted_pad() const { return want_boxes ? ((cloud.n + 63) / 64) * 64 : cloud.n; }
    float4 *d_nrm_raw = nullptr, *d_nrm_sorted = nullptr;  // optional normals
    size_t nrm_n = 0;
    size_t node_cap = 0, nrm_cap = 0, nrm_sorted_cap = 0;
    bool dirty = true;

    int set_points(const float *xyzw, size_t n, bool from_device);
    int set_normals(const float *nxyzw, size_t n, bool from_device);
    int build();             // sort + tree; clears dirty (normals are gathered by sort_normals())
    int enqueue_build();
    GraphCache build_graph;
    int sort_normals();      // d_nrm_sorted <- d_nrm_raw in the Morton order of the last build
    int reserve_sorted_normals();
    cudaEvent_t ev_nrm_up = nullptr;
    bool nrm_up_pending = false, nrm_dirty = false;
    // fills d_nrm_sorted with unit normals estimated from the k nearest neighbours of every target
    // point (principal direction of least variance, oriented towards the sensor origin)
    int estimate_normals(int k);
    bool normals_estimated = false;
    void release();
    struct NnIndex index() const;
};

__device__ __forceinline__ unsigned float_to_ordered(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ unsigned long long expand21(unsigned long long v) {
    v &= 0x1fffffull;
    v = (v | v << 32) & 0x1f00000000ffffull;
    v = (v | v << 16) & 0x1f0000ff0000ffull;
    v = (v | v << 8) & 0x100f00f00f00f00full;
    v = (v | v << 4) & 0x10c30c30c30c30c3ull;
    v = (v | v << 2) & 0x1249249249249249ull;
    return v;
}

__device__ __forceinline__ QuantParams make_quant(const unsigned *__restrict__ bbox, int bits) {
    QuantParams q;
    q.lx = ordered_to_float(bbox[0]);
    q.ly = ordered_to_float(bbox[1]);
    q.lz = ordered_to_float(bbox[2]);
    const float ext = fmaxf(fmaxf(ordered_to_float(bbox[3]) - q.lx, ordered_to_float(bbox[4]) - q.ly),
                            ordered_to_float(bbox[5]) - q.lz);
    q.qmax = (float) ((1u << bits) - 1u);
    q.scale = (ext > 0.0f && isfinite(ext)) ? q.qmax / ext : 0.0f;
    q.bits = bits;
    if (!(ext >= 0.0f)) q.lx = q.ly = q.lz = 0.0f;  // empty cloud
    return q;
}

// monotone non-decreasing in v (fp32 subtract, multiply by a non-negative scale, clamp, truncate)
__device__ __forceinline__ unsigned quant_axis(float v, float lo, float scale, float qmax) {
    return (unsigned) fminf(fmaxf(__fmul_rn(__fsub_rn(v, lo), scale), 0.0f), qmax);
}

__device__ __forceinline__ unsigned long long morton_code(const QuantParams &q, float x, float y, float z) {
    return expand21(quant_axis(x, q.lx, q.scale, q.qmax)) | (expand21(quant_axis(y, q.ly, q.scale, q.qmax)) << 1) |
           (expand21(quant_axis(z, q.lz, q.scale, q.qmax)) << 2);
}

constexpr int kStackDepth = 96;  // >= 63 key bits + 27 index bits of radix-tree depth
constexpr int kLinkDone = (int) 0x80000000;

// Leaf link: -1 - (start | (cnt-1) << 27); start < 2^27 points, cnt <= 16.
__host__ __device__ __forceinline__ int make_leaf_link(int start, int cnt) { return -1 - (start | ((cnt - 1) << 27)); }
__device__ __forceinline__ int leaf_start(int link) { return (-1 - link) & 0x7ffffff; }
__device__ __forceinline__ int leaf_count(int link) { return (((-1 - link) >> 27) & 15) + 1; }

__device__ __forceinline__ void scan_leaf(float qx, float qy, float qz, const float4 *__restrict__ pts, int link,
                                          float &best, int &best_idx, int &best_pos) {
    const int start = leaf_start(link), cnt = leaf_count(link);
    float4 p[kLeaf];
#pragma unroll
    for (int k = 0; k < kLeaf; ++k)
        if (k < cnt) p[k] = __ldg(pts + start + k);
#pragma unroll
    for (int k = 0; k < kLeaf; ++k) {
        if (k < cnt) {
            const float d = l2_simple(qx, qy, qz, p[k].x, p[k].y, p[k].z);
            const int idx = __float_as_int(p[k].w);
            if (d < best || (d == best && idx < best_idx)) {
                best = d;
                best_idx = idx;
                best_pos = start + k;
            }
        }
    }
}

// Everything a search needs about the target.
struct NnIndex {
    const TNode *nodes;
    const TreeRoot *root;
    const float4 *pts;                   // Morton-sorted, w = original index
    const CellEntry *cells;              // entry table (kCellBits per axis), or nullptr
    const unsigned *bbox;                // quantisation frame of the Morton keys
    int key_bits;
};

// Ordered top-down search of the subtree under internal node `link` (its box is already known to
// be within the bound).  "while-while" with one pop site: all lanes descend until each holds a
// leaf, has hit a dead end (both children beyond the bound) or is done; then the lanes holding a
// leaf scan it together, and all lanes pop their next subtree together.  (Popping inside the
// descent on a dead end costs a divergent pass of ~2 active lanes per occurrence: -10 % kernel
// time on the 1 M-point lidar workload, profiles/r01_nn_variants.md.)
constexpr int kLinkPop = (int) 0x80000001;  // dead end; leaf links are >= -2^30, so no collision
#if WCU_LEAF == 1
// Single-point leaves: one uniform loop.  A node step measures both children; a point child is
// consumed on the spot (its "box" distance is its exact distance), a subtree child within the bound
// is entered (the nearer one first, the other pushed).  A lane whose step leaves it without a
// subtree pops one stack entry per round; a stale entry (bound tightened since the push) just costs
// that lane one idle round.  Every lane of a warp therefore runs the same few instructions per
// round, whatever mix of descending, point testing and popping the lanes are in.
// distances are >= 0, so (distance bits, original index) compares as one unsigned 64-bit key:
// "closer, or as close with a lower index"
__device__ __forceinline__ unsigned long long nn_key(float d2, unsigned idx) {
    return ((unsigned long long) __float_as_uint(d2) << 32) | idx;
}
__device__ __forceinline__ float key_bound(unsigned long long key) { return __uint_as_float((unsigned) (key >> 32)); }

// The walk proper.  `link`: the first node to step into (kLinkPop: start by popping).  The stack holds
// (box bound, link bits) entries; its newest entry `top` lives in registers, so a pop needs no load on
// its critical path (the entry below it is fetched from local memory for the *next* pop), and its
// bottom is a sentinel that no bound admits.  (slots / top / sp are separate objects on purpose: a
// struct holding the dynamically indexed array would be placed in local memory as a whole.)
__device__ __forceinline__ void walk(float qx, float qy, float qz, const TNode *__restrict__ nodes, int link,
                                     float2 *slots, float2 &top, int &sp, unsigned long long &best_key, int &best_pos) {
    for (;;) {
        if (link >= 0) {
            const float4 a = __ldg(&nodes[link].lo0), b = __ldg(&nodes[link].hi0);
            const float4 c = __ldg(&nodes[link].lo1), d = __ldg(&nodes[link].hi1);
            const int l0 = __float_as_int(a.w), l1 = __float_as_int(c.w);
            const float d0 = aabb_dist(qx, qy, qz, a, b), d1 = aabb_dist(qx, qy, qz, c, d);
            const unsigned long long k0 = nn_key(d0, __float_as_uint(b.w)), k1 = nn_key(d1, __float_as_uint(d.w));
            if (l0 < 0 && k0 < best_key) {
                best_key = k0;
                best_pos = ~l0;
            }
            if (l1 < 0 && k1 < best_key) {
                best_key = k1;
                best_pos = ~l1;
            }
            const float bound = key_bound(best_key);
            const bool in0 = l0 >= 0 && d0 <= bound, in1 = l1 >= 0 && d1 <= bound;
            if (in0 && in1) {
                const bool right_first = d1 < d0;
                slots[sp++] = top;
                top = right_first ? make_float2(d0, a.w) : make_float2(d1, c.w);
                link = right_first ? l1 : l0;
            } else {
                link = in0 ? l0 : (in1 ? l1 : kLinkPop);
            }
        }
        if (link < 0) {
            const int tl = __float_as_int(top.y);
            if (tl == kLinkDone) break;
            link = (top.x <= key_bound(best_key)) ? tl : kLinkPop;
            top = slots[--sp];
        }
    }
}

__device__ __forceinline__ void subtree_search(float qx, float qy, float qz, const TNode *__restrict__ nodes,
                                               const float4 *__restrict__ pts, int link, float &best, int &best_idx,
                                               int &best_pos) {
    (void) pts;
    float2 slots[kStackDepth];
    float2 top = make_float2(INFINITY, __int_as_float(kLinkDone));
    int sp = 0;
    unsigned long long best_key = nn_key(best, (unsigned) best_idx);
    walk(qx, qy, qz, nodes, link, slots, top, sp, best_key, best_pos);
    best = key_bound(best_key);
    best_idx = (int) (unsigned) best_key;
}
#else
__device__ __forceinline__ void subtree_search(float qx, float qy, float qz, const TNode *__restrict__ nodes,
                                               const float4 *__restrict__ pts, int link, float &best, int &best_idx,
                                               int &best_pos) {
    float2 stack[kStackDepth];  // (box bound, link bits): one 8-byte local load per pop
    int sp = 0;
    for (;;) {
        while (link >= 0) {
            const float4 a = __ldg(&nodes[link].lo0), b = __ldg(&nodes[link].hi0);
            const float4 c = __ldg(&nodes[link].lo1), d = __ldg(&nodes[link].hi1);
            const float d0 = aabb_dist(qx, qy, qz, a, b), d1 = aabb_dist(qx, qy, qz, c, d);
            const bool right_first = d1 < d0;
            const float dn = right_first ? d1 : d0, df = right_first ? d0 : d1;
            const float ln = right_first ? c.w : a.w, lf = right_first ? a.w : c.w;
            if (dn <= best) {
                if (df <= best) stack[sp++] = make_float2(df, lf);
                link = __float_as_int(ln);
            } else {
                link = kLinkPop;
            }
        }
        if (link != kLinkPop) {
            if (link == kLinkDone) return;
            scan_leaf(qx, qy, qz, pts, link, best, best_idx, best_pos);
        }
        link = kLinkDone;
        while (sp > 0) {
            const float2 e = stack[--sp];
            if (e.x <= best) {
                link = __float_as_int(e.y);
                break;
            }
        }
    }
}

#endif

inline NnIndex TargetIndex::index() const {
    NnIndex ix;
    ix.nodes = d_nodes;
    ix.root = d_root;
    ix.pts = cloud.d_sorted;
    ix.cells = (kCellBits > 0 && cloud.key_bits >= kCellBits) ? d_cells : nullptr;
    ix.bbox = cloud.d_bbox;
    ix.key_bits = cloud.key_bits;
    return ix;
}

// Exact k-NN (k <= kMaxKnn) under l2_simple, ordered by (distance, original index) - the order
// FLANN's sorted k-NN result set yields up to exact ties.  d[]/idx[]/pos[] are kept sorted; a
// subtree is skipped only when its box bound exceeds the current k-th distance.
constexpr int kMaxKnn = 32;

struct KnnList {
    float d[kMaxKnn];
    int idx[kMaxKnn];
    int pos[kMaxKnn];
    int k;
    __device__ __forceinline__ void init(int k_) {
        k = k_;
        for (int i = 0; i < k_; ++i) {
            d[i] = INFINITY;
            idx[i] = 0x7fffffff;
            pos[i] = -1;
        }
    }
    __device__ __forceinline__ float worst() const { return d[k - 1]; }
    __device__ __forceinline__ void offer(float dist, int index, int position) {
        if (!(dist < d[k - 1] || (dist == d[k - 1] && index < idx[k - 1]))) return;
        int p = k - 1;
        while (p > 0 && (d[p - 1] > dist || (d[p - 1] == dist && idx[p - 1] > index))) {
            d[p] = d[p - 1];
            idx[p] = idx[p - 1];
            pos[p] = pos[p - 1];
            --p;
        }
        d[p] = dist;
        idx[p] = index;
        pos[p] = position;
    }
};

__device__ __forceinline__ void knn_search(float qx, float qy, float qz, const NnIndex &ix, KnnList &out) {
    const float4 rlo = __ldg(&ix.root->lo), rhi = __ldg(&ix.root->hi);
    if (__float_as_int(rhi.w) <= 0) return;
    int link = __float_as_int(rlo.w);
    int stack_link[kStackDepth];
    float stack_d[kStackDepth];
    int sp = 0;
    for (;;) {
        while (link >= 0) {
            const float4 a = __ldg(&ix.nodes[link].lo0), b = __ldg(&ix.nodes[link].hi0);
            const float4 c = __ldg(&ix.nodes[link].lo1), d = __ldg(&ix.nodes[link].hi1);
            const float d0 = aabb_dist(qx, qy, qz, a, b), d1 = aabb_dist(qx, qy, qz, c, d);
            const bool right_first = d1 < d0;
            const float dn = right_first ? d1 : d0, df = right_first ? d0 : d1;
            const int ln = __float_as_int(right_first ? c.w : a.w), lf = __float_as_int(right_first ? a.w : c.w);
            if (dn <= out.worst()) {
                if (df <= out.worst()) {
                    stack_link[sp] = lf;
                    stack_d[sp] = df;
                    ++sp;
                }
                link = ln;
            } else {
                link = kLinkPop;
            }
        }
        if (link != kLinkPop) {
            if (link == kLinkDone) return;
            const int start = leaf_start(link), cnt = leaf_count(link);
            for (int k = 0; k < cnt; ++k) {
                const float4 p = __ldg(ix.pts + start + k);
                out.offer(l2_simple(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w), start + k);
            }
        }
        link = kLinkDone;
        while (sp > 0) {
            --sp;
            if (stack_d[sp] <= out.worst()) {
                link = stack_link[sp];
                break;
            }
        }
    }
}

// Exact 1-NN of (qx,qy,qz) under l2_simple with the lowest original index among exact ties.
// best / best_idx / best_pos come in as the current bound (the max-correspondence threshold with
// best_idx = INT_MAX and best_pos = -1, possibly improved by a warm-start candidate) and leave as
// the result.  A subtree is skipped only when its box bound is strictly greater than best, so
// equal-distance candidates are always examined: the result is exact.
//
// Measured alternatives that did NOT beat this ordered top-down walk on the 1 M-point lidar
// workload (profiles/r01_nn_variants.md): a stack-free threaded traversal (fixed child order costs
// more node visits than the stack saves), a shared-memory stack (the carve-out shrinks L1 and the
// walk is L1-latency bound), and a bottom-up search confined to the Morton cell of the bound
// (balls straddle coarse cell boundaries too often).
__device__ __forceinline__ void nn_search(float qx, float qy, float qz, const NnIndex &ix, float &best, int &best_idx,
                                          int &best_pos) {
    const float4 rlo = __ldg(&ix.root->lo), rhi = __ldg(&ix.root->hi);
    if (__float_as_int(rhi.w) <= 0 || aabb_dist(qx, qy, qz, rlo, rhi) > best) return;
    const int root_link = __float_as_int(rlo.w);
    if (root_link < 0) {
        scan_leaf(qx, qy, qz, ix.pts, root_link, best, best_idx, best_pos);
        return;
    }
    subtree_search(qx, qy, qz, ix.nodes, ix.pts, root_link, best, best_idx, best_pos);
}

#if WCU_LEAF == 1
// 1-NN through the entry table.  A query with a candidate (best = its distance, best_pos >= 0) has
// a ball that usually spans at most two coarse cells per axis, and every target point inside the
// ball then lives under the entries of those <= 8 cells (the quantisation behind the Morton keys is
// monotone per axis, and the radius is grown by more than the fp32 rounding of the per-axis gaps) -
// so the walk starts there, the query's own cell first.  A query without a candidate first searches
// its own cell's subtree, which is exact for that cell and usually yields the neighbour or a tight
// bound, and then only has the other cells of its ball left.  Wider balls, and targets without a
// table, start at the root.
__device__ __forceinline__ void nn_search_cells(float qx, float qy, float qz, const NnIndex &ix, float &best,
                                                int &best_idx, int &best_pos) {
    const float4 rlo = __ldg(&ix.root->lo), rhi = __ldg(&ix.root->hi);
    if (__float_as_int(rhi.w) <= 0) return;
    const int root_link = __float_as_int(rlo.w);
    if (root_link < 0 || ix.cells == nullptr) {
        nn_search(qx, qy, qz, ix, best, best_idx, best_pos);
        return;
    }
    const QuantParams qp = make_quant(ix.bbox, ix.key_bits);
    const int shift = ix.key_bits - kCellBits;
    const int ox = (int) (quant_axis(qx, qp.lx, qp.scale, qp.qmax) >> shift), oy = (int) (quant_axis(qy, qp.ly, qp.scale, qp.qmax) >> shift),
              oz = (int) (quant_axis(qz, qp.lz, qp.scale, qp.qmax) >> shift);
    const int own = ox | (oy << kCellBits) | (oz << (2 * kCellBits));
    float2 slots[kStackDepth];
    float2 top = make_float2(INFINITY, __int_as_float(kLinkDone));
    int sp = 0;
    unsigned long long best_key = nn_key(best, (unsigned) best_idx);
    int link = kLinkPop;
    auto enter = [&](int cell) {
        const int l = __ldg(ix.cells + cell);
        if (l == kCellEmpty) return;
        if (l < 0) {  // a single point
            const float4 p = __ldg(ix.pts + ~l);
            const unsigned long long k = nn_key(l2_simple(qx, qy, qz, p.x, p.y, p.z), __float_as_uint(p.w));
            if (k < best_key) {
                best_key = k;
                best_pos = ~l;
            }
            return;
        }
        if (link >= 0) {
            slots[sp++] = top;
            top = make_float2(0.0f, __int_as_float(link));  // bound 0: always admitted; its node step prunes
        }
        link = l;
    };
    bool own_done = false;
    if (best_pos < 0) {  // no candidate: the own cell first, on its own
        enter(own);
        walk(qx, qy, qz, ix.nodes, link, slots, top, sp, best_key, best_pos);
        own_done = true;
        link = kLinkPop;
        top = make_float2(INFINITY, __int_as_float(kLinkDone));
        sp = 0;
    }
    const float bound = key_bound(best_key);
    const float r = sqrtf(bound) * 1.00001f + fmaxf(fmaxf(fabsf(qx), fabsf(qy)), fabsf(qz)) * 1e-6f + 1e-30f;
    const int ax = (int) (quant_axis(qx - r, qp.lx, qp.scale, qp.qmax) >> shift), bx = (int) (quant_axis(qx + r, qp.lx, qp.scale, qp.qmax) >> shift);
    const int ay = (int) (quant_axis(qy - r, qp.ly, qp.scale, qp.qmax) >> shift), by = (int) (quant_axis(qy + r, qp.ly, qp.scale, qp.qmax) >> shift);
    const int az = (int) (quant_axis(qz - r, qp.lz, qp.scale, qp.qmax) >> shift), bz = (int) (quant_axis(qz + r, qp.lz, qp.scale, qp.qmax) >> shift);
    if (best_pos < 0 || bx - ax > 1 || by - ay > 1 || bz - az > 1) {
        if (aabb_dist(qx, qy, qz, rlo, rhi) <= bound) link = root_link;
    } else {
        for (int cz = az; cz <= bz; ++cz)
            for (int cy = ay; cy <= by; ++cy)
                for (int cx = ax; cx <= bx; ++cx) {
                    const int cell = cx | (cy << kCellBits) | (cz << (2 * kCellBits));
                    if (cell != own) enter(cell);
                }
        if (!own_done) enter(own);  // last = first link to walk
    }
    walk(qx, qy, qz, ix.nodes, link, slots, top, sp, best_key, best_pos);
    best = key_bound(best_key);
    best_idx = (int) (unsigned) best_key;
}
#else
__device__ __forceinline__ void nn_search_cells(float qx, float qy, float qz, const NnIndex &ix, float &best,
                                                int &best_idx, int &best_pos) {
    nn_search(qx, qy, qz, ix, best, best_idx, best_pos);  // tuning builds with multi-point leaves: no entry table
}
#endif

}  // namespace wavecu
