// Build side of index.cuh: bounding box, Morton keys, radix sort, gather, AABB tree.
// The sort is cub::DeviceRadixSort (CUDA toolkit library, build step only - it runs once per
// align() target, not per iteration); everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cuda/atomic>

#include "../../include/wavecu.h"
#include "index.cuh"
#include "linalg.cuh"
#include "voxel.cuh"

namespace wavecu {

namespace {

constexpr int kBuildThreads = 256;
#ifndef WCU_LBVH_FENCE
#define WCU_LBVH_FENCE 0
#endif

__global__ void bbox_init_kernel(unsigned *bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0xff800000u;  // lo = ordered(+inf)
    else if (threadIdx.x < 6) bbox[threadIdx.x] = 0x007fffffu;  // hi = ordered(-inf)
    else if (threadIdx.x < 8) bbox[threadIdx.x] = 0u;  // [6] = number of finite points
}

__global__ void __launch_bounds__(kBuildThreads) bbox_kernel(const float4 *__restrict__ pts, size_t n, unsigned *bbox) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    unsigned cnt = 0;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (finite3(p.x, p.y, p.z)) {
            ++cnt;
            lo[0] = fminf(lo[0], p.x); hi[0] = fmaxf(hi[0], p.x);
            lo[1] = fminf(lo[1], p.y); hi[1] = fmaxf(hi[1], p.y);
            lo[2] = fminf(lo[2], p.z); hi[2] = fmaxf(hi[2], p.z);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    __shared__ float s_lo[kBuildThreads / 32][3], s_hi[kBuildThreads / 32][3];
    __shared__ unsigned s_cnt[kBuildThreads / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            s_lo[warp][d] = lo[d];
            s_hi[warp][d] = hi[d];
        }
        s_cnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        const int d = threadIdx.x;
        float l = s_lo[0][d], h = s_hi[0][d];
        for (int w = 1; w < kBuildThreads / 32; ++w) {
            l = fminf(l, s_lo[w][d]);
            h = fmaxf(h, s_hi[w][d]);
        }
        if (l <= h) {
            atomicMin(&bbox[d], float_to_ordered(l));
            atomicMax(&bbox[3 + d], float_to_ordered(h));
        }
    } else if (threadIdx.x == 3) {
        unsigned c = 0;
        for (int w = 0; w < kBuildThreads / 32; ++w) c += s_cnt[w];
        if (c) atomicAdd(&bbox[6], c);
    }
}

// Packed sort: key << 24 | cloud index in one 64-bit word, sorted on the key bits alone (a stable radix sort keeps
// the indices of equal keys ascending, exactly the order the (key, index) pair sort gives) - a third fewer bytes per
// radix pass (113 against 131 us for 1 M keys, tools/probes/sort_probe.cu).  Clouds of 2^24 points or more, or keys
// wider than 40 bits, take the pair sort.
constexpr int kPackShift = 24;

__global__ void __launch_bounds__(kBuildThreads) morton_kernel(const float4 *__restrict__ pts, size_t n,
                                                               const unsigned *__restrict__ bbox, int bits,
                                                               unsigned long long *keys, unsigned *vals) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const QuantParams qp = make_quant(bbox, bits);
    const float4 p = pts[i];
    unsigned long long key = 1ull << (3 * bits);  // non-finite points sort to the end and become pads
    if (finite3(p.x, p.y, p.z)) key = morton_code(qp, p.x, p.y, p.z);
    if (vals) {
        keys[i] = key;
        vals[i] = (unsigned) i;
    } else {
        keys[i] = (key << kPackShift) | (unsigned long long) i;   // packed: the cloud index rides in the low bits
    }
}

__global__ void __launch_bounds__(kBuildThreads) gather_kernel(const float4 *__restrict__ pts,
                                                               const unsigned long long *keys,   // may alias packed
                                                               const unsigned *vals, size_t n,   // may alias vals_out
                                                               size_t n_pad, int bits, float4 *out,
                                                               const float4 *__restrict__ extra_in, float4 *extra_out,
                                                               unsigned long long *packed, unsigned *vals_out) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    float4 o = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (packed && i < n) {   // packed sort: split the word back into the sorted key and the permutation, in place
        const unsigned long long w = packed[i];
        packed[i] = w >> kPackShift;
        vals_out[i] = (unsigned) (w & ((1ull << kPackShift) - 1ull));
    }
    if (i < n && (keys[i] >> (3 * bits)) == 0ull) {
        const unsigned src = vals[i];
        const float4 p = pts[src];
        o = make_float4(p.x, p.y, p.z, __int_as_float((int) src));
        if (extra_in) e = extra_in[src];
    }
    out[i] = o;
    if (extra_out) extra_out[i] = e;
}

// strict total order on the gaps between consecutive sorted keys: the radix tree is the Cartesian
// tree of this order.  Equal keys fall back to the index bits (a balanced radix tree over the run).
__device__ __forceinline__ bool gap_less(const unsigned long long *__restrict__ keys, int a, int b) {
    const unsigned long long xa = keys[a] ^ keys[a + 1], xb = keys[b] ^ keys[b + 1];
    if (xa != xb) return xa < xb;
    const unsigned ia = (unsigned) a ^ (unsigned) (a + 1), ib = (unsigned) b ^ (unsigned) (b + 1);
    if (ia != ib) return ia < ib;
    return a < b;
}

// Agglomerative LBVH construction, one launch: thread i starts as the single-point node [i,i] and
// climbs.  A node [l,r] hangs under the smaller of its two bounding gaps (l-1 and r); it writes its
// box and link into that parent's record, then exchanges its far bound through other[parent]:
// the first child to arrive retires, the second continues upward with the merged range and box.
// Subtrees of <= kLeaf points are referenced as leaves (runs of the sorted array).
// number of leading key bits (of the 3 * bits Morton bits) two sorted keys share
__device__ __forceinline__ int shared_prefix(unsigned long long a, unsigned long long b, int bits) {
    const unsigned long long x = a ^ b;
    return x ? __clzll((long long) x) - (64 - 3 * bits) : 3 * bits;
}

// (cx, cy, cz) of the coarse cell of a Morton key -> linear index
__device__ __forceinline__ unsigned cell_of_key(unsigned long long key, int bits) {
    const unsigned long long c = key >> (3 * (bits - kCellBits));  // 3 * kCellBits interleaved bits, x lowest
    unsigned cx = 0, cy = 0, cz = 0;
#pragma unroll
    for (int b = 0; b < kCellBits; ++b) {
        cx |= (unsigned) ((c >> (3 * b)) & 1ull) << b;
        cy |= (unsigned) ((c >> (3 * b + 1)) & 1ull) << b;
        cz |= (unsigned) ((c >> (3 * b + 2)) & 1ull) << b;
    }
    return cx | (cy << kCellBits) | (cz << (2 * kCellBits));
}

__global__ void __launch_bounds__(kBuildThreads) lbvh_kernel(const unsigned long long *__restrict__ keys,
                                                             const float4 *__restrict__ pts,
                                                             const unsigned *__restrict__ bbox, TNode *nodes, int *other,
                                                             TreeRoot *root, CellEntry *cells, int bits) {
    const int n = (int) bbox[6];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && n == 0) {
        root->lo = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(-1));
        root->hi = make_float4(-INFINITY, -INFINITY, -INFINITY, __int_as_float(0));
    }
    if (i >= n) return;
    int l = i, r = i;
    const float4 p = pts[i];
    float lox = p.x, loy = p.y, loz = p.z, hix = p.x, hiy = p.y, hiz = p.z;
    int link = make_leaf_link(i, 1);
    float tag = p.w;  // a single-point child carries the point's original index in hi.w
    bool below = cells != nullptr;  // this node still lies inside one coarse cell
    for (;;) {
        // The keys of [l, r] share `own` leading bits; the node is the root of a coarse cell when that is
        // at least the cell prefix and its parent's shared prefix is shorter.  Once a node is above the
        // cell level so are all its ancestors: the long upper part of the climb skips the key loads.
        int own = 0;
        if (below) {
            own = shared_prefix(keys[l], keys[r], bits);
            below = own >= 3 * kCellBits;
        }
        if (l == 0 && r == n - 1) {
            root->lo = make_float4(lox, loy, loz, __int_as_float(link));
            root->hi = make_float4(hix, hiy, hiz, __int_as_float(r - l + 1));
            if (below) cells[cell_of_key(keys[l], bits)] = link;
            return;
        }
        bool parent_right;  // parent is gap r: this node is its left child
        if (l == 0) parent_right = true;
        else if (r == n - 1) parent_right = false;
        else parent_right = gap_less(keys, r, l - 1);
        const int par = parent_right ? r : l - 1;
        if (below && shared_prefix(keys[par], keys[par + 1], bits) < 3 * kCellBits)
            cells[cell_of_key(keys[l], bits)] = link;
        TNode *nd = nodes + par;
        if (parent_right) {
            nd->lo0 = make_float4(lox, loy, loz, __int_as_float(link));
            nd->hi0 = make_float4(hix, hiy, hiz, tag);
        } else {
            nd->lo1 = make_float4(lox, loy, loz, __int_as_float(link));
            nd->hi1 = make_float4(hix, hiy, hiz, tag);
        }
#if WCU_LBVH_FENCE
        __threadfence();
        const int prev = atomicExch(&other[par], parent_right ? l : r);
#else
        // release our half of the record, acquire the sibling's: one acq_rel exchange instead of a
        // sequentially consistent fence + relaxed exchange on every level of the climb
        const int prev = cuda::atomic_ref<int, cuda::thread_scope_device>(other[par]).exchange(
            parent_right ? l : r, cuda::memory_order_acq_rel);
#endif
        if (prev == -1) return;
        float4 slo, shi;
        if (parent_right) {
            r = prev;
            slo = __ldcg(&nd->lo1);
            shi = __ldcg(&nd->hi1);
        } else {
            l = prev;
            slo = __ldcg(&nd->lo0);
            shi = __ldcg(&nd->hi0);
        }
        lox = fminf(lox, slo.x); loy = fminf(loy, slo.y); loz = fminf(loz, slo.z);
        hix = fmaxf(hix, shi.x); hiy = fmaxf(hiy, shi.y); hiz = fmaxf(hiz, shi.z);
        link = (r - l + 1 <= kLeaf) ? make_leaf_link(l, r - l + 1) : par;
        tag = 0.0f;
    }
}

// Surface normal of every Morton-sorted target point from its k nearest neighbours (itself
// included): fp64 covariance, 3x3 Jacobi, direction of least variance, flipped towards the sensor
// origin (pcl::flipNormalTowardsViewpoint with the default viewpoint).  Used by the point-to-plane
// estimator when the caller supplies no normals.
__global__ void __launch_bounds__(128) normals_kernel(NnIndex ix, int n, int k, float4 *nrm) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const float4 q = ix.pts[s];
    if (__float_as_int(q.w) == 0x7fffffff) {
        nrm[s] = make_float4(NAN, NAN, NAN, 0.f);
        return;
    }
    KnnList nb;
    nb.init(k);
    knn_search(q.x, q.y, q.z, ix, nb);
    double mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    int m = 0;
    for (int j = 0; j < k; ++j) {
        if (nb.pos[j] < 0) break;
        const float4 p = __ldg(ix.pts + nb.pos[j]);
        const double x = p.x, y = p.y, z = p.z;
        mean[0] += x; mean[1] += y; mean[2] += z;
        cov[0] += x * x; cov[1] += x * y; cov[2] += x * z; cov[4] += y * y; cov[5] += y * z; cov[8] += z * z;
        ++m;
    }
    if (m < 3) {
        nrm[s] = make_float4(NAN, NAN, NAN, 0.f);
        return;
    }
    for (int d = 0; d < 3; ++d) mean[d] /= (double) m;
    cov[0] = cov[0] / m - mean[0] * mean[0];
    cov[1] = cov[1] / m - mean[0] * mean[1];
    cov[2] = cov[2] / m - mean[0] * mean[2];
    cov[4] = cov[4] / m - mean[1] * mean[1];
    cov[5] = cov[5] / m - mean[1] * mean[2];
    cov[8] = cov[8] / m - mean[2] * mean[2];
    cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
    double U[9];
    eig_sym3_desc(cov, U);
    double nx = U[2], ny = U[5], nz = U[8];  // column of the smallest |eigenvalue|
    if (nx * (double) q.x + ny * (double) q.y + nz * (double) q.z > 0) {
        nx = -nx; ny = -ny; nz = -nz;
    }
    nrm[s] = make_float4((float) nx, (float) ny, (float) nz, 0.f);
}

// Box pyramid, levels 0 and 1: one thread per run of 8 sorted points (its 128 B are contiguous), then
// the 8 threads of a 64-point run combine their boxes with shuffles.  n_pad is a multiple of 64.
__global__ void __launch_bounds__(kBuildThreads) boxes_low_kernel(const float4 *__restrict__ pts, int n_runs8,
                                                                  float4 *lv0, float4 *lv1) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // n_runs8 is a multiple of 8: whole groups of 8 lanes exit together
    if (i >= n_runs8) return;
    float lox = INFINITY, loy = INFINITY, loz = INFINITY, hix = -INFINITY, hiy = -INFINITY, hiz = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 p = pts[(size_t) i * 8 + k];
        lox = fminf(lox, p.x); loy = fminf(loy, p.y); loz = fminf(loz, p.z);
        hix = fmaxf(hix, p.x); hiy = fmaxf(hiy, p.y); hiz = fmaxf(hiz, p.z);
    }
    lv0[2 * (size_t) i] = make_float4(lox, loy, loz, 0.f);
    lv0[2 * (size_t) i + 1] = make_float4(hix, hiy, hiz, 0.f);
    const unsigned group = 0xffu << ((threadIdx.x & 31) & ~7);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        lox = fminf(lox, __shfl_xor_sync(group, lox, o)); loy = fminf(loy, __shfl_xor_sync(group, loy, o));
        loz = fminf(loz, __shfl_xor_sync(group, loz, o)); hix = fmaxf(hix, __shfl_xor_sync(group, hix, o));
        hiy = fmaxf(hiy, __shfl_xor_sync(group, hiy, o)); hiz = fmaxf(hiz, __shfl_xor_sync(group, hiz, o));
    }
    if ((i & 7) == 0) {
        lv1[2 * (size_t) (i >> 3)] = make_float4(lox, loy, loz, 0.f);
        lv1[2 * (size_t) (i >> 3) + 1] = make_float4(hix, hiy, hiz, 0.f);
    }
}

// Levels >= 2, one level per launch (they shrink by 8 each: a handful of short launches inside the
// build graph); a missing child (the last box of a level) is skipped.
__global__ void __launch_bounds__(kBuildThreads) boxes_up_kernel(const float4 *__restrict__ child, int n_child,
                                                                 float4 *parent, int n_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_parent) return;
    float lox = INFINITY, loy = INFINITY, loz = INFINITY, hix = -INFINITY, hiy = -INFINITY, hiz = -INFINITY;
    for (int k = 0; k < 8; ++k) {
        const int c = i * 8 + k;
        if (c >= n_child) break;
        const float4 lo = child[2 * (size_t) c], hi = child[2 * (size_t) c + 1];
        lox = fminf(lox, lo.x); loy = fminf(loy, lo.y); loz = fminf(loz, lo.z);
        hix = fmaxf(hix, hi.x); hiy = fmaxf(hiy, hi.y); hiz = fmaxf(hiz, hi.z);
    }
    parent[2 * (size_t) i] = make_float4(lox, loy, loz, 0.f);
    parent[2 * (size_t) i + 1] = make_float4(hix, hiy, hiz, 0.f);
}

template <class T>
int grow(T *&ptr, size_t &cap, size_t want) {
    if (want <= cap) return WAVECU_OK;
    if (ptr) WCU_CHECK(cudaFree(ptr));
    ptr = nullptr;
    cap = 0;
    const size_t alloc = want + want / 8 + 64;
    WCU_CHECK(cudaMalloc(reinterpret_cast<void **>(&ptr), alloc * sizeof(T)));
    cap = alloc;
    return WAVECU_OK;
}

}  // namespace

void launch_bbox(const float4 *d_pts, size_t n, unsigned *d_bbox8, cudaStream_t stream) {
    bbox_init_kernel<<<1, 32, 0, stream>>>(d_bbox8);
    if (n) {
        const int grid = (int) std::min<size_t>((n + kBuildThreads - 1) / kBuildThreads, 148 * 4);
        bbox_kernel<<<grid, kBuildThreads, 0, stream>>>(d_pts, n, d_bbox8);
    }
}

int MortonCloud::reserve(size_t n_points, size_t sorted_points) {
    if (n_points > cap) {
        for (void *p : {(void *) d_raw, (void *) d_keys, (void *) d_keys_alt, (void *) d_vals, (void *) d_vals_alt})
            if (p) WCU_CHECK(cudaFree(p));
        d_raw = nullptr; d_keys = d_keys_alt = nullptr; d_vals = d_vals_alt = nullptr;
        cap = 0;
        const size_t a = n_points + n_points / 8 + 64;
        WCU_CHECK(cudaMalloc((void **) &d_raw, a * sizeof(float4)));
        WCU_CHECK(cudaMalloc((void **) &d_keys, a * sizeof(unsigned long long)));
        WCU_CHECK(cudaMalloc((void **) &d_keys_alt, a * sizeof(unsigned long long)));
        WCU_CHECK(cudaMalloc((void **) &d_vals, a * sizeof(unsigned)));
        WCU_CHECK(cudaMalloc((void **) &d_vals_alt, a * sizeof(unsigned)));
        cap = a;
        size_t need = 0;
        cub::DoubleBuffer<unsigned long long> kb(d_keys, d_keys_alt);
        cub::DoubleBuffer<unsigned> vb(d_vals, d_vals_alt);
        WCU_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, need, kb, vb, (int) a, 0, 64, stream));
        if (need > tmp_bytes) {
            if (d_tmp) WCU_CHECK(cudaFree(d_tmp));
            d_tmp = nullptr;
            WCU_CHECK(cudaMalloc(&d_tmp, need));
            tmp_bytes = need;
        }
    }
    if (!d_bbox) WCU_CHECK(cudaMalloc((void **) &d_bbox, 8 * sizeof(unsigned)));
    if (sorted_points > sorted_cap) {
        if (d_sorted) WCU_CHECK(cudaFree(d_sorted));
        d_sorted = nullptr;
        sorted_cap = 0;
        const size_t a = sorted_points + 64;
        WCU_CHECK(cudaMalloc((void **) &d_sorted, a * sizeof(float4)));
        sorted_cap = a;
    }
    return WAVECU_OK;
}

int MortonCloud::upload(const float *xyzw, size_t n_points, bool from_device) {
    WCU_CHECK(cudaSetDevice(device));
    int rc = reserve(n_points, 0);
    if (rc) return rc;
    n = n_points;
    // host clouds cross PCIe on the copy stream; device-resident ones are a ~10 us D2D copy that simply
    // goes in front of the sort on `stream` (which the caller has ordered behind the producer)
    const bool async_copy = copy_stream && !from_device;
    cudaStream_t cs = async_copy ? copy_stream : stream;
    if (async_copy) {
        if (!ev_up) {
            const unsigned fl = getenv("WAVECU_TIMELINE") ? cudaEventDefault : cudaEventDisableTiming;
            WCU_CHECK(cudaEventCreateWithFlags(&ev_up, fl));
            WCU_CHECK(cudaEventCreateWithFlags(&ev_used, fl));
        }
        if (used_pending) WCU_CHECK(cudaStreamWaitEvent(copy_stream, ev_used, 0));
    }
    if (n)
        WCU_CHECK(cudaMemcpyAsync(d_raw, xyzw, n * sizeof(float4),
                                  from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, cs));
    up_pending = false;
    if (async_copy) {
        WCU_CHECK(cudaEventRecord(ev_up, copy_stream));
        up_pending = true;
    }
    return WAVECU_OK;
}

static bool pack_enabled() {
    static const bool on = [] {
        const char *e = getenv("WAVECU_PACKED_SORT");
        return !(e && *e == '0');
    }();
    return on;
}

bool graphs_enabled() {
    static const bool on = [] {
        const char *e = getenv("WAVECU_NO_GRAPH");
        return !(e && *e && *e != '0');
    }();
    return on;
}

namespace {
// Replays `cache` if its key still matches, otherwise captures `enqueue` (thread-local capture: other
// host threads keep using CUDA) and instantiates it.  `counter` is the owner's launch statistic.
template <class F>
int run_captured(GraphCache &cache, cudaStream_t stream, const unsigned long long (&key)[8], long long &counter,
                 F &&enqueue) {
    if (!graphs_enabled()) return enqueue();
    if (!cache.matches(key)) {
        if (!cache.seen_before(key)) return enqueue();   // first sight of this key: plain launches
        cache.release();
        const long long before = counter;
        WCU_CHECK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue();
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(stream, &graph);
        const long long per_replay = counter - before;
        counter = before;
        if (rc != WAVECU_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            if (rc != WAVECU_OK) return rc;
            return enqueue();  // not capturable here: plain launches
        }
        const cudaError_t ie = cudaGraphInstantiate(&cache.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            cache.exec = nullptr;
            cudaGetLastError();
            return enqueue();
        }
        for (int i = 0; i < 8; ++i) cache.key[i] = key[i];
        cache.launches = per_replay;
    }
    WCU_CHECK(cudaGraphLaunch(cache.exec, stream));
    counter += cache.launches;
    return WAVECU_OK;
}
}  // namespace

int MortonCloud::enqueue_sort(size_t n_sorted_pad, const float4 *d_extra_in, float4 *d_extra_out) {
    // WAVECU_TIMELINE + WAVECU_NO_GRAPH: events between the phases (not capturable into the cached graph)
    const bool timeline = getenv("WAVECU_TIMELINE") && !graphs_enabled();
    if (timeline && !ev_tl[0])
        for (auto &e : ev_tl) WCU_CHECK(cudaEventCreate(&e));
    bbox_init_kernel<<<1, 32, 0, stream>>>(d_bbox);
    ++launches;
    d_keys_sorted = d_keys;
    d_vals_sorted = d_vals;
    bool packed = false;
    if (n) {
        const int grid = (int) std::min<size_t>((n + kBuildThreads - 1) / kBuildThreads, 148 * 4);
        bbox_kernel<<<grid, kBuildThreads, 0, stream>>>(d_raw, n, d_bbox);
        packed = pack_enabled() && n_sorted_pad >= n && n < ((size_t) 1 << kPackShift) && 3 * key_bits + 1 + kPackShift <= 64;
        morton_kernel<<<(unsigned) ((n + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0, stream>>>(
            d_raw, n, d_bbox, key_bits, d_keys, packed ? nullptr : d_vals);
        launches += 2;
        if (timeline) WCU_CHECK(cudaEventRecord(ev_tl[0], stream));
        cub::DoubleBuffer<unsigned long long> kb(d_keys, d_keys_alt);
        cub::DoubleBuffer<unsigned> vb(d_vals, d_vals_alt);
        size_t need = tmp_bytes;
        if (packed)
            WCU_CHECK(cub::DeviceRadixSort::SortKeys(d_tmp, need, kb, (int) n, kPackShift, kPackShift + 3 * key_bits + 1, stream));
        else
            WCU_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, need, kb, vb, (int) n, 0, 3 * key_bits + 1, stream));
        launches += 2 + (3 * key_bits + 8) / 8;  // histogram + scan + onesweep passes (CUB-internal)
        d_keys_sorted = kb.Current();
        d_vals_sorted = vb.Current();   // packed: the gather below fills it
        if (timeline) WCU_CHECK(cudaEventRecord(ev_tl[1], stream));
    }
    if (n_sorted_pad) {
        gather_kernel<<<(unsigned) ((n_sorted_pad + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0, stream>>>(
            d_raw, d_keys_sorted, d_vals_sorted, n, n_sorted_pad, key_bits, d_sorted, d_extra_in, d_extra_out,
            packed ? d_keys_sorted : nullptr, packed ? d_vals_sorted : nullptr);
        ++launches;
        if (timeline) WCU_CHECK(cudaEventRecord(ev_tl[2], stream));
    }
    WCU_CHECK(cudaGetLastError());
    return WAVECU_OK;
}

int MortonCloud::pre_sort(size_t n_sorted_pad) {
    int rc = reserve(n, n_sorted_pad);
    if (rc) return rc;
    if (copy_stream && up_pending) WCU_CHECK(cudaStreamWaitEvent(stream, ev_up, 0));
    return WAVECU_OK;
}

int MortonCloud::post_sort() {
    if (copy_stream && ev_used) {  // the next host upload must not overwrite d_raw under this sort
        WCU_CHECK(cudaEventRecord(ev_used, stream));
        used_pending = true;
    }
    return WAVECU_OK;
}

int MortonCloud::sort(size_t n_sorted_pad, const float4 *d_extra_in, float4 *d_extra_out) {
    int rc = pre_sort(n_sorted_pad);
    if (rc) return rc;
    const unsigned long long key[8] = {(unsigned long long) n, (unsigned long long) n_sorted_pad,
                                       (unsigned long long) (uintptr_t) d_raw, (unsigned long long) (uintptr_t) d_sorted,
                                       (unsigned long long) (uintptr_t) d_keys, (unsigned long long) (uintptr_t) d_extra_in,
                                       (unsigned long long) (uintptr_t) d_extra_out, (unsigned long long) key_bits};
    rc = run_captured(sort_graph, stream, key, launches,
                      [&] { return enqueue_sort(n_sorted_pad, d_extra_in, d_extra_out); });
    if (rc) return rc;
    return post_sort();
}

void MortonCloud::release() {
    sort_graph.release();
    if (ev_up) cudaEventDestroy(ev_up);
    if (ev_used) cudaEventDestroy(ev_used);
    for (auto &e : ev_tl) {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    ev_up = ev_used = nullptr;
    up_pending = used_pending = false;
    for (void *p : {(void *) d_raw, (void *) d_sorted, (void *) d_bbox, (void *) d_keys, (void *) d_keys_alt,
                    (void *) d_vals, (void *) d_vals_alt, d_tmp})
        if (p) cudaFree(p);
    d_raw = d_sorted = nullptr; d_bbox = nullptr; d_keys = d_keys_alt = nullptr; d_vals = d_vals_alt = nullptr;
    d_keys_sorted = nullptr; d_vals_sorted = nullptr;
    d_tmp = nullptr;
    cap = sorted_cap = tmp_bytes = n = 0;
}

int TargetIndex::set_points(const float *xyzw, size_t n, bool from_device) {
    dirty = true;
    nrm_n = 0;
    normals_estimated = false;
    return cloud.upload(xyzw, n, from_device);
}

int TargetIndex::set_normals(const float *nxyzw, size_t n, bool from_device) {
    WCU_CHECK(cudaSetDevice(cloud.device));
    if (n != cloud.n) {
        set_last_error("normals count differs from the target size");
        return WAVECU_ERR_ARG;
    }
    int rc = grow(d_nrm_raw, nrm_cap, n);
    if (rc) return rc;
    const bool async_copy = cloud.copy_stream && !from_device;
    cudaStream_t cs = async_copy ? cloud.copy_stream : cloud.stream;
    if (n)
        WCU_CHECK(cudaMemcpyAsync(d_nrm_raw, nxyzw, n * sizeof(float4),
                                  from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, cs));
    nrm_up_pending = false;
    if (async_copy) {
        if (!ev_nrm_up)
            WCU_CHECK(cudaEventCreateWithFlags(&ev_nrm_up, getenv("WAVECU_TIMELINE") ? cudaEventDefault
                                                                                    : cudaEventDisableTiming));
        WCU_CHECK(cudaEventRecord(ev_nrm_up, cloud.copy_stream));
        nrm_up_pending = true;
    }
    nrm_n = n;
    nrm_dirty = true;   // the tree does not depend on the normals: only their gather is redone
    return WAVECU_OK;
}

namespace {
__global__ void __launch_bounds__(kBuildThreads) gather_extra_kernel(const float4 *__restrict__ in,
                                                                     const unsigned long long *__restrict__ keys,
                                                                     const unsigned *__restrict__ vals, size_t n,
                                                                     int bits, float4 *out) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((keys[i] >> (3 * bits)) == 0ull) e = in[vals[i]];
    out[i] = e;
}
}  // namespace

int TargetIndex::reserve_sorted_normals() {
    const size_t n = cloud.n;
    if (n > nrm_sorted_cap) {
        if (d_nrm_sorted) WCU_CHECK(cudaFree(d_nrm_sorted));
        d_nrm_sorted = nullptr;
        WCU_CHECK(cudaMalloc((void **) &d_nrm_sorted, (n + n / 8 + 64) * sizeof(float4)));
        nrm_sorted_cap = n + n / 8 + 64;
    }
    return WAVECU_OK;
}

int TargetIndex::sort_normals() {
    WCU_CHECK(cudaSetDevice(cloud.device));
    const size_t n = cloud.n;
    if (nrm_n != n) {
        nrm_dirty = false;
        return WAVECU_OK;
    }
    const int rc = reserve_sorted_normals();
    if (rc) return rc;
    if (cloud.copy_stream && nrm_up_pending) WCU_CHECK(cudaStreamWaitEvent(cloud.stream, ev_nrm_up, 0));
    if (n) {
        gather_extra_kernel<<<(unsigned) ((n + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0, cloud.stream>>>(
            d_nrm_raw, cloud.d_keys_sorted, cloud.d_vals_sorted, n, cloud.key_bits, d_nrm_sorted);
        ++cloud.launches;
        WCU_CHECK(cudaGetLastError());
    }
    nrm_dirty = false;
    normals_estimated = false;
    return WAVECU_OK;
}

int TargetIndex::build() {
    WCU_CHECK(cudaSetDevice(cloud.device));
    const size_t n = cloud.n;
    if (n >= ((size_t) 1 << 27)) {
        set_last_error("target cloud too large for 27-bit leaf links (>= 134M points)");
        return WAVECU_ERR_ARG;
    }
    if (n + 1 > node_cap) {
        if (d_nodes) WCU_CHECK(cudaFree(d_nodes));
        if (d_other) WCU_CHECK(cudaFree(d_other));
        d_nodes = nullptr; d_other = nullptr; node_cap = 0;
        const size_t a = n + n / 8 + 64;
        WCU_CHECK(cudaMalloc((void **) &d_nodes, a * sizeof(TNode)));
        WCU_CHECK(cudaMalloc((void **) &d_other, a * sizeof(int)));
        node_cap = a;
    }
    if (!d_root) WCU_CHECK(cudaMalloc((void **) &d_root, sizeof(TreeRoot)));
    if (kCellBits > 0 && !d_cells) WCU_CHECK(cudaMalloc((void **) &d_cells, kCellCount * sizeof(CellEntry)));
    int rc = cloud.pre_sort(std::max<size_t>(sorted_pad(), 1));
    if (rc) return rc;
    if (want_boxes) {
        // level sizes: 8-point runs, then /8 per level until at most kBoxTop boxes (at least two levels)
        size_t total = 0;
        int cnt[kBoxLevelsMax], nl = 0;
        size_t c = sorted_pad() / 8;
        for (;;) {
            cnt[nl++] = (int) c;
            total += 2 * c;
            if ((nl >= 2 && c <= (size_t) kBoxTop) || nl == kBoxLevelsMax) break;
            c = (c + 7) / 8;
        }
        if (cnt[nl - 1] > kBoxTop) {
            set_last_error("target cloud too large for the box pyramid");
            return WAVECU_ERR_ARG;
        }
        if (total > boxes_cap) {
            if (d_boxes) WCU_CHECK(cudaFree(d_boxes));
            d_boxes = nullptr;
            boxes_cap = 0;
            const size_t a = total + total / 8 + 64;
            WCU_CHECK(cudaMalloc((void **) &d_boxes, a * sizeof(float4)));
            boxes_cap = a;
        }
        size_t off = 0;
        for (int k = 0; k < kBoxLevelsMax; ++k) {
            boxes.lv[k] = k < nl ? d_boxes + off : nullptr;
            boxes.cnt[k] = k < nl ? cnt[k] : 0;
            if (k < nl) off += 2 * (size_t) cnt[k];
        }
        boxes.n_levels = nl;
    }
    const unsigned long long key[8] = {(unsigned long long) n + ((unsigned long long) want_boxes << 62),
                                       (unsigned long long) (uintptr_t) cloud.d_raw ^ ((unsigned long long) (uintptr_t) d_boxes << 1),
                                       (unsigned long long) (uintptr_t) cloud.d_sorted,
                                       (unsigned long long) (uintptr_t) cloud.d_keys,
                                       (unsigned long long) (uintptr_t) d_nodes, (unsigned long long) (uintptr_t) d_other,
                                       (unsigned long long) (uintptr_t) d_root, (unsigned long long) cloud.key_bits};
    rc = run_captured(build_graph, cloud.stream, key, cloud.launches, [&] { return enqueue_build(); });
    if (rc) return rc;
    rc = cloud.post_sort();
    if (rc) return rc;
    if (nrm_n) nrm_dirty = true;
    dirty = false;
    normals_estimated = false;
    return WAVECU_OK;
}

// Morton sort + one-launch radix-tree construction, as one capturable launch sequence
int TargetIndex::enqueue_build() {
    const size_t n = cloud.n;
    int rc = cloud.enqueue_sort(std::max<size_t>(sorted_pad(), 1), nullptr, nullptr);
    if (rc) return rc;
    if (want_boxes && n) {
        const int runs8 = boxes.cnt[0];
        boxes_low_kernel<<<(unsigned) ((runs8 + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0, cloud.stream>>>(
            cloud.d_sorted, runs8, const_cast<float4 *>(boxes.lv[0]), const_cast<float4 *>(boxes.lv[1]));
        ++cloud.launches;
        for (int k = 2; k < boxes.n_levels; ++k) {
            boxes_up_kernel<<<(unsigned) ((boxes.cnt[k] + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0,
                              cloud.stream>>>(boxes.lv[k - 1], boxes.cnt[k - 1], const_cast<float4 *>(boxes.lv[k]),
                                              boxes.cnt[k]);
            ++cloud.launches;
        }
    }
    if (n) WCU_CHECK(cudaMemsetAsync(d_other, 0xff, n * sizeof(int), cloud.stream));
    CellEntry *cells = (kCellBits > 0 && cloud.key_bits >= kCellBits) ? d_cells : nullptr;
    if (cells) WCU_CHECK(cudaMemsetAsync(cells, 0x80, kCellCount * sizeof(CellEntry), cloud.stream));  // kCellEmpty
    lbvh_kernel<<<(unsigned) std::max<size_t>(1, (n + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0,
                  cloud.stream>>>(cloud.d_keys_sorted, cloud.d_sorted, cloud.d_bbox, d_nodes, d_other, d_root, cells,
                                  cloud.key_bits);
    ++cloud.launches;
    WCU_CHECK(cudaGetLastError());
    return WAVECU_OK;
}

int TargetIndex::estimate_normals(int k) {
    WCU_CHECK(cudaSetDevice(cloud.device));
    const size_t n = cloud.n;
    if (n > nrm_sorted_cap) {
        if (d_nrm_sorted) WCU_CHECK(cudaFree(d_nrm_sorted));
        d_nrm_sorted = nullptr;
        WCU_CHECK(cudaMalloc((void **) &d_nrm_sorted, (n + n / 8 + 64) * sizeof(float4)));
        nrm_sorted_cap = n + n / 8 + 64;
    }
    if (n) {
        k = std::max(3, std::min(k, kMaxKnn));
        normals_kernel<<<(unsigned) ((n + 127) / 128), 128, 0, cloud.stream>>>(index(), (int) n, k, d_nrm_sorted);
        ++cloud.launches;
        WCU_CHECK(cudaGetLastError());
    }
    normals_estimated = true;
    return WAVECU_OK;
}

void TargetIndex::release() {
    build_graph.release();
    cloud.release();
    for (void *p : {(void *) d_nodes, (void *) d_other, (void *) d_root, (void *) d_nrm_raw, (void *) d_nrm_sorted,
                    (void *) d_cells, (void *) d_boxes})
        if (p) cudaFree(p);
    d_cells = nullptr;
    d_boxes = nullptr;
    boxes_cap = 0;
    if (ev_nrm_up) cudaEventDestroy(ev_nrm_up);
    ev_nrm_up = nullptr;
    nrm_up_pending = nrm_dirty = false;
    d_nodes = nullptr; d_other = nullptr; d_root = nullptr; d_nrm_raw = d_nrm_sorted = nullptr;
    node_cap = nrm_cap = nrm_sorted_cap = nrm_n = 0;
}

}  // namespace wavecu
