// Build side of index.cuh: bounding box, Morton keys, radix sort, gather, AABB tree.
// The sort is cub::DeviceRadixSort (CUDA toolkit library, build step only - it runs once per
// align() target, not per iteration); everything else is hand-written.
#include <cub/device/device_radix_sort.cuh>

#include "../../include/wavecu.h"
#include "index.cuh"

namespace wavecu {

namespace {

constexpr int kBuildThreads = 256;

__global__ void bbox_init_kernel(unsigned *bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0xff800000u;  // lo = ordered(+inf)
    else if (threadIdx.x < 6) bbox[threadIdx.x] = 0x007fffffu;  // hi = ordered(-inf)
}

__global__ void __launch_bounds__(kBuildThreads) bbox_kernel(const float4 *__restrict__ pts, size_t n, unsigned *bbox) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        if (finite3(p.x, p.y, p.z)) {
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
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (lo[d] <= hi[d]) {
                atomicMin(&bbox[d], float_to_ordered(lo[d]));
                atomicMax(&bbox[3 + d], float_to_ordered(hi[d]));
            }
        }
    }
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

__global__ void __launch_bounds__(kBuildThreads) morton_kernel(const float4 *__restrict__ pts, size_t n,
                                                               const unsigned *__restrict__ bbox,
                                                               unsigned long long *keys, unsigned *vals) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float lx = ordered_to_float(bbox[0]), ly = ordered_to_float(bbox[1]), lz = ordered_to_float(bbox[2]);
    const float ext = fmaxf(fmaxf(ordered_to_float(bbox[3]) - lx, ordered_to_float(bbox[4]) - ly),
                            ordered_to_float(bbox[5]) - lz);
    const float scale = (ext > 0.0f && isfinite(ext)) ? 2097151.0f / ext : 0.0f;
    const float4 p = pts[i];
    unsigned long long key = ~0ull;  // non-finite points sort to the end and become pads
    if (finite3(p.x, p.y, p.z)) {
        const unsigned qx = (unsigned) fminf(fmaxf((p.x - lx) * scale, 0.0f), 2097151.0f);
        const unsigned qy = (unsigned) fminf(fmaxf((p.y - ly) * scale, 0.0f), 2097151.0f);
        const unsigned qz = (unsigned) fminf(fmaxf((p.z - lz) * scale, 0.0f), 2097151.0f);
        key = expand21(qx) | (expand21(qy) << 1) | (expand21(qz) << 2);
    }
    keys[i] = key;
    vals[i] = (unsigned) i;
}

__global__ void __launch_bounds__(kBuildThreads) gather_kernel(const float4 *__restrict__ pts,
                                                               const unsigned long long *__restrict__ keys,
                                                               const unsigned *__restrict__ vals, size_t n,
                                                               size_t n_pad, float4 *out,
                                                               const float4 *__restrict__ extra_in, float4 *extra_out) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    float4 o = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7fffffff));
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n && keys[i] != ~0ull) {
        const unsigned src = vals[i];
        const float4 p = pts[src];
        o = make_float4(p.x, p.y, p.z, __int_as_float((int) src));
        if (extra_in) e = extra_in[src];
    }
    out[i] = o;
    if (extra_out) extra_out[i] = e;
}

// One thread per leaf slot: leaf AABB, then climb; the second thread to reach a parent merges.
__global__ void __launch_bounds__(kBuildThreads) tree_kernel(const float4 *__restrict__ pts, int P, Node *nodes,
                                                             int *flags) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= P) return;
    float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
#pragma unroll
    for (int k = 0; k < kLeaf; ++k) {
        const float4 p = pts[(size_t) j * kLeaf + k];
        if (p.x != INFINITY) {
            lo.x = fminf(lo.x, p.x); hi.x = fmaxf(hi.x, p.x);
            lo.y = fminf(lo.y, p.y); hi.y = fmaxf(hi.y, p.y);
            lo.z = fminf(lo.z, p.z); hi.z = fmaxf(hi.z, p.z);
        }
    }
    unsigned node = (unsigned) (P + j);
    nodes[node].lo = lo;
    nodes[node].hi = hi;
    while (node > 1u) {
        const unsigned parent = node >> 1;
        __threadfence();
        if (atomicAdd(&flags[parent], 1) == 0) return;  // sibling not ready yet: it will merge
        const unsigned sib = node ^ 1u;
        const float4 slo = __ldcg(&nodes[sib].lo), shi = __ldcg(&nodes[sib].hi);
        lo.x = fminf(lo.x, slo.x); lo.y = fminf(lo.y, slo.y); lo.z = fminf(lo.z, slo.z);
        hi.x = fmaxf(hi.x, shi.x); hi.y = fmaxf(hi.y, shi.y); hi.z = fmaxf(hi.z, shi.z);
        nodes[parent].lo = lo;
        nodes[parent].hi = hi;
        node = parent;
    }
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
    if (!d_bbox) WCU_CHECK(cudaMalloc((void **) &d_bbox, 6 * sizeof(unsigned)));
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
    if (n)
        WCU_CHECK(cudaMemcpyAsync(d_raw, xyzw, n * sizeof(float4),
                                  from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, stream));
    return WAVECU_OK;
}

int MortonCloud::sort(size_t n_sorted_pad, const float4 *d_extra_in, float4 *d_extra_out) {
    int rc = reserve(n, n_sorted_pad);
    if (rc) return rc;
    bbox_init_kernel<<<1, 32, 0, stream>>>(d_bbox);
    ++launches;
    if (n) {
        const int grid = (int) std::min<size_t>((n + kBuildThreads - 1) / kBuildThreads, 148 * 8);
        bbox_kernel<<<grid, kBuildThreads, 0, stream>>>(d_raw, n, d_bbox);
        morton_kernel<<<(unsigned) ((n + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0, stream>>>(
            d_raw, n, d_bbox, d_keys, d_vals);
        launches += 2;
        cub::DoubleBuffer<unsigned long long> kb(d_keys, d_keys_alt);
        cub::DoubleBuffer<unsigned> vb(d_vals, d_vals_alt);
        size_t need = tmp_bytes;
        WCU_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, need, kb, vb, (int) n, 0, 64, stream));
        launches += 8;  // histogram + onesweep passes (approximate; CUB-internal)
        if (kb.Current() != d_keys) std::swap(d_keys, d_keys_alt);
        if (vb.Current() != d_vals) std::swap(d_vals, d_vals_alt);
    }
    if (n_sorted_pad) {
        gather_kernel<<<(unsigned) ((n_sorted_pad + kBuildThreads - 1) / kBuildThreads), kBuildThreads, 0, stream>>>(
            d_raw, d_keys, d_vals, n, n_sorted_pad, d_sorted, d_extra_in, d_extra_out);
        ++launches;
    }
    WCU_CHECK(cudaGetLastError());
    return WAVECU_OK;
}

void MortonCloud::release() {
    for (void *p : {(void *) d_raw, (void *) d_sorted, (void *) d_bbox, (void *) d_keys, (void *) d_keys_alt,
                    (void *) d_vals, (void *) d_vals_alt, d_tmp})
        if (p) cudaFree(p);
    d_raw = d_sorted = nullptr; d_bbox = nullptr; d_keys = d_keys_alt = nullptr; d_vals = d_vals_alt = nullptr;
    d_tmp = nullptr;
    cap = sorted_cap = tmp_bytes = n = 0;
}

int TargetIndex::set_points(const float *xyzw, size_t n, bool from_device) {
    dirty = true;
    nrm_n = 0;
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
    if (n)
        WCU_CHECK(cudaMemcpyAsync(d_nrm_raw, nxyzw, n * sizeof(float4),
                                  from_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, cloud.stream));
    nrm_n = n;
    dirty = true;
    return WAVECU_OK;
}

int TargetIndex::build() {
    WCU_CHECK(cudaSetDevice(cloud.device));
    const size_t leaves = (cloud.n + kLeaf - 1) / kLeaf;
    int p = 1;
    while ((size_t) p < leaves) p <<= 1;
    P = p;
    const size_t n_pad = (size_t) P * kLeaf;
    if ((size_t) 2 * P > node_cap) {
        if (d_nodes) WCU_CHECK(cudaFree(d_nodes));
        if (d_flags) WCU_CHECK(cudaFree(d_flags));
        d_nodes = nullptr; d_flags = nullptr; node_cap = 0;
        WCU_CHECK(cudaMalloc((void **) &d_nodes, (size_t) 2 * P * sizeof(Node)));
        WCU_CHECK(cudaMalloc((void **) &d_flags, (size_t) P * sizeof(int)));
        node_cap = (size_t) 2 * P;
        if (d_nrm_sorted) WCU_CHECK(cudaFree(d_nrm_sorted));
        d_nrm_sorted = nullptr;
    }
    if (nrm_n && !d_nrm_sorted) WCU_CHECK(cudaMalloc((void **) &d_nrm_sorted, (size_t) (node_cap / 2) * kLeaf * sizeof(float4)));
    int rc = cloud.sort(n_pad, nrm_n ? d_nrm_raw : nullptr, nrm_n ? d_nrm_sorted : nullptr);
    if (rc) return rc;
    WCU_CHECK(cudaMemsetAsync(d_flags, 0, (size_t) P * sizeof(int), cloud.stream));
    tree_kernel<<<(P + kBuildThreads - 1) / kBuildThreads, kBuildThreads, 0, cloud.stream>>>(cloud.d_sorted, P, d_nodes,
                                                                                              d_flags);
    ++cloud.launches;
    WCU_CHECK(cudaGetLastError());
    dirty = false;
    return WAVECU_OK;
}

void TargetIndex::release() {
    cloud.release();
    for (void *p : {(void *) d_nodes, (void *) d_flags, (void *) d_nrm_raw, (void *) d_nrm_sorted})
        if (p) cudaFree(p);
    d_nodes = nullptr; d_flags = nullptr; d_nrm_raw = d_nrm_sorted = nullptr;
    node_cap = nrm_cap = nrm_n = 0;
}

}  // namespace wavecu
