#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <limits>

#include "../../include/wavecu.h"
#include "index.cuh"
#include "voxel.cuh"

namespace wavecu {

namespace {

__global__ void voxel_key_kernel(const float4 *__restrict__ in, size_t n, GridDesc g, unsigned *keys, unsigned *vals) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = in[i];
    unsigned key = 0xffffffffu;
    if (finite3(p.x, p.y, p.z)) {
        // ijk = (int)(floor(p * inv) - (float) min_b), voxel_grid.hpp
        const int i0 = (int) __fsub_rn(floorf(__fmul_rn(p.x, g.inv)), (float) g.min_b[0]);
        const int i1 = (int) __fsub_rn(floorf(__fmul_rn(p.y, g.inv)), (float) g.min_b[1]);
        const int i2 = (int) __fsub_rn(floorf(__fmul_rn(p.z, g.inv)), (float) g.min_b[2]);
        key = (unsigned) (i0 * g.mul[0] + i1 * g.mul[1] + i2 * g.mul[2]);
    }
    keys[i] = key;
    vals[i] = (unsigned) i;
}

__global__ void voxel_head_kernel(const unsigned *__restrict__ keys, size_t n, int *flags) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = keys[i];
    flags[i] = (k != 0xffffffffu && (i == 0 || keys[i - 1] != k)) ? 1 : 0;
}

// fp32 sums in ascending cloud index (the sort is stable), centroid = sum / n.  One thread per voxel for
// runs of up to kShortRun points (the usual case at the reference's leaf sizes); longer runs - thousands of
// points per voxel at the coarse multiscale levels of a dense scan - are handed to voxel_long_kernel,
// because a single thread chasing vals[] -> in[] point by point took 0.6 ms per cloud and level there.
constexpr int kShortRun = 32;

__global__ void voxel_centroid_kernel(const float4 *__restrict__ in, const unsigned *__restrict__ keys,
                                      const unsigned *__restrict__ vals, const int *__restrict__ pos, size_t n,
                                      float4 *out, int *long_count, int *long_list) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = keys[i];
    const bool head = (k != 0xffffffffu) && (i == 0 || keys[i - 1] != k);
    if (!head) return;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    size_t j = i;
    for (; j < n && keys[j] == k; ++j) {
        if (j - i == (size_t) kShortRun) {   // a long run: a warp takes it from the start
            long_list[atomicAdd(long_count, 1)] = (int) i;
            return;
        }
        const float4 p = in[vals[j]];
        sx = __fadd_rn(sx, p.x);
        sy = __fadd_rn(sy, p.y);
        sz = __fadd_rn(sz, p.z);
    }
    const float cnt = (float) (j - i);
    out[pos[i]] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 1.0f);
}

// one warp per long run: the lanes gather 4 x 32 points per trip, the sums stay one sequential chain of fp32
// adds in run order (lane values handed over by shuffles), so the result is bit-identical to the serial loop
__global__ void __launch_bounds__(256) voxel_long_kernel(const float4 *__restrict__ in, const unsigned *__restrict__ keys,
                                                         const unsigned *__restrict__ vals, const int *__restrict__ pos,
                                                         size_t n, float4 *out, const int *__restrict__ long_count,
                                                         const int *__restrict__ long_list) {
    const int total = *long_count;
    const int lane = threadIdx.x & 31;
    for (int w = (int) ((blockIdx.x * (size_t) blockDim.x + threadIdx.x) >> 5); w < total;
         w += (int) ((gridDim.x * (size_t) blockDim.x) >> 5)) {
        const size_t start = (size_t) long_list[w];
        const unsigned k = keys[start];
        float sx = 0.f, sy = 0.f, sz = 0.f;
        size_t count = 0;
        constexpr int kTrip = 4;
        bool more = true;
        for (size_t base = start; more; base += 32 * kTrip) {
            float4 p[kTrip];
            unsigned member[kTrip];
#pragma unroll
            for (int u = 0; u < kTrip; ++u) {
                const size_t j = base + (size_t) u * 32 + lane;
                const bool mine = j < n && keys[j] == k;
                member[u] = __ballot_sync(0xffffffffu, mine);
                p[u] = mine ? in[vals[j]] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kTrip; ++u) {
                const int cnt = __popc(member[u]);   // the run is contiguous: lanes 0 .. cnt-1
#pragma unroll
                for (int l = 0; l < 32; ++l) {
                    const float vx = __shfl_sync(0xffffffffu, p[u].x, l), vy = __shfl_sync(0xffffffffu, p[u].y, l),
                                vz = __shfl_sync(0xffffffffu, p[u].z, l);
                    if (l < cnt) {
                        sx = __fadd_rn(sx, vx);
                        sy = __fadd_rn(sy, vy);
                        sz = __fadd_rn(sz, vz);
                    }
                }
                count += (size_t) cnt;
                if (cnt < 32) more = false;
            }
        }
        if (lane == 0) {
            const float cnt = (float) count;
            out[pos[start]] = make_float4(__fdiv_rn(sx, cnt), __fdiv_rn(sy, cnt), __fdiv_rn(sz, cnt), 1.0f);
        }
    }
}

// pos is the exclusive scan of the head flags: total = pos[n-1] + head(n-1)
__global__ void voxel_count_kernel(const unsigned *__restrict__ keys, const int *__restrict__ pos, size_t n, int *total) {
    if (threadIdx.x != 0) return;
    const unsigned k = keys[n - 1];
    const bool head = (k != 0xffffffffu) && (n == 1 || keys[n - 2] != k);
    *total = pos[n - 1] + (head ? 1 : 0);
}

__global__ void affine3d_kernel(float4 *cloud, size_t n, double m00, double m01, double m02, double m03, double m10,
                                double m11, double m12, double m13, double m20, double m21, double m22, double m23) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = cloud[i];
    const double x = p.x, y = p.y, z = p.z;
    const float ox = (float) (((m00 * x + m01 * y) + m02 * z) + m03);
    const float oy = (float) (((m10 * x + m11 * y) + m12 * z) + m13);
    const float oz = (float) (((m20 * x + m21 * y) + m22 * z) + m23);
    cloud[i] = make_float4(ox, oy, oz, p.w);
}

}  // namespace

int affine3d_inplace(float4 *d_cloud, size_t n, const double T[12], cudaStream_t stream) {
    if (!n) return WAVECU_OK;
    affine3d_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(d_cloud, n, T[0], T[1], T[2], T[3], T[4], T[5],
                                                                       T[6], T[7], T[8], T[9], T[10], T[11]);
    WCU_CHECK(cudaGetLastError());
    return WAVECU_OK;
}

int VoxelWork::reserve(size_t n) {
    WCU_CHECK(cudaSetDevice(device));
    if (!d_bbox) {
        WCU_CHECK(cudaMalloc((void **) &d_bbox, 8 * sizeof(unsigned)));
        WCU_CHECK(cudaHostAlloc((void **) &h_bbox, 8 * sizeof(unsigned), cudaHostAllocDefault));
        WCU_CHECK(cudaHostAlloc((void **) &h_count, sizeof(int), cudaHostAllocDefault));
    }
    if (n <= cap) return WAVECU_OK;
    for (void *p : {(void *) d_keys, (void *) d_keys_alt, (void *) d_vals, (void *) d_vals_alt, (void *) d_pos, d_tmp})
        if (p) WCU_CHECK(cudaFree(p));
    d_keys = d_keys_alt = d_vals = d_vals_alt = nullptr;
    d_pos = nullptr;
    d_tmp = nullptr;
    cap = 0;
    const size_t a = n + n / 8 + 64;
    WCU_CHECK(cudaMalloc((void **) &d_keys, a * sizeof(unsigned)));
    WCU_CHECK(cudaMalloc((void **) &d_keys_alt, a * sizeof(unsigned)));
    WCU_CHECK(cudaMalloc((void **) &d_vals, a * sizeof(unsigned)));
    WCU_CHECK(cudaMalloc((void **) &d_vals_alt, a * sizeof(unsigned)));
    WCU_CHECK(cudaMalloc((void **) &d_pos, (a + 1) * sizeof(int)));
    size_t need_sort = 0, need_scan = 0;
    cub::DoubleBuffer<unsigned> kb(d_keys, d_keys_alt), vb(d_vals, d_vals_alt);
    WCU_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, need_sort, kb, vb, (int) a, 0, 32, stream));
    WCU_CHECK(cub::DeviceScan::ExclusiveSum(nullptr, need_scan, d_pos, d_pos, (int) a, stream));
    tmp_bytes = std::max(need_sort, need_scan);
    WCU_CHECK(cudaMalloc(&d_tmp, tmp_bytes));
    cap = a;
    return WAVECU_OK;
}

int VoxelWork::prepare(const float4 *d_in, size_t n, float leaf, int *status) {
    *status = -1;
    n_voxels = 0;
    n_sorted = n;
    if (n == 0) return WAVECU_OK;
    int rc = reserve(n);
    if (rc) return rc;
    // getMinMax3D
    launch_bbox(d_in, n, d_bbox, stream);
    launches += 2;
    WCU_CHECK(cudaMemcpyAsync(h_bbox, d_bbox, 8 * sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    WCU_CHECK(cudaStreamSynchronize(stream));
    auto decode = [](unsigned u) {
        const unsigned b = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
        float f;
        memcpy(&f, &b, 4);
        return f;
    };
    if (h_bbox[6] == 0) return WAVECU_OK;  // no finite point
    float mn[3], mx[3];
    for (int d = 0; d < 3; ++d) {
        mn[d] = decode(h_bbox[d]);
        mx[d] = decode(h_bbox[3 + d]);
    }
    const float inv = 1.0f / leaf;
    const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
    const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
    const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
    if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
        *status = 0;  // "Leaf size is too small for the input dataset. Integer indices would overflow."
        return WAVECU_OK;
    }
    GridDesc g;
    g.inv = inv;
    for (int d = 0; d < 3; ++d) {
        g.min_b[d] = static_cast<int>(std::floor(mn[d] * inv));
        g.div_b[d] = static_cast<int>(std::floor(mx[d] * inv)) - g.min_b[d] + 1;
    }
    g.mul[0] = 1;
    g.mul[1] = g.div_b[0];
    g.mul[2] = g.div_b[0] * g.div_b[1];
    grid = g;
    const unsigned blocks = (unsigned) ((n + 255) / 256);
    voxel_key_kernel<<<blocks, 256, 0, stream>>>(d_in, n, g, d_keys, d_vals);
    cub::DoubleBuffer<unsigned> kb(d_keys, d_keys_alt), vb(d_vals, d_vals_alt);
    size_t need = tmp_bytes;
    WCU_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, need, kb, vb, (int) n, 0, 32, stream));
    if (kb.Current() != d_keys) std::swap(d_keys, d_keys_alt);
    if (vb.Current() != d_vals) std::swap(d_vals, d_vals_alt);
    voxel_head_kernel<<<blocks, 256, 0, stream>>>(d_keys, n, d_pos);
    need = tmp_bytes;
    WCU_CHECK(cub::DeviceScan::ExclusiveSum(d_tmp, need, d_pos, d_pos, (int) n, stream));
    voxel_count_kernel<<<1, 32, 0, stream>>>(d_keys, d_pos, n, d_pos + n);
    launches += 3 + 6 + 2;
    WCU_CHECK(cudaMemcpyAsync(h_count, d_pos + n, sizeof(int), cudaMemcpyDeviceToHost, stream));
    WCU_CHECK(cudaStreamSynchronize(stream));
    WCU_CHECK(cudaGetLastError());
    n_voxels = *h_count;
    *status = 1;
    return WAVECU_OK;
}

int VoxelWork::filter(const float4 *d_in, size_t n, float leaf, float4 *d_out, size_t *n_out, int *filtered) {
    *n_out = 0;
    if (filtered) *filtered = 1;
    if (n == 0) return WAVECU_OK;
    int status = 0;
    int rc = prepare(d_in, n, leaf, &status);
    if (rc) return rc;
    if (status < 0) return WAVECU_OK;  // no finite point: empty output
    if (status == 0) {                 // overflow rule: output = input
        if (d_out != d_in) WCU_CHECK(cudaMemcpyAsync(d_out, d_in, n * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        *n_out = n;
        if (filtered) *filtered = 0;
        return WAVECU_OK;
    }
    // long-run list: worst case one run per kShortRun + 1 points
    const size_t list_need = n / (kShortRun + 1) + 2;
    if (list_need > long_cap) {
        if (d_long) WCU_CHECK(cudaFree(d_long));
        d_long = nullptr;
        WCU_CHECK(cudaMalloc((void **) &d_long, (list_need + 64 + 1) * sizeof(int)));
        long_cap = list_need + 64;
    }
    WCU_CHECK(cudaMemsetAsync(d_long, 0, sizeof(int), stream));   // d_long[0] = count, list behind it
    voxel_centroid_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, stream>>>(d_in, d_keys, d_vals, d_pos, n, d_out, d_long,
                                                                            d_long + 1);
    voxel_long_kernel<<<148 * 4, 256, 0, stream>>>(d_in, d_keys, d_vals, d_pos, n, d_out, d_long, d_long + 1);
    launches += 2;
    WCU_CHECK(cudaGetLastError());
    *n_out = (size_t) n_voxels;
    return WAVECU_OK;
}

void VoxelWork::release() {
    if (d_long) cudaFree(d_long);
    d_long = nullptr;
    long_cap = 0;
    for (void *p : {(void *) d_keys, (void *) d_keys_alt, (void *) d_vals, (void *) d_vals_alt, (void *) d_pos,
                    (void *) d_bbox, d_tmp})
        if (p) cudaFree(p);
    if (h_bbox) cudaFreeHost(h_bbox);
    if (h_count) cudaFreeHost(h_count);
    d_keys = d_keys_alt = d_vals = d_vals_alt = nullptr;
    d_pos = nullptr;
    d_bbox = nullptr;
    d_tmp = nullptr;
    h_bbox = nullptr;
    h_count = nullptr;
    cap = tmp_bytes = 0;
}

}  // namespace wavecu

using namespace wavecu;

extern "C" int wavecu_voxel_grid(int device, const float *xyzw, size_t n, float leaf, float *out_xyzw, size_t *n_out,
                                 int *filtered) {
    if (!n_out || (n && (!xyzw || !out_xyzw)) || !(leaf > 0)) return WAVECU_ERR_ARG;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_last_error("no such CUDA device (libwavecu has no CPU fallback)");
        return WAVECU_ERR_CUDA;
    }
    *n_out = 0;
    if (n == 0) return WAVECU_OK;
    WCU_CHECK(cudaSetDevice(device));
    VoxelWork w;
    w.device = device;
    WCU_CHECK(cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking));
    float4 *d_in = nullptr, *d_out = nullptr;
    int rc = WAVECU_OK;
    do {
        if (cudaMalloc((void **) &d_in, n * sizeof(float4)) != cudaSuccess ||
            cudaMalloc((void **) &d_out, n * sizeof(float4)) != cudaSuccess) {
            set_last_error("cudaMalloc failed in wavecu_voxel_grid");
            rc = WAVECU_ERR_CUDA;
            break;
        }
        cudaMemcpyAsync(d_in, xyzw, n * sizeof(float4), cudaMemcpyHostToDevice, w.stream);
        rc = w.filter(d_in, n, leaf, d_out, n_out, filtered);
        if (rc) break;
        cudaMemcpyAsync(out_xyzw, d_out, *n_out * sizeof(float4), cudaMemcpyDeviceToHost, w.stream);
        if (cudaStreamSynchronize(w.stream) != cudaSuccess) {
            set_last_error("stream synchronize failed in wavecu_voxel_grid");
            rc = WAVECU_ERR_CUDA;
        }
    } while (0);
    if (d_in) cudaFree(d_in);
    if (d_out) cudaFree(d_out);
    w.release();
    cudaStreamDestroy(w.stream);
    return rc;
}
