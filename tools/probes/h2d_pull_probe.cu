// Scratch probe: how much does an in-flight host-to-device copy slow down a concurrently running radix sort, and
// does it matter whether the bytes come in through the copy engine (cudaMemcpyAsync from pinned memory) or are
// pulled by a few SMs reading the mapped pinned buffer directly?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/h2d_pull_probe tools/probes/h2d_pull_probe.cu
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void pull_kernel(const float4 *__restrict__ host, float4 *dev, size_t n) {
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(host + i));
        dev[i] = v;
    }
}
__global__ void fill_keys(unsigned long long *k, unsigned *v, size_t n) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i < n) { k[i] = (i * 0x9E3779B97F4A7C15ull) >> 24; v[i] = (unsigned) i; }
}

int main(int argc, char **argv) {
    const size_t n = 1 << 20, copy_bytes = 256u << 20;
    const int pull_blocks = argc > 1 ? atoi(argv[1]) : 16, pull_threads = argc > 2 ? atoi(argv[2]) : 512;
    float4 *h; CK(cudaHostAlloc((void **) &h, copy_bytes, cudaHostAllocDefault));
    memset(h, 1, copy_bytes);
    float4 *d; CK(cudaMalloc((void **) &d, copy_bytes));
    unsigned long long *k0, *k1; unsigned *v0, *v1;
    CK(cudaMalloc((void **) &k0, n * 8)); CK(cudaMalloc((void **) &k1, n * 8)); CK(cudaMalloc((void **) &v0, n * 4)); CK(cudaMalloc((void **) &v1, n * 4));
    size_t tmp_bytes = 0; cub::DoubleBuffer<unsigned long long> kb(k0, k1); cub::DoubleBuffer<unsigned> vb(v0, v1);
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, kb, vb, (int) n, 0, 40));
    void *tmp; CK(cudaMalloc(&tmp, tmp_bytes));
    cudaStream_t s_work, s_copy; CK(cudaStreamCreate(&s_work)); CK(cudaStreamCreate(&s_copy));
    cudaEvent_t e0, e1, c0, c1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&c0)); CK(cudaEventCreate(&c1));
    auto victim = [&]() {
        fill_keys<<<(unsigned) (n / 256), 256, 0, s_work>>>(k0, v0, n);
        CK(cudaEventRecord(e0, s_work));
        for (int r = 0; r < 4; ++r) {
            cub::DoubleBuffer<unsigned long long> a(k0, k1); cub::DoubleBuffer<unsigned> b(v0, v1);
            CK(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, a, b, (int) n, 0, 40, s_work));
        }
        CK(cudaEventRecord(e1, s_work));
    };
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 0; mode < 3; ++mode) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(c0, s_copy));
            if (mode == 1) CK(cudaMemcpyAsync(d, h, copy_bytes, cudaMemcpyHostToDevice, s_copy));
            if (mode == 2) pull_kernel<<<pull_blocks, pull_threads, 0, s_copy>>>(h, d, copy_bytes / 16);
            CK(cudaEventRecord(c1, s_copy));
            victim();
            CK(cudaDeviceSynchronize());
            float ms_v, ms_c; CK(cudaEventElapsedTime(&ms_v, e0, e1)); CK(cudaEventElapsedTime(&ms_c, c0, c1));
            printf("%-12s 4 sorts of 1M pairs: %.3f ms (%.1f us each)   copy: %.3f ms = %.1f GB/s\n",
                   mode == 0 ? "quiet" : mode == 1 ? "copy engine" : "SM pull", ms_v, ms_v * 250, ms_c, mode ? copy_bytes / ms_c * 1e-6 : 0.0);
        }
    // pull bandwidth alone for a few shapes
    for (int blocks : {4, 8, 16, 32, 64})
        for (int threads : {256, 1024}) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(c0, s_copy));
            pull_kernel<<<blocks, threads, 0, s_copy>>>(h, d, copy_bytes / 16);
            CK(cudaEventRecord(c1, s_copy));
            CK(cudaDeviceSynchronize());
            float ms_c; CK(cudaEventElapsedTime(&ms_c, c0, c1));
            printf("pull %3d x %4d: %.1f GB/s\n", blocks, threads, copy_bytes / ms_c * 1e-6);
        }
    return 0;
}
