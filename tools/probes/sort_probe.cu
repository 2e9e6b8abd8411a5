// Scratch probe: cub radix sort of 1 M Morton keys - (u64 key, u32 value) pairs over 40 bits against packed u64 keys
// (key << 24 | index) sorted on bits [24, 64), and (u32 key, u32 value) pairs over 32 bits.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/sort_probe tools/probes/sort_probe.cu
#include <cub/device/device_radix_sort.cuh>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
__global__ void fill(unsigned long long *k, unsigned long long *pk, unsigned *k32, unsigned *v, size_t n) {
    size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i < n) {
        unsigned long long key = (i * 0x9E3779B97F4A7C15ull) >> 24;
        k[i] = key; pk[i] = (key << 24) | i; k32[i] = (unsigned) (key >> 8); v[i] = (unsigned) i;
    }
}
int main() {
    const size_t n = 1 << 20;
    unsigned long long *k0, *k1, *p0, *p1; unsigned *v0, *v1, *q0, *q1;
    CK(cudaMalloc(&k0, n * 8)); CK(cudaMalloc(&k1, n * 8)); CK(cudaMalloc(&p0, n * 8)); CK(cudaMalloc(&p1, n * 8));
    CK(cudaMalloc(&v0, n * 4)); CK(cudaMalloc(&v1, n * 4)); CK(cudaMalloc(&q0, n * 4)); CK(cudaMalloc(&q1, n * 4));
    size_t tb = 0, t2 = 0, t3 = 0;
    { cub::DoubleBuffer<unsigned long long> a(k0, k1); cub::DoubleBuffer<unsigned> b(v0, v1); CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, a, b, (int) n, 0, 40)); }
    { cub::DoubleBuffer<unsigned long long> a(p0, p1); CK(cub::DeviceRadixSort::SortKeys(nullptr, t2, a, (int) n, 24, 64)); }
    { cub::DoubleBuffer<unsigned> a(q0, q1); cub::DoubleBuffer<unsigned> b(v0, v1); CK(cub::DeviceRadixSort::SortPairs(nullptr, t3, a, b, (int) n, 0, 32)); }
    void *tmp; CK(cudaMalloc(&tmp, std::max(tb, std::max(t2, t3))));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int rep = 0; rep < 3; ++rep)
        for (int mode = 0; mode < 3; ++mode) {
            float tot = 0;
            for (int r = 0; r < 10; ++r) {
                fill<<<(unsigned) (n / 256), 256>>>(k0, p0, q0, v0, n);
                CK(cudaEventRecord(e0));
                if (mode == 0) { cub::DoubleBuffer<unsigned long long> a(k0, k1); cub::DoubleBuffer<unsigned> b(v0, v1); CK(cub::DeviceRadixSort::SortPairs(tmp, tb, a, b, (int) n, 0, 40)); }
                if (mode == 1) { cub::DoubleBuffer<unsigned long long> a(p0, p1); CK(cub::DeviceRadixSort::SortKeys(tmp, t2, a, (int) n, 24, 64)); }
                if (mode == 2) { cub::DoubleBuffer<unsigned> a(q0, q1); cub::DoubleBuffer<unsigned> b(v0, v1); CK(cub::DeviceRadixSort::SortPairs(tmp, t3, a, b, (int) n, 0, 32)); }
                CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); tot += ms;
            }
            const char *names[] = {"pairs u64+u32, 40 bits", "packed u64 keys, bits 24-64", "pairs u32+u32, 32 bits"};
            printf("%-30s %.1f us\n", names[mode], tot * 100);
        }
    return 0;
}
