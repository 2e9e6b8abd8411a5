// Scratch probe: host-to-device copy bandwidth from ordinary pinned memory against write-combined pinned memory, one
// copy at a time and two copies on two streams.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/h2d_wc_probe tools/probes/h2d_wc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
int main() {
    const size_t bytes = 256u << 20;
    void *h0, *h1, *d0, *d1;
    CK(cudaHostAlloc(&h0, bytes, cudaHostAllocDefault));
    CK(cudaHostAlloc(&h1, bytes, cudaHostAllocWriteCombined));
    memset(h0, 1, bytes); memset(h1, 1, bytes);
    CK(cudaMalloc(&d0, bytes)); CK(cudaMalloc(&d1, bytes));
    cudaStream_t s0, s1; CK(cudaStreamCreate(&s0)); CK(cudaStreamCreate(&s1));
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 3; ++rep)
        for (int mode = 0; mode < 4; ++mode) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(a, s0));
            if (mode == 0) CK(cudaMemcpyAsync(d0, h0, bytes, cudaMemcpyHostToDevice, s0));
            if (mode == 1) CK(cudaMemcpyAsync(d0, h1, bytes, cudaMemcpyHostToDevice, s0));
            if (mode == 2) for (int c = 0; c < 16; ++c) CK(cudaMemcpyAsync((char *) d0 + c * (bytes / 16), (char *) h0 + c * (bytes / 16), bytes / 16, cudaMemcpyHostToDevice, s0));
            if (mode == 3) {  // two streams, half each
                CK(cudaMemcpyAsync(d1, h1, bytes / 2, cudaMemcpyHostToDevice, s1));
                CK(cudaMemcpyAsync(d0, h0, bytes / 2, cudaMemcpyHostToDevice, s0));
                CK(cudaStreamSynchronize(s1));
            }
            CK(cudaEventRecord(b, s0));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            const char *names[] = {"pinned", "write-combined", "pinned, 16 MB pieces", "two streams"};
            printf("%-22s %.3f ms = %.1f GB/s\n", names[mode], ms, bytes / ms * 1e-6);
        }
    return 0;
}
