// Stand-alone exact 1-NN (wavecu_nn_*): the correspondence kernel without the estimator.
// Replaces pcl::KdTreeFLANN::setInputCloud + nearestKSearch(k = 1) as estimateLUMold uses them
// (wave_matching/src/icp_pcl_functions.cpp:67-80); also the kernel the roofline figure is taken on.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "../../include/wavecu.h"
#include "index.cuh"

namespace wavecu {

namespace {
// per host thread: MultiMatcher workers each drive their own handle, and a failure in one of them
// must not be reported with (or overwritten by) another thread's message
thread_local std::string g_last_error;
}  // namespace

void set_last_error(const std::string &msg) {
    g_last_error = msg;
}

constexpr int kNnThreads = 128;

// queries in Morton order (w = original index); results written in original query order
__global__ void __launch_bounds__(kNnThreads) nn_kernel(const float4 *__restrict__ q_sorted, int nq,
                                                         NnIndex ix, float thr, int *out_idx,
                                                         float *out_d2) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nq) return;
    const float4 q = q_sorted[s];
    const int orig = __float_as_int(q.w);
    if (orig == 0x7fffffff) return;  // pad (non-finite query): left at -1 / inf by the caller's fill
    float best = thr;
    int best_idx = 0x7fffffff, best_pos = -1;
    nn_search_cells(q.x, q.y, q.z, ix, best, best_idx, best_pos);
    out_idx[orig] = best_pos >= 0 ? best_idx : -1;
    out_d2[orig] = best_pos >= 0 ? best : INFINITY;
}

__global__ void nn_fill_kernel(int *idx, float *d2, size_t n) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i < n) {
        idx[i] = -1;
        d2[i] = INFINITY;
    }
}

struct NnHandle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    TargetIndex tgt;
    MortonCloud q;
    int *d_idx = nullptr;
    float *d_d2 = nullptr;
    size_t out_cap = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;

    static float threshold(double max_dist) {
        if (!(max_dist > 0)) return FLT_MAX;
        const double max2 = max_dist * max_dist;
        float f = (max2 >= (double) FLT_MAX) ? FLT_MAX : (float) max2;
        if ((double) f > max2) f = std::nextafterf(f, -INFINITY);
        return f;
    }

    int search(const float *queries, bool q_on_device, size_t nq, double max_dist, int *idx, float *d2,
               bool out_on_device, int repeats, float *elapsed_ms) {
        WCU_CHECK(cudaSetDevice(device));
        if (tgt.dirty) {
            const int rc = tgt.build();
            if (rc) return rc;
        }
        int rc = q.upload(queries, nq, q_on_device);
        if (rc) return rc;
        int *o_idx = idx;
        float *o_d2 = d2;
        if (!out_on_device) {
            if (nq > out_cap) {
                if (d_idx) WCU_CHECK(cudaFree(d_idx));
                if (d_d2) WCU_CHECK(cudaFree(d_d2));
                d_idx = nullptr; d_d2 = nullptr; out_cap = 0;
                WCU_CHECK(cudaMalloc((void **) &d_idx, (nq + 64) * sizeof(int)));
                WCU_CHECK(cudaMalloc((void **) &d_d2, (nq + 64) * sizeof(float)));
                out_cap = nq + 64;
            }
            o_idx = d_idx;
            o_d2 = d_d2;
        }
        if (nq == 0) return WAVECU_OK;
        rc = q.sort(nq);
        if (rc) return rc;
        nn_fill_kernel<<<(unsigned) ((nq + 255) / 256), 256, 0, stream>>>(o_idx, o_d2, nq);
        const float thr = threshold(max_dist);
        repeats = std::max(repeats, 1);
        if (elapsed_ms) WCU_CHECK(cudaEventRecord(e0, stream));
        for (int r = 0; r < repeats; ++r)
            nn_kernel<<<(unsigned) ((nq + kNnThreads - 1) / kNnThreads), kNnThreads, 0, stream>>>(
                q.d_sorted, (int) nq, tgt.index(), thr, o_idx, o_d2);
        if (elapsed_ms) WCU_CHECK(cudaEventRecord(e1, stream));
        WCU_CHECK(cudaGetLastError());
        if (!out_on_device) {
            WCU_CHECK(cudaMemcpyAsync(idx, d_idx, nq * sizeof(int), cudaMemcpyDeviceToHost, stream));
            WCU_CHECK(cudaMemcpyAsync(d2, d_d2, nq * sizeof(float), cudaMemcpyDeviceToHost, stream));
        }
        WCU_CHECK(cudaStreamSynchronize(stream));
        if (elapsed_ms) WCU_CHECK(cudaEventElapsedTime(elapsed_ms, e0, e1));
        return WAVECU_OK;
    }
};

}  // namespace wavecu

using namespace wavecu;

struct wavecu_nn {
    NnHandle h;
};

extern "C" {

const char *wavecu_last_error(void) {
    return g_last_error.c_str();  // the calling thread's own last message
}

int wavecu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int wavecu_nn_create(int device, void *stream, wavecu_nn **out) {
    if (!out) return WAVECU_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_last_error("no such CUDA device (libwavecu has no CPU fallback)");
        return WAVECU_ERR_CUDA;
    }
    wavecu_nn *w = new wavecu_nn();
    NnHandle &h = w->h;
    h.device = device;
    h.stream = (cudaStream_t) stream;
    WCU_CHECK(cudaSetDevice(device));
    if (!h.stream) {
        WCU_CHECK(cudaStreamCreateWithFlags(&h.stream, cudaStreamNonBlocking));
        h.own_stream = true;
    }
    h.q.key_bits = 12;
    h.tgt.cloud.device = h.q.device = device;
    h.tgt.cloud.stream = h.q.stream = h.stream;
    WCU_CHECK(cudaEventCreate(&h.e0));
    WCU_CHECK(cudaEventCreate(&h.e1));
    *out = w;
    return WAVECU_OK;
}

int wavecu_nn_destroy(wavecu_nn *w) {
    if (!w) return WAVECU_OK;
    NnHandle &h = w->h;
    cudaSetDevice(h.device);
    h.tgt.release();
    h.q.release();
    if (h.d_idx) cudaFree(h.d_idx);
    if (h.d_d2) cudaFree(h.d_d2);
    if (h.e0) cudaEventDestroy(h.e0);
    if (h.e1) cudaEventDestroy(h.e1);
    if (h.own_stream && h.stream) cudaStreamDestroy(h.stream);
    delete w;
    return WAVECU_OK;
}

int wavecu_nn_set_target(wavecu_nn *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    return w->h.tgt.set_points(xyzw, n, false);
}

int wavecu_nn_search(wavecu_nn *w, const float *q_xyzw, size_t nq, double max_dist, int *idx, float *dist2) {
    if (!w || (nq && (!q_xyzw || !idx || !dist2))) return WAVECU_ERR_ARG;
    return w->h.search(q_xyzw, false, nq, max_dist, idx, dist2, false, 1, nullptr);
}

int wavecu_nn_search_device(wavecu_nn *w, const void *d_q_xyzw, size_t nq, double max_dist, void *d_idx,
                            void *d_dist2, int repeats, float *elapsed_ms) {
    if (!w || (nq && (!d_q_xyzw || !d_idx || !d_dist2))) return WAVECU_ERR_ARG;
    return w->h.search((const float *) d_q_xyzw, true, nq, max_dist, (int *) d_idx, (float *) d_dist2, true, repeats,
                       elapsed_ms);
}

}  // extern "C"
