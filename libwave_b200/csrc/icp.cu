// Host side of the ICP handle behind include/wavecu.h (wavecu_icp_*): device buffers, the
// per-align() launch sequence and the C ABI.  Reference behaviour restated: pcl::Registration::
// align + IterativeClosestPoint::computeTransformation as driven by wave_matching/src/icp.cpp
// :47-50 (parameters) and :75-133 (ICPMatcher::match).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/wavecu.h"
#include "icp_kernels.cuh"
#include "index.cuh"

namespace wavecu {

namespace {

struct SetupArgs {
    const unsigned *src_bbox, *tgt_bbox;
    MatchConsts *mc;
    IcpState *st;
    Acc128 *acc;
    double max_corr;
};

__device__ double bbox_max_abs(const unsigned *bb) {
    double m = 0.0;
    for (int d = 0; d < 3; ++d) {
        const float lo = ordered_to_float(bb[d]), hi = ordered_to_float(bb[3 + d]);
        if (lo <= hi) m = fmax(m, fmax(fabs((double) lo), fabs((double) hi)));
    }
    return m;
}

// Fixed-point exponents (DESIGN.md "Estimator arithmetic"), the fp32 distance threshold, and a
// fresh iteration state - all on the device so that align() needs no host round trip to start.
__global__ void setup_kernel(SetupArgs a) {
    const int t = threadIdx.x;
    for (int i = t; i < kAccSlots * kMaxAcc; i += blockDim.x) {
        a.acc[i].lo = 0;
        a.acc[i].hi = 0;
    }
    if (t != 0) return;
    MatchConsts &mc = *a.mc;
    for (int d = 0; d < 3; ++d) {
        mc.src_lo[d] = ordered_to_float(a.src_bbox[d]);
        mc.src_hi[d] = ordered_to_float(a.src_bbox[3 + d]);
        mc.tgt_lo[d] = ordered_to_float(a.tgt_bbox[d]);
        mc.tgt_hi[d] = ordered_to_float(a.tgt_bbox[3 + d]);
    }
    const double ms = bbox_max_abs(a.src_bbox), mt = bbox_max_abs(a.tgt_bbox);
    const double reach = fmin(a.max_corr, 2.0 * ms + mt);
    const double M = fmax(mt + reach, 1.0);
    int E, Ed;
    frexp(M, &E);
    const double max2 = a.max_corr * a.max_corr;
    const double dmax = fmax(fmin(max2, 12.0 * M * M), 1e-30);
    frexp(dmax, &Ed);
    mc.k_lin = 50 - E;
    mc.k_quad = 50 - 2 * E;
    mc.k_d2 = 50 - Ed;
    mc.s_lin = ldexp(1.0, mc.k_lin);
    mc.s_quad = ldexp(1.0, mc.k_quad);
    mc.s_d2 = ldexp(1.0, mc.k_d2);
    mc.s_plane = ldexp(1.0, mc.k_quad - 3);
    // largest fp32 value f with (double) f <= max_corr^2 (correspondence_estimation.hpp compares
    // the fp32 distance against the double threshold and keeps equality)
    float f = (max2 >= (double) FLT_MAX) ? FLT_MAX : (float) max2;
    if ((double) f > max2) f = nextafterf(f, -INFINITY);
    mc.thr = f;

    IcpState &st = *a.st;
    for (int i = 0; i < 16; ++i) st.T_inc[i] = st.T_final[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    st.prev_mse = DBL_MAX;
    st.iter = 0;
    st.done = 0;
    st.converged = 0;
    st.state = WAVECU_CONV_NOT_CONVERGED;
    st.n_corr = 0;
}

__global__ void fill_int_kernel(int *p, int v, size_t n) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// out[i] = T (x) raw[i] in fp32 (pcl transformCloud with the final transform)
__global__ void apply_final_kernel(const float4 *__restrict__ raw, size_t n, const IcpState *st, float4 *out) {
    const size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
    if (i >= n) return;
    float T[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) T[k] = st->T_final[k];
    float4 p = raw[i];
    if (finite3(p.x, p.y, p.z)) {
        const float x = xform_row(T + 0, p.x, p.y, p.z), y = xform_row(T + 4, p.x, p.y, p.z),
                    z = xform_row(T + 8, p.x, p.y, p.z);
        p = make_float4(x, y, z, p.w);
    }
    out[i] = p;
}

// Morton order -> original source order
__global__ void unsort_corr_kernel(const float4 *__restrict__ cur, const int *__restrict__ nn_idx,
                                   const float *__restrict__ nn_d2, int n_sorted, int *out_idx, float *out_d2) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sorted) return;
    const float4 c = cur[s];
    const int orig = __float_as_int(c.w);
    if (orig == 0x7fffffff) return;  // pad
    out_idx[orig] = nn_idx[s];
    out_d2[orig] = nn_d2[s];
}

}  // namespace

struct IcpHandle {
    wavecu_icp_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    MortonCloud src;
    TargetIndex tgt;
    bool src_dirty = true;

    // iteration buffers
    int *d_nn_pos = nullptr, *d_nn_idx = nullptr;
    float *d_nn_d2 = nullptr;
    int *d_out_idx = nullptr;
    float *d_out_d2 = nullptr;
    float4 *d_aligned = nullptr;
    size_t iter_cap = 0;
    MatchConsts *d_mc = nullptr;
    IcpState *d_st = nullptr;
    Acc128 *d_acc = nullptr;
    TraceRow *d_trace = nullptr;
    int trace_cap = 0;

    // pinned host mirrors
    int *h_done = nullptr;  // ring of kDepth flags
    IcpState *h_st = nullptr;
    static constexpr int kDepth = 3;
    cudaEvent_t ev_ring[kDepth] = {nullptr, nullptr, nullptr};

    // results of the last align
    bool have_result = false;
    size_t result_n_src = 0;
    IcpState last;
    std::vector<TraceRow> trace;

    // profiling
    bool profiling = false;
    wavecu_stats stats{};
    std::vector<cudaEvent_t> ev_pool;

    int init();
    int ensure_iter_buffers(size_t n_src_pad, int max_iter);
    int align(double *T_out, int *converged, int *iterations, int *state);
    void release();
};

int IcpHandle::init() {
    WCU_CHECK(cudaSetDevice(device));
    if (!stream) {
        WCU_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }
    src.key_bits = 10;  // the source order only has to be spatially coherent: 31 sorted bits, 4 passes
    if (const char *e = getenv("WAVECU_SRC_BITS")) src.key_bits = std::max(1, std::min(21, atoi(e)));  // tuning knob
    if (const char *e = getenv("WAVECU_TGT_BITS")) tgt.cloud.key_bits = std::max(1, std::min(21, atoi(e)));
    src.device = tgt.cloud.device = device;
    src.stream = tgt.cloud.stream = stream;
    cudaFuncSetAttribute(correspond_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1);
    WCU_CHECK(cudaMalloc((void **) &d_mc, sizeof(MatchConsts)));
    WCU_CHECK(cudaMalloc((void **) &d_st, sizeof(IcpState)));
    WCU_CHECK(cudaMalloc((void **) &d_acc, sizeof(Acc128) * kAccSlots * kMaxAcc));
    WCU_CHECK(cudaHostAlloc((void **) &h_done, sizeof(int) * kDepth, cudaHostAllocDefault));
    WCU_CHECK(cudaHostAlloc((void **) &h_st, sizeof(IcpState), cudaHostAllocDefault));
    for (int i = 0; i < kDepth; ++i) WCU_CHECK(cudaEventCreateWithFlags(&ev_ring[i], cudaEventDisableTiming));
    return WAVECU_OK;
}

int IcpHandle::ensure_iter_buffers(size_t n_src_pad, int max_iter) {
    if (n_src_pad > iter_cap) {
        for (void *p : {(void *) d_nn_pos, (void *) d_nn_idx, (void *) d_nn_d2, (void *) d_out_idx, (void *) d_out_d2,
                        (void *) d_aligned})
            if (p) WCU_CHECK(cudaFree(p));
        d_nn_pos = d_nn_idx = d_out_idx = nullptr;
        d_nn_d2 = d_out_d2 = nullptr;
        d_aligned = nullptr;
        iter_cap = 0;
        const size_t a = n_src_pad + n_src_pad / 8 + 64;
        WCU_CHECK(cudaMalloc((void **) &d_nn_pos, a * sizeof(int)));
        WCU_CHECK(cudaMalloc((void **) &d_nn_idx, a * sizeof(int)));
        WCU_CHECK(cudaMalloc((void **) &d_nn_d2, a * sizeof(float)));
        WCU_CHECK(cudaMalloc((void **) &d_out_idx, a * sizeof(int)));
        WCU_CHECK(cudaMalloc((void **) &d_out_d2, a * sizeof(float)));
        WCU_CHECK(cudaMalloc((void **) &d_aligned, a * sizeof(float4)));
        iter_cap = a;
    }
    if (max_iter > trace_cap) {
        if (d_trace) WCU_CHECK(cudaFree(d_trace));
        d_trace = nullptr;
        WCU_CHECK(cudaMalloc((void **) &d_trace, sizeof(TraceRow) * (size_t) max_iter));
        trace_cap = max_iter;
    }
    return WAVECU_OK;
}

int IcpHandle::align(double *T_out, int *converged, int *iterations, int *state) {
    WCU_CHECK(cudaSetDevice(device));
    have_result = false;
    const size_t n_src = src.n, n_tgt = tgt.cloud.n;
    const int max_iter = std::max(prm.max_iter, 1);
    if (prm.estimator == WAVECU_EST_POINT_TO_PLANE && tgt.nrm_n != n_tgt) {
        set_last_error("point-to-plane estimator needs target normals (wavecu_icp_set_target_normals)");
        return WAVECU_ERR_STATE;
    }
    stats = wavecu_stats{};
    const long long launches0 = src.launches + tgt.cloud.launches;
    cudaEvent_t e_begin = nullptr, e_built = nullptr, e_end = nullptr;
    size_t ev_used = 0;
    auto next_event = [&]() -> cudaEvent_t {
        if (ev_used == ev_pool.size()) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            ev_pool.push_back(e);
        }
        return ev_pool[ev_used++];
    };
    if (profiling) {
        e_begin = next_event();
        e_built = next_event();
        e_end = next_event();
        WCU_CHECK(cudaEventRecord(e_begin, stream));
    }

    // ---- build: Morton-sort both clouds, AABB tree over the target ----
    const size_t n_src_pad = n_src;
    int rc = ensure_iter_buffers(std::max<size_t>(n_src_pad, 1), max_iter);
    if (rc) return rc;
    if (tgt.dirty) {
        rc = tgt.build();
        if (rc) return rc;
    }
    // the working cloud is consumed by the iterations, so the source is re-sorted for every align
    rc = src.sort(std::max<size_t>(n_src_pad, 1));
    if (rc) return rc;
    src_dirty = false;
    long long extra_launches = 0;
    if (n_src) {
        fill_int_kernel<<<(unsigned) ((n_src + 255) / 256), 256, 0, stream>>>(d_nn_pos, -1, n_src);
        ++extra_launches;
    }
    SetupArgs sa{src.d_bbox, tgt.cloud.d_bbox, d_mc, d_st, d_acc, prm.max_corr};
    setup_kernel<<<1, 256, 0, stream>>>(sa);
    ++extra_launches;
    WCU_CHECK(cudaGetLastError());
    if (profiling) WCU_CHECK(cudaEventRecord(e_built, stream));

    // ---- iterate: fused correspondence+reduction kernel, then the one-thread estimator ----
    IterArgs ia;
    ia.cur = src.d_sorted;
    ia.n_src = (int) n_src;
    ia.ix = tgt.index();
    ia.tgt = tgt.cloud.d_sorted;
    ia.nrm = tgt.d_nrm_sorted;
    ia.nn_pos = d_nn_pos;
    ia.nn_idx = d_nn_idx;
    ia.nn_d2 = d_nn_d2;
    ia.st = d_st;
    ia.mc = d_mc;
    ia.acc = d_acc;
    SolveArgs so{d_st, d_mc, d_acc, d_trace, max_iter, prm.t_eps, prm.fit_eps};
    const unsigned grid_nn = (unsigned) std::max<size_t>(1, (n_src + kIterThreads - 1) / kIterThreads);
    const unsigned grid_red = (unsigned) std::max<size_t>(
        1, (n_src + kReduceThreads * kReducePerThread - 1) / (kReduceThreads * kReducePerThread));
    std::vector<cudaEvent_t> it_ev;
    int launched = 0;
    bool finished = (n_src == 0 || n_tgt == 0);  // initCompute fails -> converged_ = false
    for (int k = 0; !finished && k < max_iter; ++k) {
        if (profiling) {
            it_ev.push_back(next_event());
            WCU_CHECK(cudaEventRecord(it_ev.back(), stream));
        }
        correspond_kernel<<<grid_nn, kIterThreads, 0, stream>>>(ia);
        if (profiling) {
            it_ev.push_back(next_event());
            WCU_CHECK(cudaEventRecord(it_ev.back(), stream));
        }
        if (prm.estimator == WAVECU_EST_POINT_TO_PLANE) {
            reduce_kernel<WAVECU_EST_POINT_TO_PLANE><<<grid_red, kReduceThreads, 0, stream>>>(ia);
            solve_kernel<WAVECU_EST_POINT_TO_PLANE><<<1, 64, 0, stream>>>(so);
        } else {
            reduce_kernel<WAVECU_EST_SVD><<<grid_red, kReduceThreads, 0, stream>>>(ia);
            solve_kernel<WAVECU_EST_SVD><<<1, 64, 0, stream>>>(so);
        }
        if (profiling) {
            it_ev.push_back(next_event());
            WCU_CHECK(cudaEventRecord(it_ev.back(), stream));
        }
        ++launched;
        const int slot = k % kDepth;
        WCU_CHECK(cudaMemcpyAsync(&h_done[slot], &d_st->done, sizeof(int), cudaMemcpyDeviceToHost, stream));
        WCU_CHECK(cudaEventRecord(ev_ring[slot], stream));
        // look at the flag of iteration k-(kDepth-1): the device always has work queued behind it
        if (k >= kDepth - 1) {
            const int old = (k - (kDepth - 1)) % kDepth;
            WCU_CHECK(cudaEventSynchronize(ev_ring[old]));
            if (h_done[old]) finished = true;
        }
    }
    if (profiling) WCU_CHECK(cudaEventRecord(e_end, stream));
    WCU_CHECK(cudaMemcpyAsync(h_st, d_st, sizeof(IcpState), cudaMemcpyDeviceToHost, stream));
    WCU_CHECK(cudaStreamSynchronize(stream));
    WCU_CHECK(cudaGetLastError());
    last = *h_st;
    if (n_src == 0 || n_tgt == 0) {
        last.converged = 0;
        last.state = WAVECU_CONV_NOT_CONVERGED;
    }
    trace.resize((size_t) std::max(0, last.iter));
    if (last.iter > 0)
        WCU_CHECK(cudaMemcpy(trace.data(), d_trace, sizeof(TraceRow) * (size_t) last.iter, cudaMemcpyDeviceToHost));
    have_result = true;
    result_n_src = n_src;

    stats.iterate_launches = last.iter;
    stats.kernel_launches = (src.launches + tgt.cloud.launches - launches0) + extra_launches + 3LL * launched;
    stats.pairs = (long long) last.iter * (long long) n_src;
    if (profiling) {
        float ms = 0;
        cudaEventElapsedTime(&ms, e_begin, e_built);
        stats.build_ms = ms;
        cudaEventElapsedTime(&ms, e_begin, e_end);
        stats.total_ms = ms;
        // only iterations that did work count (speculative launches past `done` return at once)
        for (int k = 0; k < last.iter + (last.state == WAVECU_CONV_NO_CORRESPONDENCES ? 1 : 0) &&
                        (size_t) (3 * k + 2) < it_ev.size(); ++k) {
            cudaEventElapsedTime(&ms, it_ev[3 * k], it_ev[3 * k + 1]);
            stats.iterate_ms += ms;
            cudaEventElapsedTime(&ms, it_ev[3 * k + 1], it_ev[3 * k + 2]);
            stats.solve_ms += ms;
        }
    }
    if (T_out)
        for (int i = 0; i < 16; ++i) T_out[i] = (double) last.T_final[i];
    if (converged) *converged = last.converged;
    if (iterations) *iterations = last.iter;
    if (state) *state = last.state;
    return WAVECU_OK;
}

void IcpHandle::release() {
    cudaSetDevice(device);
    src.release();
    tgt.release();
    for (void *p : {(void *) d_nn_pos, (void *) d_nn_idx, (void *) d_nn_d2, (void *) d_out_idx, (void *) d_out_d2,
                    (void *) d_aligned, (void *) d_mc, (void *) d_st, (void *) d_acc, (void *) d_trace})
        if (p) cudaFree(p);
    if (h_done) cudaFreeHost(h_done);
    if (h_st) cudaFreeHost(h_st);
    for (auto &e : ev_ring)
        if (e) cudaEventDestroy(e);
    for (auto e : ev_pool) cudaEventDestroy(e);
    if (own_stream && stream) cudaStreamDestroy(stream);
}

}  // namespace wavecu

using namespace wavecu;

struct wavecu_icp {
    IcpHandle h;
};

extern "C" {

void wavecu_icp_default_params(wavecu_icp_params *p) {
    if (!p) return;
    p->max_corr = 3;
    p->max_iter = 100;
    p->t_eps = 1e-8;
    p->fit_eps = 1e-2;
    p->lidar_ang_covar = 7.78e-9;
    p->lidar_lin_covar = 2.5e-4;
    p->multiscale_steps = 3;
    p->res = 0.1f;
    p->covar_estimator = WAVECU_INFO_LUM;
    p->estimator = WAVECU_EST_SVD;
}

int wavecu_icp_create(const wavecu_icp_params *params, int device, void *stream, wavecu_icp **out) {
    if (!out) return WAVECU_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) {
        set_last_error("no such CUDA device (libwavecu has no CPU fallback)");
        return WAVECU_ERR_CUDA;
    }
    wavecu_icp *w = new wavecu_icp();
    if (params) w->h.prm = *params;
    else wavecu_icp_default_params(&w->h.prm);
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

int wavecu_icp_destroy(wavecu_icp *w) {
    if (!w) return WAVECU_OK;
    w->h.release();
    delete w;
    return WAVECU_OK;
}

int wavecu_icp_set_params(wavecu_icp *w, const wavecu_icp_params *params) {
    if (!w || !params) return WAVECU_ERR_ARG;
    w->h.prm = *params;
    return WAVECU_OK;
}

int wavecu_icp_set_source(wavecu_icp *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    w->h.src_dirty = true;
    w->h.have_result = false;
    return w->h.src.upload(xyzw, n, false);
}
int wavecu_icp_set_source_device(wavecu_icp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    w->h.src_dirty = true;
    w->h.have_result = false;
    return w->h.src.upload((const float *) d, n, true);
}
int wavecu_icp_set_target(wavecu_icp *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    w->h.have_result = false;
    return w->h.tgt.set_points(xyzw, n, false);
}
int wavecu_icp_set_target_device(wavecu_icp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    w->h.have_result = false;
    return w->h.tgt.set_points((const float *) d, n, true);
}
int wavecu_icp_set_target_normals(wavecu_icp *w, const float *nxyzw, size_t n) {
    if (!w || (!nxyzw && n)) return WAVECU_ERR_ARG;
    return w->h.tgt.set_normals(nxyzw, n, false);
}
int wavecu_icp_set_target_normals_device(wavecu_icp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    return w->h.tgt.set_normals((const float *) d, n, true);
}

int wavecu_icp_align(wavecu_icp *w, double T_out[16], int *converged, int *iterations, int *state) {
    if (!w) return WAVECU_ERR_ARG;
    return w->h.align(T_out, converged, iterations, state);
}

int wavecu_icp_match(wavecu_icp *w, double T_out[16], int *converged, int *iterations) {
    if (!w) return WAVECU_ERR_ARG;
    if (w->h.prm.res > 0) {
        set_last_error("voxel-filtered / multiscale match() is not built yet: set res <= 0");
        return WAVECU_ERR_STATE;
    }
    // full-resolution branch, src/icp.cpp:123-131
    return w->h.align(T_out, converged, iterations, nullptr);
}

int wavecu_icp_correspondences(wavecu_icp *w, int *idx_query, int *idx_match, float *dist2, size_t *n) {
    if (!w || !n) return WAVECU_ERR_ARG;
    IcpHandle &h = w->h;
    if (!h.have_result) {
        set_last_error("no align() result");
        return WAVECU_ERR_STATE;
    }
    WCU_CHECK(cudaSetDevice(h.device));
    const size_t ns = h.result_n_src;
    *n = 0;
    if (ns == 0 || h.last.iter == 0) return WAVECU_OK;
    fill_int_kernel<<<(unsigned) ((ns + 255) / 256), 256, 0, h.stream>>>(h.d_out_idx, -1, ns);
    unsort_corr_kernel<<<(unsigned) ((ns + 255) / 256), 256, 0, h.stream>>>(h.src.d_sorted, h.d_nn_idx, h.d_nn_d2,
                                                                              (int) ns, h.d_out_idx, h.d_out_d2);
    std::vector<int> idx(ns);
    std::vector<float> d2(ns);
    WCU_CHECK(cudaMemcpyAsync(idx.data(), h.d_out_idx, ns * sizeof(int), cudaMemcpyDeviceToHost, h.stream));
    WCU_CHECK(cudaMemcpyAsync(d2.data(), h.d_out_d2, ns * sizeof(float), cudaMemcpyDeviceToHost, h.stream));
    WCU_CHECK(cudaStreamSynchronize(h.stream));
    size_t c = 0;
    for (size_t i = 0; i < ns; ++i) {
        if (idx[i] < 0) continue;
        if (idx_query) idx_query[c] = (int) i;
        if (idx_match) idx_match[c] = idx[i];
        if (dist2) dist2[c] = d2[i];
        ++c;
    }
    *n = c;
    return WAVECU_OK;
}

int wavecu_icp_aligned(wavecu_icp *w, float *xyzw, size_t *n) {
    if (!w || !n) return WAVECU_ERR_ARG;
    IcpHandle &h = w->h;
    if (!h.have_result) {
        set_last_error("no align() result");
        return WAVECU_ERR_STATE;
    }
    *n = h.result_n_src;
    if (!xyzw || h.result_n_src == 0) return WAVECU_OK;
    WCU_CHECK(cudaSetDevice(h.device));
    const size_t ns = h.result_n_src;
    apply_final_kernel<<<(unsigned) ((ns + 255) / 256), 256, 0, h.stream>>>(h.src.d_raw, ns, h.d_st, h.d_aligned);
    WCU_CHECK(cudaMemcpyAsync(xyzw, h.d_aligned, ns * sizeof(float4), cudaMemcpyDeviceToHost, h.stream));
    WCU_CHECK(cudaStreamSynchronize(h.stream));
    return WAVECU_OK;
}

int wavecu_icp_trace(wavecu_icp *w, double *mse, int *n_corr, float *T_inc, int *n) {
    if (!w || !n) return WAVECU_ERR_ARG;
    IcpHandle &h = w->h;
    if (!h.have_result) {
        set_last_error("no align() result");
        return WAVECU_ERR_STATE;
    }
    *n = (int) h.trace.size();
    for (size_t i = 0; i < h.trace.size(); ++i) {
        if (mse) mse[i] = h.trace[i].mse;
        if (n_corr) n_corr[i] = h.trace[i].n_corr;
        if (T_inc) std::memcpy(T_inc + 16 * i, h.trace[i].T, sizeof(float) * 16);
    }
    return WAVECU_OK;
}

int wavecu_icp_info(wavecu_icp *w, int method, double info_out[36]) {
    (void) method;
    (void) info_out;
    if (!w) return WAVECU_ERR_ARG;
    set_last_error("information estimators are not built yet");
    return WAVECU_ERR_STATE;
}

int wavecu_icp_set_profiling(wavecu_icp *w, int enabled) {
    if (!w) return WAVECU_ERR_ARG;
    w->h.profiling = enabled != 0;
    return WAVECU_OK;
}

int wavecu_icp_stats(wavecu_icp *w, wavecu_stats *out) {
    if (!w || !out) return WAVECU_ERR_ARG;
    *out = w->h.stats;
    return WAVECU_OK;
}

}  // extern "C"
