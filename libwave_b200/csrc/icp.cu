// Host side of the ICP handle behind include/wavecu.h (wavecu_icp_*): device buffers, the
// per-align() launch sequence and the C ABI.  Reference behaviour restated: pcl::Registration::
// align + IterativeClosestPoint::computeTransformation as driven by wave_matching/src/icp.cpp
// :47-50 (parameters) and :75-133 (ICPMatcher::match).
#include <sched.h>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/wavecu.h"
#include "icp_kernels.cuh"
#include "index.cuh"
#include "info_kernels.cuh"
#include "tile_nn.cuh"
#include "voxel.cuh"

namespace wavecu {

namespace {

struct SetupArgs {
    const unsigned *src_bbox, *tgt_bbox;
    MatchConsts *mc;
    IcpState *st;
    Acc128 *acc;
    double max_corr;
    int *fb_count;
    unsigned *ticket;
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
    if (a.fb_count) *a.fb_count = 0;
    if (a.ticket) *a.ticket = 0u;
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
    st.pad = 0;
    st.fb_total = 0;
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
    // Three streams per handle.  `stream` (the caller's, or our own): target build, iterations, results.
    // `aux`: the source cloud's sort, which overlaps the target build.  `copy`: every upload, so that
    // the H2D copy of one cloud overlaps the sort / tree build of the previous one (a 1 M-point
    // match moves 48 MB over PCIe - as long as the whole device-side build).
    cudaStream_t stream = nullptr, aux = nullptr, copy = nullptr;
    bool own_stream = false;
    cudaEvent_t ev_main = nullptr, ev_src_sorted = nullptr, ev_first = nullptr;
    bool first_recorded = false, first_marked = false;
    long long launches_mark = 0;
    MortonCloud src;
    TargetIndex tgt;
    bool src_dirty = true;
    bool src_sorted_fresh = false;   // src.d_sorted holds an unconsumed Morton sort of the uploaded source
    // A page-locked host source is not copied at once: setRef comes before setTarget, but it is the target
    // whose tree the first iteration waits for, so its copy goes first and the source follows it
    // (flushed by set_target, or by whatever needs the source next).
    const float *pending_src = nullptr;
    size_t pending_src_n = 0;

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
    // tiled correspondence kernel (tile_nn.cuh): queries it hands to the LBVH walk
    // scan-to-map batches: the target index of another handle on the same device, read-only here
    // (wavecu_icp_share_target); its owner records ev_tgt_ready behind the build
    IcpHandle *tgt_owner = nullptr;
    cudaEvent_t ev_tgt_ready = nullptr;
    unsigned long long tgt_epoch = 0, tgt_epoch_seen = 0;
    TargetIndex &target() { return tgt_owner ? tgt_owner->tgt : tgt; }
    bool use_tile = false;
    bool tgt_set_since_align = false;   // set_target came before set_source for the coming match
    bool use_fused = true;    // one launch per iteration (iterate_kernel); WAVECU_FUSED=0: correspond / reduce / solve
    unsigned *d_ticket = nullptr;
    int *d_fb_count = nullptr, *d_fb_list = nullptr;

    // pinned host mirrors
    int *h_progress = nullptr;  // mapped: (launch number << 1) | done, written by the solve kernel
    IcpState *h_st = nullptr;
    TraceRow *h_trace = nullptr;  // pinned: the first kTracePrefetch trace rows
    static constexpr int kTracePrefetch = 16;
    static constexpr int kDepth = 3;

    // results of the last align
    bool have_result = false;
    size_t result_n_src = 0;
    IcpState last;
    double result[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // Matcher::result of the last match
    double *d_censi = nullptr;      // per-block partial sums of the Censi estimator
    double *h_censi = nullptr;
    int censi_blocks = 0;
    std::vector<TraceRow> trace;

    // voxel-filtered / multiscale path: the clouds as the caller gave them
    VoxelWork vox;
    float4 *d_orig_src = nullptr, *d_orig_tgt = nullptr;
    size_t orig_src_cap = 0, orig_tgt_cap = 0, n_orig_src = 0, n_orig_tgt = 0;
    bool src_is_user = true, tgt_is_user = true;   // working clouds still hold the caller's data
    bool have_orig_src = false, have_orig_tgt = false;
    MatchConsts last_mc{};
    int *d_pos2 = nullptr;  // estimateLUMold correspondences
    size_t pos2_cap = 0;
    Acc128 *h_acc = nullptr;  // pinned mirror of the accumulators

    // profiling
    bool profiling = false;
    wavecu_stats stats{};
    std::vector<cudaEvent_t> ev_pool;

    int init();
    int on_set_source(bool from_device);
    int on_set_target(bool from_device);
    int set_source(const float *xyzw, size_t n, bool from_device);
    int flush_pending_source();
    int set_target(const float *xyzw, size_t n, bool from_device);
    int set_target_normals(const float *nxyzw, size_t n, bool from_device);
    int mark_first();
    int ensure_iter_buffers(size_t n_src_pad, int max_iter);
    int align(double *T_out, int *converged, int *iterations, int *state);
    int stash_originals();
    int restore_originals();
    int load_level(float leaf, const double *running);
    int match(double *T_out, int *converged, int *iterations);
    int info(int method, double *info_out);
    int info_censi(double *info_out);
    int read_acc(int n_values, __int128 *out);
    void release();
};

int IcpHandle::init() {
    WCU_CHECK(cudaSetDevice(device));
    if (!stream) {
        WCU_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }
    // the source order only has to make the queries of a warp neighbours: 12 bits per axis (37 sorted
    // bits, 5 passes on the aux stream, off the critical path) measured 4 % faster searches than 10
    src.key_bits = 12;
    // target: 13 bits (extent / 8192 cells, 5 radix passes) measured best for ~1 M-point lidar targets -
    // finer keys do not improve the tree (12: +2 % search time, 14-17: equal) and cost another sort pass
    // (profiles/r01_nn_variants.md).  GICP keeps 15: its fp64 cost sums run in Morton order, and its
    // iteration-for-iteration agreement with the oracle was established with that order.
    tgt.cloud.key_bits = 13;
    if (const char *e = getenv("WAVECU_SRC_BITS")) src.key_bits = std::max(1, std::min(21, atoi(e)));  // tuning knob
    if (const char *e = getenv("WAVECU_TGT_BITS")) tgt.cloud.key_bits = std::max(1, std::min(21, atoi(e)));
    WCU_CHECK(cudaStreamCreateWithFlags(&aux, cudaStreamNonBlocking));
    WCU_CHECK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
    WCU_CHECK(cudaEventCreateWithFlags(&ev_main, cudaEventDisableTiming));
    WCU_CHECK(cudaEventCreateWithFlags(&ev_src_sorted, getenv("WAVECU_TIMELINE") ? cudaEventDefault
                                                                                  : cudaEventDisableTiming));
    WCU_CHECK(cudaEventCreate(&ev_first));
    if (const char *e = getenv("WAVECU_NN")) use_tile = std::string(e) == "tile";   // default of wavecu_icp_set_search
    tgt.want_boxes = use_tile;
    WCU_CHECK(cudaFuncSetAttribute(correspond_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int) sizeof(TileSmem)));
    if (const char *e = getenv("WAVECU_FUSED")) use_fused = atoi(e) != 0;   // tuning knob
    WCU_CHECK(cudaMalloc((void **) &d_fb_count, sizeof(int)));
    WCU_CHECK(cudaMemset(d_fb_count, 0, sizeof(int)));
    WCU_CHECK(cudaMalloc((void **) &d_ticket, sizeof(unsigned)));
    WCU_CHECK(cudaMemset(d_ticket, 0, sizeof(unsigned)));
    src.device = tgt.cloud.device = device;
    src.stream = aux;
    tgt.cloud.stream = stream;
    src.copy_stream = tgt.cloud.copy_stream = copy;
    // shared-memory carve-out left at the driver default (32 KB here): measured 5 % faster than MaxL1
    if (const char *e = getenv("WAVECU_CARVEOUT"))  // tuning knob
        cudaFuncSetAttribute(correspond_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
    WCU_CHECK(cudaMalloc((void **) &d_mc, sizeof(MatchConsts)));
    WCU_CHECK(cudaMalloc((void **) &d_st, sizeof(IcpState)));
    WCU_CHECK(cudaMalloc((void **) &d_acc, sizeof(Acc128) * kAccSlots * kMaxAcc));
    WCU_CHECK(cudaHostAlloc((void **) &h_progress, sizeof(int), cudaHostAllocMapped));
    WCU_CHECK(cudaHostAlloc((void **) &h_st, sizeof(IcpState), cudaHostAllocDefault));
    WCU_CHECK(cudaHostAlloc((void **) &h_trace, sizeof(TraceRow) * kTracePrefetch, cudaHostAllocDefault));
    WCU_CHECK(cudaHostAlloc((void **) &h_acc, sizeof(Acc128) * kAccSlots * kMaxAcc, cudaHostAllocDefault));
    vox.device = device;
    vox.stream = stream;
    return WAVECU_OK;
}

int IcpHandle::mark_first() {
    if (!first_marked) {
        launches_mark = src.launches + tgt.cloud.launches;
        first_marked = true;
    }
    if (profiling && !first_recorded) {
        WCU_CHECK(cudaEventRecord(ev_first, stream));
        first_recorded = true;
    }
    return WAVECU_OK;
}

// Uploads are asynchronous; at full resolution the source sort (aux stream) and the target build
// (main stream) are queued right behind their copies, so they run while the next cloud is still
// crossing PCIe.  Device-resident inputs were produced on the caller's stream: the copy waits for it.
int IcpHandle::on_set_source(bool from_device) {
    src_is_user = true;
    have_orig_src = false;
    src_dirty = true;
    have_result = false;
    src_sorted_fresh = false;
    int rc = mark_first();
    if (rc) return rc;
    if (from_device) {  // the D2D copy and the sort run on aux: order them behind the caller's stream
        WCU_CHECK(cudaEventRecord(ev_main, stream));
        WCU_CHECK(cudaStreamWaitEvent(aux, ev_main, 0));
    }
    return WAVECU_OK;
}

int IcpHandle::on_set_target(bool from_device) {
    tgt_is_user = true;
    have_orig_tgt = false;
    have_result = false;
    (void) from_device;  // device-resident targets are copied on the main stream itself
    return mark_first();
}

int IcpHandle::set_source(const float *xyzw, size_t n, bool from_device) {
    WCU_CHECK(cudaSetDevice(device));
    int rc = on_set_source(from_device);
    if (rc) return rc;
    pending_src = nullptr;
    if (!from_device && n) {
        // the caller keeps page-locked buffers untouched until match() returns (wavecu.h): defer
        // (only while the target of this match has not been given yet: a caller who sets the target first gets
        // the source copy right behind it, ahead of the normals)
        cudaPointerAttributes at;
        if (!tgt_set_since_align && cudaPointerGetAttributes(&at, xyzw) == cudaSuccess && at.type == cudaMemoryTypeHost) {
            pending_src = xyzw;
            pending_src_n = n;
            return WAVECU_OK;
        }
        cudaGetLastError();
    }
    rc = src.upload(xyzw, n, from_device);
    if (rc) return rc;
    if (!(prm.res > 0) && n) {  // full resolution: this cloud is the working cloud - sort it right away
        rc = src.sort(n);
        if (rc) return rc;
        WCU_CHECK(cudaEventRecord(ev_src_sorted, aux));
        src_sorted_fresh = true;
    }
    return WAVECU_OK;
}

int IcpHandle::flush_pending_source() {
    if (!pending_src) return WAVECU_OK;
    const float *p = pending_src;
    const size_t n = pending_src_n;
    pending_src = nullptr;
    int rc = src.upload(p, n, false);
    if (rc) return rc;
    if (!(prm.res > 0) && n) {
        rc = src.sort(n);
        if (rc) return rc;
        WCU_CHECK(cudaEventRecord(ev_src_sorted, aux));
        src_sorted_fresh = true;
    }
    return WAVECU_OK;
}

int IcpHandle::set_target(const float *xyzw, size_t n, bool from_device) {
    WCU_CHECK(cudaSetDevice(device));
    int rc = on_set_target(from_device);
    if (rc) return rc;
    tgt_owner = nullptr;   // a target of its own ends the sharing of another handle's
    tgt_set_since_align = true;
    rc = tgt.set_points(xyzw, n, from_device);
    if (rc) return rc;
    if (!(prm.res > 0) && n) {  // full resolution: build the search tree behind the copy
        rc = tgt.build();
        if (rc) return rc;
    }
    return flush_pending_source();  // the deferred source copy goes behind the target's
}

int IcpHandle::set_target_normals(const float *nxyzw, size_t n, bool from_device) {
    WCU_CHECK(cudaSetDevice(device));
    return tgt.set_normals(nxyzw, n, from_device);
}

int IcpHandle::ensure_iter_buffers(size_t n_src_pad, int max_iter) {
    if (n_src_pad > iter_cap) {
        for (void *p : {(void *) d_nn_pos, (void *) d_nn_idx, (void *) d_nn_d2, (void *) d_out_idx, (void *) d_out_d2,
                        (void *) d_aligned, (void *) d_fb_list})
            if (p) WCU_CHECK(cudaFree(p));
        d_nn_pos = d_nn_idx = d_out_idx = d_fb_list = nullptr;
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
        WCU_CHECK(cudaMalloc((void **) &d_fb_list, a * sizeof(int)));
        iter_cap = a;
    }
    if (max_iter > trace_cap) {
        if (d_trace) WCU_CHECK(cudaFree(d_trace));
        d_trace = nullptr;
        WCU_CHECK(cudaMalloc((void **) &d_trace, sizeof(TraceRow) * (size_t) max_iter));
        WCU_CHECK(cudaMemsetAsync(d_trace, 0, sizeof(TraceRow) * (size_t) max_iter, stream));   // rows of launches past convergence are read back unwritten
        trace_cap = max_iter;
    }
    return WAVECU_OK;
}

int IcpHandle::align(double *T_out, int *converged, int *iterations, int *state) {
    WCU_CHECK(cudaSetDevice(device));
    {
        const int rc = flush_pending_source();
        if (rc) return rc;
    }
    have_result = false;
    tgt_set_since_align = false;
    TargetIndex &TG = target();
    if (tgt_owner) {
        if (TG.dirty) {
            set_last_error("the shared target has not been built (wavecu_icp_build_target on its owner)");
            return WAVECU_ERR_STATE;
        }
        if (tgt_epoch_seen != tgt_owner->tgt_epoch) {   // once per (re)build of the map
            WCU_CHECK(cudaStreamWaitEvent(stream, tgt_owner->ev_tgt_ready, 0));
            tgt_epoch_seen = tgt_owner->tgt_epoch;
        }
    }
    const size_t n_src = src.n, n_tgt = TG.cloud.n;
    const int max_iter = std::max(prm.max_iter, 1);
    stats = wavecu_stats{};
    const long long launches0 = first_marked ? launches_mark : src.launches + tgt.cloud.launches;
    first_marked = false;
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

    bool late_normals = false;
    // ---- build: Morton-sort both clouds, AABB tree over the target ----
    const size_t n_src_pad = n_src;
    int rc = ensure_iter_buffers(std::max<size_t>(n_src_pad, 1), max_iter);
    if (rc) return rc;
    // the working cloud is consumed by the iterations, so the source is re-sorted for every align
    // (on the aux stream, concurrently with the target build) unless set_source already queued it
    if (!src_sorted_fresh) {
        WCU_CHECK(cudaEventRecord(ev_main, stream));   // load_level / restore write src.d_raw on the main stream
        WCU_CHECK(cudaStreamWaitEvent(aux, ev_main, 0));
        rc = src.sort(std::max<size_t>(n_src_pad, 1));
        if (rc) return rc;
        WCU_CHECK(cudaEventRecord(ev_src_sorted, aux));
    }
    if (!tgt_owner && TG.dirty) {
        rc = TG.build();
        if (rc) return rc;
    }
    if (prm.estimator == WAVECU_EST_POINT_TO_PLANE) {
        if (TG.nrm_n == n_tgt) {
            // the caller's normals are only read by the first reduction: their gather into Morton order is
            // queued behind the first correspondence launch, so that search does not wait for their upload
            // ... which only pays while that upload is still crossing PCIe; device-resident normals are
            // gathered right away, and every iteration then runs as one fused launch
            late_normals = !tgt_owner && TG.nrm_dirty && TG.nrm_up_pending;
            rc = TG.reserve_sorted_normals();
            if (rc) return rc;
            if (!tgt_owner && TG.nrm_dirty && !late_normals) {
                rc = TG.sort_normals();
                if (rc) return rc;
            }
        } else if (!TG.normals_estimated) {
            // no normals from the caller: estimate them on the target's own tree (k = 10 neighbours)
            rc = TG.estimate_normals(10);
            if (rc) return rc;
        }
    }
    WCU_CHECK(cudaStreamWaitEvent(stream, ev_src_sorted, 0));
    src_sorted_fresh = false;
    src_dirty = false;
    long long extra_launches = 0;
    SetupArgs sa{src.d_bbox, TG.cloud.d_bbox, d_mc, d_st, d_acc, prm.max_corr, d_fb_count, d_ticket};
    setup_kernel<<<1, 256, 0, stream>>>(sa);
    ++extra_launches;
    WCU_CHECK(cudaGetLastError());
    if (profiling) WCU_CHECK(cudaEventRecord(e_built, stream));

    // ---- iterate: fused correspondence+reduction kernel, then the one-thread estimator ----
    IterArgs ia;
    ia.cur = src.d_sorted;
    ia.n_src = (int) n_src;
    ia.ix = TG.index();
    ia.tgt = TG.cloud.d_sorted;
    ia.nrm = TG.d_nrm_sorted;
    ia.nn_pos = d_nn_pos;
    ia.nn_idx = d_nn_idx;
    ia.nn_d2 = d_nn_d2;
    ia.st = d_st;
    ia.mc = d_mc;
    ia.acc = d_acc;
    int *d_progress = nullptr;
    WCU_CHECK(cudaHostGetDevicePointer((void **) &d_progress, h_progress, 0));
    *(volatile int *) h_progress = 0;
    SolveArgs so{d_st, d_mc, d_acc, d_trace, max_iter, prm.t_eps, prm.fit_eps, d_progress, 0, d_fb_count};
    const unsigned grid_nn = (unsigned) std::max<size_t>(1, (n_src + kIterThreads - 1) / kIterThreads);
    const bool tiled = use_tile && TG.want_boxes;
    const unsigned grid_tile = (unsigned) std::max<size_t>(1, (n_src + kTileQ - 1) / kTileQ);
    const unsigned grid_fb = std::min<unsigned>(grid_nn, 148u * 8u);
    const TileFallback fb{d_fb_count, d_fb_list};
    const unsigned grid_red = (unsigned) std::max<size_t>(1, (n_src + kReduceThreads - 1) / kReduceThreads);
    std::vector<cudaEvent_t> it_ev;
    int launched = 0;
    long long launches_total = 0;
    bool finished = (n_src == 0 || n_tgt == 0);  // initCompute fails -> converged_ = false
    for (int k = 0; !finished && k < max_iter; ++k) {
        if (profiling) {
            it_ev.push_back(next_event());
            WCU_CHECK(cudaEventRecord(it_ev.back(), stream));
        }
        so.launch = k + 1;
        // one launch per iteration, except while the caller's normals are still on their way (they are
        // gathered behind the first search) and with the tiled search (its own kernel pair)
        const bool fused = use_fused && !tiled && !late_normals;
        if (fused) {
            const FusedArgs fa{ia, so, d_ticket};
            if (prm.estimator == WAVECU_EST_POINT_TO_PLANE)
                iterate_kernel<WAVECU_EST_POINT_TO_PLANE><<<grid_nn, kIterThreads, 0, stream>>>(fa);
            else
                iterate_kernel<WAVECU_EST_SVD><<<grid_nn, kIterThreads, 0, stream>>>(fa);
            launches_total += 1;
        } else if (tiled) {
            correspond_tile_kernel<<<grid_tile, kTileQ, sizeof(TileSmem), stream>>>(ia, TG.boxes, fb);
            tile_fallback_kernel<<<grid_fb, kIterThreads, 0, stream>>>(ia, fb);
            launches_total += 4;
        } else {
            correspond_kernel<<<grid_nn, kIterThreads, 0, stream>>>(ia);
            launches_total += 3;
        }
        if (profiling) {
            it_ev.push_back(next_event());
            WCU_CHECK(cudaEventRecord(it_ev.back(), stream));
        }
        if (late_normals) {
            late_normals = false;
            rc = TG.sort_normals();
            if (rc) return rc;
        }
        if (!fused) {
            if (prm.estimator == WAVECU_EST_POINT_TO_PLANE) {
                reduce_kernel<WAVECU_EST_POINT_TO_PLANE><<<grid_red, kReduceThreads, 0, stream>>>(ia);
                solve_kernel<WAVECU_EST_POINT_TO_PLANE><<<1, 64, 0, stream>>>(so);
            } else {
                reduce_kernel<WAVECU_EST_SVD><<<grid_red, kReduceThreads, 0, stream>>>(ia);
                solve_kernel<WAVECU_EST_SVD><<<1, 64, 0, stream>>>(so);
            }
        }
        if (profiling) {
            it_ev.push_back(next_event());
            WCU_CHECK(cudaEventRecord(it_ev.back(), stream));
        }
        ++launched;
        // follow the device kDepth-1 iterations behind (it always has work queued): the solve kernel of
        // launch j publishes (j << 1) | done in mapped host memory
        if (k >= kDepth - 1) {
            const int need = k - (kDepth - 1) + 1;
            int v, spins = 0;
            while (((v = *(volatile int *) h_progress) >> 1) < need) {
                if (++spins % 4096 == 0) {
                    const cudaError_t q = cudaStreamQuery(stream);
                    if (q != cudaErrorNotReady) {  // finished (or failed) without publishing: stop waiting
                        if (q != cudaSuccess) WCU_CHECK(q);
                        v = *(volatile int *) h_progress;
                        if ((v >> 1) < need) v |= 1;
                        break;
                    }
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
                // a thread that has been waiting for a while (~0.1 ms) gives its core away now and then: boxes with fewer
                // cores than matcher threads (8 ranks x several matchers on 32 cores); with a free core it returns at once
                if (spins > 2048 && (spins & 255) == 255) sched_yield();
            }
            if (v & 1) finished = true;
        }
    }
    if (profiling) WCU_CHECK(cudaEventRecord(e_end, stream));
    WCU_CHECK(cudaMemcpyAsync(h_st, d_st, sizeof(IcpState), cudaMemcpyDeviceToHost, stream));
    WCU_CHECK(cudaMemcpyAsync(&last_mc, d_mc, sizeof(MatchConsts), cudaMemcpyDeviceToHost, stream));
    // most matches stop within a few iterations: their trace rows ride along with the state read-back
    const int n_pre = std::min(launched, kTracePrefetch);
    if (n_pre > 0)
        WCU_CHECK(cudaMemcpyAsync(h_trace, d_trace, sizeof(TraceRow) * (size_t) n_pre, cudaMemcpyDeviceToHost, stream));
    WCU_CHECK(cudaStreamSynchronize(stream));
    WCU_CHECK(cudaGetLastError());
    last = *h_st;
    if (n_src == 0 || n_tgt == 0) {
        last.converged = 0;
        last.state = WAVECU_CONV_NOT_CONVERGED;
    }
    trace.resize((size_t) std::max(0, last.iter));
    if (last.iter > 0 && last.iter <= n_pre)
        std::memcpy(trace.data(), h_trace, sizeof(TraceRow) * (size_t) last.iter);
    else if (last.iter > 0)
        WCU_CHECK(cudaMemcpy(trace.data(), d_trace, sizeof(TraceRow) * (size_t) last.iter, cudaMemcpyDeviceToHost));
    have_result = true;
    result_n_src = n_src;

    stats.iterate_launches = last.iter;
    stats.kernel_launches = (src.launches + tgt.cloud.launches - launches0) + extra_launches + launches_total;
    stats.pairs = (long long) last.iter * (long long) n_src;
    stats.fallback_queries = last.fb_total;
    if (profiling) {
        float ms = 0;
        // the build may have been queued by set_source / set_target already: count from the first of them
        cudaEvent_t e_start = first_recorded ? ev_first : e_begin;
        cudaEventElapsedTime(&ms, e_start, e_built);
        stats.build_ms = ms;
        cudaEventElapsedTime(&ms, e_start, e_end);
        stats.total_ms = ms;
        if (getenv("WAVECU_TIMELINE") && first_recorded) {  // debugging aid: where the streams were, in ms
            auto at = [&](cudaEvent_t e) {
                float t = -1.0f;
                if (e && cudaEventElapsedTime(&t, ev_first, e) != cudaSuccess) { cudaGetLastError(); t = -1.0f; }
                return t;
            };
            fprintf(stderr, "[wavecu timeline ms] src_up %.3f tgt_up %.3f nrm_up %.3f src_sorted %.3f tgt_used %.3f "
                            "align_begin %.3f built %.3f end %.3f\n", at(src.ev_up), at(tgt.cloud.ev_up),
                    at(tgt.ev_nrm_up), at(ev_src_sorted), at(tgt.cloud.ev_used), at(e_begin), at(e_built), at(e_end));
            if (src.ev_tl[0])
                fprintf(stderr, "[wavecu timeline ms] src keys %.3f sorted %.3f gathered %.3f | tgt keys %.3f sorted %.3f "
                                "gathered %.3f\n", at(src.ev_tl[0]), at(src.ev_tl[1]), at(src.ev_tl[2]),
                        at(tgt.cloud.ev_tl[0]), at(tgt.cloud.ev_tl[1]), at(tgt.cloud.ev_tl[2]));
        }
        first_recorded = false;
        // only iterations that did work count (speculative launches past `done` return at once)
        for (int k = 0; k < last.iter + (last.state == WAVECU_CONV_NO_CORRESPONDENCES ? 1 : 0) &&
                        (size_t) (3 * k + 2) < it_ev.size(); ++k) {
            cudaEventElapsedTime(&ms, it_ev[3 * k], it_ev[3 * k + 1]);
            stats.iterate_ms += ms;
            cudaEventElapsedTime(&ms, it_ev[3 * k + 1], it_ev[3 * k + 2]);
            stats.solve_ms += ms;
        }
    }
    for (int i = 0; i < 16; ++i) result[i] = (double) last.T_final[i];
    if (T_out)
        for (int i = 0; i < 16; ++i) T_out[i] = (double) last.T_final[i];
    if (converged) *converged = last.converged;
    if (iterations) *iterations = last.iter;
    if (state) *state = last.state;
    return WAVECU_OK;
}

// ---- ICPMatcher::match(), src/icp.cpp:75-133 -------------------------------------------------------
int IcpHandle::stash_originals() {
    // the caller's clouds may still be in flight on the copy / aux streams
    WCU_CHECK(cudaStreamSynchronize(copy));
    WCU_CHECK(cudaStreamSynchronize(aux));
    if (src_is_user) {
        if (src.n > orig_src_cap) {
            if (d_orig_src) WCU_CHECK(cudaFree(d_orig_src));
            d_orig_src = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_orig_src, (src.n + 64) * sizeof(float4)));
            orig_src_cap = src.n + 64;
        }
        if (src.n) WCU_CHECK(cudaMemcpyAsync(d_orig_src, src.d_raw, src.n * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        n_orig_src = src.n;
        have_orig_src = true;
        src_is_user = false;
    }
    if (tgt_is_user) {
        if (tgt.cloud.n > orig_tgt_cap) {
            if (d_orig_tgt) WCU_CHECK(cudaFree(d_orig_tgt));
            d_orig_tgt = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_orig_tgt, (tgt.cloud.n + 64) * sizeof(float4)));
            orig_tgt_cap = tgt.cloud.n + 64;
        }
        if (tgt.cloud.n)
            WCU_CHECK(cudaMemcpyAsync(d_orig_tgt, tgt.cloud.d_raw, tgt.cloud.n * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
        n_orig_tgt = tgt.cloud.n;
        have_orig_tgt = true;
        tgt_is_user = false;
    }
    return WAVECU_OK;
}

int IcpHandle::restore_originals() {
    if (!src_is_user && have_orig_src) {
        int rc = src.upload((const float *) d_orig_src, n_orig_src, true);
        if (rc) return rc;
        src_is_user = true;
        src_sorted_fresh = false;
    }
    if (!tgt_is_user && have_orig_tgt) {
        int rc = tgt.set_points((const float *) d_orig_tgt, n_orig_tgt, true);
        if (rc) return rc;
        tgt_is_user = true;
    }
    return WAVECU_OK;
}

// working clouds <- VoxelGrid(leaf) of the caller's clouds; the source additionally moved by the
// running transform (pcl::transformPointCloud with an Affine3d, src/icp.cpp:84-86)
int IcpHandle::load_level(float leaf, const double *running) {
    int rc = src.reserve(std::max<size_t>(n_orig_src, 1), 0);
    if (rc) return rc;
    size_t n_out = 0;
    rc = vox.filter(d_orig_src, n_orig_src, leaf, src.d_raw, &n_out, nullptr);
    if (rc) return rc;
    src.n = n_out;
    src_sorted_fresh = false;
    if (running) {
        rc = affine3d_inplace(src.d_raw, n_out, running, stream);
        if (rc) return rc;
    }
    rc = tgt.cloud.reserve(std::max<size_t>(n_orig_tgt, 1), 0);
    if (rc) return rc;
    rc = vox.filter(d_orig_tgt, n_orig_tgt, leaf, tgt.cloud.d_raw, &n_out, nullptr);
    if (rc) return rc;
    tgt.cloud.n = n_out;
    tgt.nrm_n = 0;
    tgt.dirty = true;
    return WAVECU_OK;
}

int IcpHandle::match(double *T_out, int *converged, int *iterations) {
    WCU_CHECK(cudaSetDevice(device));
    {
        const int rc = flush_pending_source();
        if (rc) return rc;
    }
    if (converged) *converged = 0;
    if (iterations) *iterations = 0;
    if (!(prm.res > 0)) {
        // full-resolution branch, src/icp.cpp:123-131
        int rc = restore_originals();
        if (rc) return rc;
        return align(T_out, converged, iterations, nullptr);
    }
    if (prm.estimator != WAVECU_EST_SVD) {
        set_last_error("the point-to-plane estimator needs per-point target normals: use res <= 0");
        return WAVECU_ERR_STATE;
    }
    if (tgt_owner) {
        set_last_error("a shared target is matched at full resolution only: use res <= 0");
        return WAVECU_ERR_STATE;
    }
    int rc = stash_originals();
    if (rc) return rc;
    const double user_max_corr = prm.max_corr;
    int total_iters = 0, conv = 0;
    double T[16];
    if (prm.multiscale_steps > 0) {
        double running[16];
        for (int i = 0; i < 16; ++i) running[i] = (i % 5 == 0) ? 1.0 : 0.0;
        for (int i = prm.multiscale_steps; i >= 0; --i) {
            const float leaf_size = std::pow(2, i) * prm.res;
            rc = load_level(leaf_size, running);
            if (rc) break;
            prm.max_corr = std::pow(2, i) * user_max_corr;
            int it = 0;
            rc = align(T, &conv, &it, nullptr);
            if (rc) break;
            total_iters += it;
            if (!conv) break;  // any level failing to converge fails the match, src/icp.cpp:96-98
            // running = final.cast<double>() * running (Eigen column order)
            double out[16];
            for (int r = 0; r < 4; ++r)
                for (int c = 0; c < 4; ++c) {
                    double acc = T[4 * r + 0] * running[0 * 4 + c];
                    acc = T[4 * r + 1] * running[1 * 4 + c] + acc;
                    acc = T[4 * r + 2] * running[2 * 4 + c] + acc;
                    acc = T[4 * r + 3] * running[3 * 4 + c] + acc;
                    out[4 * r + c] = acc;
                }
            std::memcpy(running, out, sizeof out);
        }
        prm.max_corr = user_max_corr;
        if (rc) return rc;
        if (conv) std::memcpy(T, running, sizeof running);
    } else {
        rc = load_level(prm.res, nullptr);
        if (rc) return rc;
        rc = align(T, &conv, &total_iters, nullptr);
        if (rc) return rc;
    }
    std::memcpy(result, T, sizeof T);  // Matcher::result of the multiscale / voxel branch
    if (T_out) std::memcpy(T_out, T, sizeof T);
    if (converged) *converged = conv;
    if (iterations) *iterations = total_iters;
    return WAVECU_OK;
}

// ---- information matrices ---------------------------------------------------------------------------
int IcpHandle::read_acc(int n_values, __int128 *out) {
    WCU_CHECK(cudaMemcpyAsync(h_acc, d_acc, sizeof(Acc128) * kAccSlots * kMaxAcc, cudaMemcpyDeviceToHost, stream));
    WCU_CHECK(cudaStreamSynchronize(stream));
    for (int i = 0; i < n_values; ++i) {
        __int128 t = 0;
        for (int sl = 0; sl < kAccSlots; ++sl) {
            const Acc128 &c = h_acc[sl * kMaxAcc + i];
            t += ((__int128) c.hi << 64) + (__int128) c.lo;
        }
        out[i] = t;
    }
    return WAVECU_OK;
}

namespace {

double fix_to_double(__int128 v, int k) {
    return acc_to_double((unsigned long long) v, (long long) (v >> 64), k);  // sign-magnitude, common.cuh
}

// N x N solve by Gaussian elimination with partial pivoting (fixed order), used column by column
// for MM.inverse() - the information estimators only
bool solve6_host(const double *A_in, const double *b_in, double *x) {
    double A[6][7];
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) A[i][j] = A_in[6 * i + j];
        A[i][6] = b_in[i];
    }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        double best = std::fabs(A[c][c]);
        for (int r = c + 1; r < 6; ++r)
            if (std::fabs(A[r][c]) > best) {
                best = std::fabs(A[r][c]);
                piv = r;
            }
        if (best == 0.0 || !std::isfinite(best)) return false;
        if (piv != c)
            for (int j = 0; j <= 6; ++j) std::swap(A[c][j], A[piv][j]);
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j <= 6; ++j) A[r][j] = A[r][j] - f * A[c][j];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = A[r][6];
        for (int j = r + 1; j < 6; ++j) s = s - A[r][j] * x[j];
        x[r] = s / A[r][r];
    }
    return true;
}

}  // namespace

int IcpHandle::info(int method, double *info_out) {
    WCU_CHECK(cudaSetDevice(device));
    if (!have_result) {
        set_last_error("estimateInfo needs a match() result");
        return WAVECU_ERR_STATE;
    }
    if (method == WAVECU_INFO_CENSI) return info_censi(info_out);
    // estimateLUM only acts if icp.hasConverged() (icp_pcl_functions.cpp:190): otherwise keep the
    // caller's matrix untouched - signalled by returning the identity the base class starts from
    const size_t ns = result_n_src;
    const int *pos = d_nn_pos;
    if (method == WAVECU_INFO_LUM && !last.converged) {
        for (int i = 0; i < 36; ++i) info_out[i] = (i % 7 == 0) ? 1.0 : 0.0;
        return WAVECU_OK;
    }
    if (ns == 0) {
        for (int i = 0; i < 36; ++i) info_out[i] = std::nan("");
        return WAVECU_OK;
    }
    if (method == WAVECU_INFO_LUMOLD) {
        if (ns > pos2_cap) {
            if (d_pos2) WCU_CHECK(cudaFree(d_pos2));
            d_pos2 = nullptr;
            WCU_CHECK(cudaMalloc((void **) &d_pos2, (ns + 64) * sizeof(int)));
            pos2_cap = ns + 64;
        }
        LumOldArgs oa;
        oa.cur = src.d_sorted;
        oa.raw = src.d_raw;
        oa.n_src = (int) ns;
        oa.ix = target().index();
        oa.warm = d_nn_pos;
        oa.st = d_st;
        const double max2 = prm.max_corr * prm.max_corr;
        float f = (max2 >= (double) FLT_MAX) ? FLT_MAX : (float) max2;
        if (!((double) f < max2)) f = std::nextafterf(f, -INFINITY);  // d2 < max^2, icp_pcl_functions.cpp:82
        oa.thr_strict = f;
        oa.pos_out = d_pos2;
        lumold_corr_kernel<<<(unsigned) ((ns + kIterThreads - 1) / kIterThreads), kIterThreads, 0, stream>>>(oa);
        pos = d_pos2;
    }
    LumArgs la;
    la.cur = src.d_sorted;
    la.raw = src.d_raw;
    la.n_src = (int) ns;
    la.tgt = target().cloud.d_sorted;
    la.pos = pos;
    la.st = d_st;
    la.acc = d_acc;
    const int k = last_mc.k_quad - 2, k_ss = last_mc.k_quad - 4;
    la.scale = std::ldexp(1.0, k);
    la.scale_ss = std::ldexp(1.0, k_ss);
    for (int i = 0; i < 6; ++i) la.D[i] = 0;
    const unsigned grid = (unsigned) std::max<size_t>(
        1, (ns + kReduceThreads * kReducePerThread - 1) / (kReduceThreads * kReducePerThread));
    zero_acc_kernel<<<1, 256, 0, stream>>>(d_acc);
    lum_kernel<0><<<grid, kReduceThreads, 0, stream>>>(la);
    __int128 tot[17];
    int rc = read_acc(17, tot);
    if (rc) return rc;
    const long long numCorr = (long long) tot[16];
    double v[15];
    for (int i = 0; i < 15; ++i) v[i] = fix_to_double(tot[i], k);
    double MM[36], MZ[6];
    for (int i = 0; i < 36; ++i) MM[i] = 0;
    auto M = [&](int r, int c) -> double & { return MM[6 * r + c]; };
    M(0, 4) = -v[1];
    M(0, 5) = v[2];
    M(1, 3) = -v[2];
    M(1, 4) = v[0];
    M(2, 3) = v[1];
    M(2, 5) = -v[0];
    M(3, 4) = -v[3];
    M(3, 5) = -v[4];
    M(4, 5) = -v[5];
    M(3, 3) = v[6];
    M(4, 4) = v[7];
    M(5, 5) = v[8];
    for (int i = 0; i < 6; ++i) MZ[i] = v[9 + i];
    M(0, 0) = M(1, 1) = M(2, 2) = static_cast<float>(numCorr);
    M(4, 0) = M(0, 4);
    M(5, 0) = M(0, 5);
    M(3, 1) = M(1, 3);
    M(4, 1) = M(1, 4);
    M(3, 2) = M(2, 3);
    M(5, 2) = M(2, 5);
    M(4, 3) = M(3, 4);
    M(5, 3) = M(3, 5);
    M(5, 4) = M(4, 5);
    // D = MM.inverse() * MZ
    double inv[36];
    bool ok = true;
    for (int c = 0; c < 6 && ok; ++c) {
        double e[6], x[6];
        for (int i = 0; i < 6; ++i) e[i] = (i == c) ? 1.0 : 0.0;
        ok = solve6_host(MM, e, x);
        for (int i = 0; i < 6; ++i) inv[6 * i + c] = x[i];
    }
    for (int r = 0; r < 6; ++r) {
        double sum = 0;
        for (int c = 0; c < 6; ++c) sum += inv[6 * r + c] * MZ[c];
        la.D[r] = ok ? sum : std::nan("");
    }
    float ss;
    if (ok) {
        zero_acc_kernel<<<1, 256, 0, stream>>>(d_acc);
        lum_kernel<1><<<grid, kReduceThreads, 0, stream>>>(la);
        rc = read_acc(2, tot);
        if (rc) return rc;
        ss = (tot[1] != 0) ? std::nanf("") : static_cast<float>(fix_to_double(tot[0], k_ss));
    } else {
        ss = std::nanf("");
    }
    zero_acc_kernel<<<1, 256, 0, stream>>>(d_acc);  // leave the accumulators clean for the next align
    WCU_CHECK(cudaGetLastError());
    if (ss < 0.0000000000001 || !std::isfinite(ss)) {
        // "Covariance matrix calculation was unsuccessful": estimateLUM returns the identity;
        // estimateLUMold falls through and overwrites it with MM / ss (icp_pcl_functions.cpp:170-178)
        for (int i = 0; i < 36; ++i) info_out[i] = (i % 7 == 0) ? 1.0 : 0.0;
        if (method == WAVECU_INFO_LUM) return WAVECU_OK;
    }
    const float rec = 1.0f / ss;
    for (int i = 0; i < 36; ++i) info_out[i] = MM[i] * rec;
    return WAVECU_OK;
}

namespace {

// Eigen 3.3 MatrixBase::eulerAngles(0, 1, 2): X-Y-Z factorisation, first angle folded into [0, pi]
void euler_angles_012(const double *R, double *out) {
    auto at = [&](int r, int c) { return R[3 * r + c]; };
    double r0 = std::atan2(at(1, 2), at(2, 2)), r1;
    const double c2 = std::sqrt(at(0, 0) * at(0, 0) + at(0, 1) * at(0, 1));
    if (r0 > 0.0) {
        r0 -= M_PI;
        r1 = std::atan2(-at(0, 2), -c2);
    } else {
        r1 = std::atan2(-at(0, 2), c2);
    }
    const double s1 = std::sin(r0), c1 = std::cos(r0);
    const double r2 = std::atan2(s1 * at(2, 0) - c1 * at(1, 0), c1 * at(1, 1) - s1 * at(2, 1));
    out[0] = -r0;
    out[1] = -r1;
    out[2] = -r2;
}

void mat3_mul(const double *A, const double *B, double *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}

bool inverse6_host(const double *A, double *inv) {
    for (int c = 0; c < 6; ++c) {
        double e[6], x[6];
        for (int i = 0; i < 6; ++i) e[i] = (i == c) ? 1.0 : 0.0;
        if (!solve6_host(A, e, x)) return false;
        for (int i = 0; i < 6; ++i) inv[6 * i + c] = x[i];
    }
    return true;
}

void mat6_mul(const double *A, const double *B, double *C) {
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            double s = 0.0;
            for (int k = 0; k < 6; ++k) s += A[6 * i + k] * B[6 * k + j];
            C[6 * i + j] = s;
        }
}

}  // namespace

// ICPMatcher::estimateCensi (src/icp.cpp:167-397): only acts if the match converged; evaluated at
// Matcher::result over the clouds and correspondences of the last align().
int IcpHandle::info_censi(double *info_out) {
    for (int i = 0; i < 36; ++i) info_out[i] = (i % 7 == 0) ? 1.0 : 0.0;
    if (!last.converged) return WAVECU_OK;
    const size_t ns = result_n_src;
    if (ns == 0) {
        for (int i = 0; i < 36; ++i) info_out[i] = std::nan("");
        return WAVECU_OK;
    }
    CensiArgs ca;
    CensiConsts &k = ca.c;
    {
        double L[9], Rp[9], eul[3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) L[3 * r + c] = result[4 * r + c];
        rotation_from_sigma(L, Rp);  // Eigen Transform::rotation(): the polar factor of the linear part
        euler_angles_012(Rp, eul);
        const double cr = std::cos(eul[0]), sr = std::sin(eul[0]), cp = std::cos(eul[1]), sp = std::sin(eul[1]),
                     cy = std::cos(eul[2]), sy = std::sin(eul[2]);
        // value, first and second derivative of each elementary rotation; R = Rz Ry Rx
        const double X[3][9] = {{1, 0, 0, 0, cr, -sr, 0, sr, cr}, {0, 0, 0, 0, -sr, -cr, 0, cr, -sr}, {0, 0, 0, 0, -cr, sr, 0, -sr, -cr}};
        const double Y[3][9] = {{cp, 0, sp, 0, 1, 0, -sp, 0, cp}, {-sp, 0, cp, 0, 0, 0, -cp, 0, -sp}, {-cp, 0, -sp, 0, 0, 0, sp, 0, -cp}};
        const double Z[3][9] = {{cy, -sy, 0, sy, cy, 0, 0, 0, 1}, {-sy, -cy, 0, cy, -sy, 0, 0, 0, 0}, {-cy, sy, 0, -sy, -cy, 0, 0, 0, 0}};
        auto prod = [&](int dz, int dy, int dx, double *out) {
            double t[9];
            mat3_mul(Z[dz], Y[dy], t);
            mat3_mul(t, X[dx], out);
        };
        prod(0, 0, 0, k.R);
        prod(0, 0, 1, k.dR[0]);
        prod(0, 1, 0, k.dR[1]);
        prod(1, 0, 0, k.dR[2]);
        prod(0, 0, 2, k.ddR[0]);
        prod(0, 1, 1, k.ddR[1]);
        prod(1, 0, 1, k.ddR[2]);
        prod(0, 2, 0, k.ddR[3]);
        prod(1, 1, 0, k.ddR[4]);
        prod(2, 0, 0, k.ddR[5]);
        k.t[0] = result[3];
        k.t[1] = result[7];
        k.t[2] = result[11];
        k.lin = prm.lidar_lin_covar;
        k.ang = prm.lidar_ang_covar;
    }
    const int blocks = (int) std::min<size_t>((ns + kCensiThreads - 1) / kCensiThreads, (size_t) 2 * 148);
    if (blocks > censi_blocks) {
        if (d_censi) WCU_CHECK(cudaFree(d_censi));
        if (h_censi) WCU_CHECK(cudaFreeHost(h_censi));
        d_censi = h_censi = nullptr;
        WCU_CHECK(cudaMalloc((void **) &d_censi, sizeof(double) * kCensiValues * (size_t) blocks));
        WCU_CHECK(cudaHostAlloc((void **) &h_censi, sizeof(double) * kCensiValues * (size_t) blocks, cudaHostAllocDefault));
        censi_blocks = blocks;
    }
    ca.cur = src.d_sorted;
    ca.raw = src.d_raw;
    ca.n_src = (int) ns;
    ca.tgt = target().cloud.d_sorted;
    ca.pos = d_nn_pos;
    ca.partial = d_censi;
    censi_kernel<<<blocks, kCensiThreads, 0, stream>>>(ca);
    WCU_CHECK(cudaGetLastError());
    WCU_CHECK(cudaMemcpyAsync(h_censi, d_censi, sizeof(double) * kCensiValues * (size_t) blocks, cudaMemcpyDeviceToHost, stream));
    WCU_CHECK(cudaStreamSynchronize(stream));
    double v[kCensiValues];
    for (int i = 0; i < kCensiValues; ++i) {
        double s = 0.0;
        for (int b = 0; b < blocks; ++b) s += h_censi[(size_t) b * kCensiValues + i];
        v[i] = s;
    }
    double H[36], M[36];
    for (int i = 0; i < 36; ++i) H[i] = M[i] = 0.0;
    int u = 0;
    for (int i = 0; i < 3; ++i) {
        H[6 * i + i] = 2.0 * v[36];
        for (int q = 0; q < 3; ++q) H[6 * i + 3 + q] = v[u++];
    }
    for (int q = 0; q < 3; ++q)
        for (int l = q; l < 3; ++l) H[6 * (3 + q) + 3 + l] = v[u++];
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) M[6 * i + j] = M[6 * j + i] = v[u++];
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < r; ++c) H[6 * r + c] = H[6 * c + r];
    double Hi[36], A[36], B[36];
    if (!inverse6_host(H, Hi)) {
        for (int i = 0; i < 36; ++i) info_out[i] = std::nan("");
        return WAVECU_OK;
    }
    mat6_mul(Hi, M, A);
    mat6_mul(A, Hi, B);
    if (!inverse6_host(B, info_out))
        for (int i = 0; i < 36; ++i) info_out[i] = std::nan("");
    return WAVECU_OK;
}

void IcpHandle::release() {
    if (d_censi) cudaFree(d_censi);
    if (h_censi) cudaFreeHost(h_censi);
    cudaSetDevice(device);
    src.release();
    tgt.release();
    for (void *p : {(void *) d_nn_pos, (void *) d_nn_idx, (void *) d_nn_d2, (void *) d_out_idx, (void *) d_out_d2,
                    (void *) d_aligned, (void *) d_mc, (void *) d_st, (void *) d_acc, (void *) d_trace,
                    (void *) d_fb_count, (void *) d_fb_list, (void *) d_ticket})
        if (p) cudaFree(p);
    vox.release();
    for (void *p : {(void *) d_orig_src, (void *) d_orig_tgt, (void *) d_pos2})
        if (p) cudaFree(p);
    if (h_acc) cudaFreeHost(h_acc);
    if (h_progress) cudaFreeHost(h_progress);
    if (h_st) cudaFreeHost(h_st);
    if (h_trace) cudaFreeHost(h_trace);
    for (auto e : ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : {ev_main, ev_src_sorted, ev_first, ev_tgt_ready})
        if (e) cudaEventDestroy(e);
    if (aux) cudaStreamDestroy(aux);
    if (copy) cudaStreamDestroy(copy);
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
    return w->h.set_source(xyzw, n, false);
}
int wavecu_icp_set_source_device(wavecu_icp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    return w->h.set_source((const float *) d, n, true);
}
int wavecu_icp_set_target(wavecu_icp *w, const float *xyzw, size_t n) {
    if (!w || (!xyzw && n)) return WAVECU_ERR_ARG;
    return w->h.set_target(xyzw, n, false);
}
int wavecu_icp_set_target_device(wavecu_icp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    return w->h.set_target((const float *) d, n, true);
}
int wavecu_icp_set_target_normals(wavecu_icp *w, const float *nxyzw, size_t n) {
    if (!w || (!nxyzw && n)) return WAVECU_ERR_ARG;
    return w->h.set_target_normals(nxyzw, n, false);
}
int wavecu_icp_set_target_normals_device(wavecu_icp *w, const void *d, size_t n) {
    if (!w || (!d && n)) return WAVECU_ERR_ARG;
    return w->h.set_target_normals((const float *) d, n, true);
}

int wavecu_icp_align(wavecu_icp *w, double T_out[16], int *converged, int *iterations, int *state) {
    if (!w) return WAVECU_ERR_ARG;
    return w->h.align(T_out, converged, iterations, state);
}

int wavecu_icp_match(wavecu_icp *w, double T_out[16], int *converged, int *iterations) {
    if (!w) return WAVECU_ERR_ARG;
    return w->h.match(T_out, converged, iterations);
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
    // no search ran (an empty cloud) -> no correspondences; a search that found fewer than three pairs leaves
    // them in correspondences_ as PCL does (state NO_CORRESPONDENCES, iteration count 0)
    if (ns == 0 || (h.last.iter == 0 && h.last.state != WAVECU_CONV_NO_CORRESPONDENCES)) return WAVECU_OK;
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
    if (!w || !info_out) return WAVECU_ERR_ARG;
    return w->h.info(method, info_out);
}

int wavecu_icp_build_target(wavecu_icp *w) {
    if (!w) return WAVECU_ERR_ARG;
    IcpHandle &h = w->h;
    WCU_CHECK(cudaSetDevice(h.device));
    if (h.tgt_owner) {
        set_last_error("this handle shares another handle's target");
        return WAVECU_ERR_STATE;
    }
    if (h.tgt.dirty) {
        const int rc = h.tgt.build();
        if (rc) return rc;
    }
    if (!h.ev_tgt_ready) WCU_CHECK(cudaEventCreateWithFlags(&h.ev_tgt_ready, cudaEventDisableTiming));
    WCU_CHECK(cudaEventRecord(h.ev_tgt_ready, h.stream));
    ++h.tgt_epoch;
    return WAVECU_OK;
}

int wavecu_icp_share_target(wavecu_icp *w, wavecu_icp *owner) {
    if (!w) return WAVECU_ERR_ARG;
    IcpHandle &h = w->h;
    if (!owner) {
        h.tgt_owner = nullptr;
        h.have_result = false;
        return WAVECU_OK;
    }
    if (owner == w || owner->h.device != h.device || owner->h.tgt_owner || !owner->h.ev_tgt_ready) {
        set_last_error("share_target: the owner must be another handle on the same device whose target was built with "
                       "wavecu_icp_build_target");
        return WAVECU_ERR_ARG;
    }
    h.tgt_owner = &owner->h;
    h.tgt_epoch_seen = 0;
    h.have_result = false;
    return WAVECU_OK;
}

int wavecu_icp_set_search(wavecu_icp *w, int mode) {
    if (!w || (mode != WAVECU_SEARCH_TREE && mode != WAVECU_SEARCH_TILED)) return WAVECU_ERR_ARG;
    const bool tile = mode == WAVECU_SEARCH_TILED;
    if (tile != w->h.use_tile) {
        w->h.use_tile = tile;
        w->h.tgt.want_boxes = tile;
        if (tile) w->h.tgt.dirty = true;   // the box pyramid is built with the tree
    }
    return WAVECU_OK;
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
