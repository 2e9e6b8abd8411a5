// wavecu_batch_*: a batch of independent ICP alignments spread over the GPUs of one process and, with
// NCCL, over the processes of one job (BASELINE.json configs[4]: 256 scan-to-map
// alignments over 8 GPUs).  Replaces what wave::MultiMatcher does with a host thread pool around PCL
// objects (wave_matching/include/wave/matching/multi_matcher.hpp:30-96,
// impl/multi_matcher_impl.hpp:19-93): here a worker is a host thread driving its own C-ABI handle (own
// streams and device buffers), several workers per device so that one scan's upload and sort overlap
// another's iterations, and a map shared by all scans is uploaded and indexed once per device
// (wavecu_icp_share_target) instead of once per job.
//
// Collectives: none on the data path - scans are independent.  ncclBroadcast moves the map from the
// root rank's device to the others when the caller has it on one rank only, and ONE ncclAllGather at
// the end leaves every rank with every scan's record.  NCCL is loaded with dlopen when a communicator
// is first asked for, so single-process users (and the CPU-only build check) need no libnccl.
#include <dlfcn.h>

#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wavecu.h"
#include "common.cuh"

namespace wavecu {

namespace {

// ---- the few NCCL entry points used, resolved at run time --------------------------------------------
struct Nccl {
    typedef struct ncclComm *comm_t;
    struct unique_id {
        char internal[128];
    };
    int (*GetUniqueId)(unique_id *) = nullptr;
    int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int, comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, comm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    void *lib = nullptr;
    static constexpr int kInt8 = 0;   // ncclInt8 / ncclChar

    bool load() {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            set_last_error(std::string("cannot load libnccl: ") + dlerror());
            return false;
        }
        auto sym = [&](const char *n) { return dlsym(lib, n); };
        GetUniqueId = (decltype(GetUniqueId)) sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank)) sym("ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy)) sym("ncclCommDestroy");
        Broadcast = (decltype(Broadcast)) sym("ncclBroadcast");
        AllGather = (decltype(AllGather)) sym("ncclAllGather");
        GetErrorString = (decltype(GetErrorString)) sym("ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !Broadcast || !AllGather || !GetErrorString) {
            set_last_error("libnccl lacks an expected symbol");
            return false;
        }
        return true;
    }
};

Nccl &nccl() {
    static Nccl n;
    return n;
}

#define WCU_NCCL(expr)                                                                         \
    do {                                                                                       \
        const int rc__ = (expr);                                                               \
        if (rc__ != 0) {                                                                       \
            set_last_error(std::string(#expr) + ": " + nccl().GetErrorString(rc__));          \
            return WAVECU_ERR_NCCL;                                                            \
        }                                                                                      \
    } while (0)

}  // namespace

struct BatchHandle {
    wavecu_icp_params prm;
    std::vector<int> devices;
    int workers_per_device = 4;
    struct Dev {
        int device = 0;
        wavecu_icp *map_owner = nullptr;           // holds the shared map and its search structure
        std::vector<wavecu_icp *> workers;
        cudaStream_t stream = nullptr;             // collectives and map staging
        void *d_map = nullptr;
        size_t map_cap = 0;
    };
    std::vector<Dev> devs;
    bool have_map = false;
    size_t map_n = 0;
    // communicator over the processes of the job: rank = this process, bound to devs[0]
    Nccl::comm_t comm = nullptr;
    int rank = 0, world = 1;
    void *d_gather = nullptr;
    size_t gather_cap = 0;

    int init() {
        for (int d : devices) {
            Dev dv;
            dv.device = d;
            WCU_CHECK(cudaSetDevice(d));
            WCU_CHECK(cudaStreamCreateWithFlags(&dv.stream, cudaStreamNonBlocking));
            int rc = wavecu_icp_create(&prm, d, nullptr, &dv.map_owner);
            if (rc) return rc;
            for (int w = 0; w < workers_per_device; ++w) {
                wavecu_icp *h = nullptr;
                rc = wavecu_icp_create(&prm, d, nullptr, &h);
                if (rc) return rc;
                dv.workers.push_back(h);
            }
            devs.push_back(dv);
        }
        return WAVECU_OK;
    }

    int share_map() {
        for (Dev &dv : devs) {
            int rc = wavecu_icp_build_target(dv.map_owner);
            if (rc) return rc;
            for (wavecu_icp *h : dv.workers) {
                rc = wavecu_icp_share_target(h, dv.map_owner);
                if (rc) return rc;
            }
        }
        have_map = true;
        return WAVECU_OK;
    }

    int set_map(const float *xyzw, size_t n) {
        for (Dev &dv : devs) {
            const int rc = wavecu_icp_set_target(dv.map_owner, xyzw, n);   // one upload + one tree per device
            if (rc) return rc;
        }
        map_n = n;
        return share_map();
    }

    // the map lives on the root rank only: stage it on the root's first device, ncclBroadcast it to every
    // rank's first device, copy on to a rank's further devices
    int broadcast_map(const float *xyzw_on_root, size_t n, int root) {
        if (!comm) {
            set_last_error("broadcast_map needs a communicator (wavecu_batch_init_comm)");
            return WAVECU_ERR_STATE;
        }
        Dev &d0 = devs[0];
        WCU_CHECK(cudaSetDevice(d0.device));
        const size_t bytes = n * 4 * sizeof(float);
        for (Dev &dv : devs)
            if (n > dv.map_cap) {
                WCU_CHECK(cudaSetDevice(dv.device));
                if (dv.d_map) WCU_CHECK(cudaFree(dv.d_map));
                dv.d_map = nullptr;
                WCU_CHECK(cudaMalloc(&dv.d_map, (n + 64) * 4 * sizeof(float)));
                dv.map_cap = n + 64;
            }
        WCU_CHECK(cudaSetDevice(d0.device));
        if (rank == root) {
            if (!xyzw_on_root && n) return WAVECU_ERR_ARG;
            WCU_CHECK(cudaMemcpyAsync(d0.d_map, xyzw_on_root, bytes, cudaMemcpyHostToDevice, d0.stream));
        }
        WCU_NCCL(nccl().Broadcast(d0.d_map, d0.d_map, bytes, Nccl::kInt8, root, comm, d0.stream));
        WCU_CHECK(cudaStreamSynchronize(d0.stream));
        for (size_t i = 0; i < devs.size(); ++i) {
            Dev &dv = devs[i];
            WCU_CHECK(cudaSetDevice(dv.device));
            if (i > 0) WCU_CHECK(cudaMemcpyPeer(dv.d_map, dv.device, d0.d_map, d0.device, bytes));
            const int rc = wavecu_icp_set_target_device(dv.map_owner, dv.d_map, n);
            if (rc) return rc;
        }
        map_n = n;
        return share_map();
    }

    int match(const float *const *scans, const size_t *n_points, const int *scan_ids, int n_scans,
              const float *const *targets, const size_t *n_targets, int with_info, wavecu_batch_record *out) {
        if (!targets && !have_map) {
            set_last_error("batch_match without per-scan targets needs a map (wavecu_batch_set_map)");
            return WAVECU_ERR_STATE;
        }
        std::atomic<int> next{0};
        std::atomic<int> failed{WAVECU_OK};
        std::vector<std::string> errors(devs.size() * (size_t) workers_per_device);
        auto work = [&](Dev *dv, wavecu_icp *h, size_t slot) {
            if (targets) wavecu_icp_share_target(h, nullptr);
            else wavecu_icp_share_target(h, dv->map_owner);
            for (;;) {
                const int i = next.fetch_add(1);
                if (i >= n_scans || failed.load() != WAVECU_OK) break;
                wavecu_batch_record &r = out[i];
                std::memset(&r, 0, sizeof r);
                r.scan_id = scan_ids ? scan_ids[i] : i;
                r.device = dv->device;
                // MultiMatcher::spin: setRef, setTarget, match, estimateInfo (impl/multi_matcher_impl.hpp:45-48)
                int rc = wavecu_icp_set_source(h, scans[i], n_points[i]);
                if (!rc && targets) rc = wavecu_icp_set_target(h, targets[i], n_targets[i]);
                if (!rc) rc = wavecu_icp_match(h, r.T, &r.converged, &r.iterations);
                for (int k = 0; k < 36; ++k) r.info[k] = (k % 7 == 0) ? 1.0 : 0.0;   // Matcher::estimateInfo default
                // ICPMatcher::estimateInfo falls through to estimateLUMold whatever the setting (src/icp.cpp:135-142)
                if (!rc && with_info) rc = wavecu_icp_info(h, WAVECU_INFO_LUMOLD, r.info);
                if (rc) {
                    errors[slot] = wavecu_last_error();
                    failed.store(rc);
                    break;
                }
            }
        };
        std::vector<std::thread> pool;
        size_t slot = 0;
        for (Dev &dv : devs)
            for (wavecu_icp *h : dv.workers) pool.emplace_back(work, &dv, h, slot++);
        for (std::thread &t : pool) t.join();
        if (failed.load() != WAVECU_OK) {
            for (const std::string &e : errors)
                if (!e.empty()) {
                    set_last_error(e);
                    break;
                }
            return failed.load();
        }
        return WAVECU_OK;
    }

    // every rank contributes its n_local records (the same count on every rank; unused slots carry scan_id
    // -1); `all` receives world * n_local records in rank order
    int allgather(const wavecu_batch_record *local, int n_local, wavecu_batch_record *all) {
        const size_t bytes = sizeof(wavecu_batch_record) * (size_t) n_local;
        if (world == 1) {
            std::memcpy(all, local, bytes);
            return WAVECU_OK;
        }
        if (!comm) {
            set_last_error("allgather over several ranks needs a communicator (wavecu_batch_init_comm)");
            return WAVECU_ERR_STATE;
        }
        Dev &d0 = devs[0];
        WCU_CHECK(cudaSetDevice(d0.device));
        const size_t need = bytes * (size_t) (world + 1);
        if (need > gather_cap) {
            if (d_gather) WCU_CHECK(cudaFree(d_gather));
            d_gather = nullptr;
            WCU_CHECK(cudaMalloc(&d_gather, need));
            gather_cap = need;
        }
        char *send = (char *) d_gather, *recv = send + bytes;
        WCU_CHECK(cudaMemcpyAsync(send, local, bytes, cudaMemcpyHostToDevice, d0.stream));
        WCU_NCCL(nccl().AllGather(send, recv, bytes, Nccl::kInt8, comm, d0.stream));
        WCU_CHECK(cudaMemcpyAsync(all, recv, bytes * (size_t) world, cudaMemcpyDeviceToHost, d0.stream));
        WCU_CHECK(cudaStreamSynchronize(d0.stream));
        return WAVECU_OK;
    }

    void release() {
        if (comm) nccl().CommDestroy(comm);
        comm = nullptr;
        for (Dev &dv : devs) {
            cudaSetDevice(dv.device);
            for (wavecu_icp *h : dv.workers) wavecu_icp_destroy(h);
            if (dv.map_owner) wavecu_icp_destroy(dv.map_owner);
            if (dv.d_map) cudaFree(dv.d_map);
            if (dv.stream) cudaStreamDestroy(dv.stream);
        }
        if (d_gather && !devs.empty()) {
            cudaSetDevice(devs[0].device);
            cudaFree(d_gather);
        }
        devs.clear();
    }
};

}  // namespace wavecu

using namespace wavecu;

struct wavecu_batch {
    BatchHandle h;
};

extern "C" {

int wavecu_batch_create(const wavecu_icp_params *params, const int *devices, int n_devices, int workers_per_device,
                        wavecu_batch **out) {
    if (!out || (n_devices > 0 && !devices) || n_devices < 0) return WAVECU_ERR_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        set_last_error("no CUDA device (libwavecu has no CPU fallback)");
        return WAVECU_ERR_CUDA;
    }
    wavecu_batch *b = new wavecu_batch();
    if (params) b->h.prm = *params;
    else wavecu_icp_default_params(&b->h.prm);
    if (n_devices == 0)
        for (int d = 0; d < count; ++d) b->h.devices.push_back(d);   // every visible device
    else
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= count) {
                delete b;
                set_last_error("no such CUDA device");
                return WAVECU_ERR_CUDA;
            }
            b->h.devices.push_back(devices[i]);
        }
    b->h.workers_per_device = workers_per_device > 0 ? workers_per_device : 4;
    const int rc = b->h.init();
    if (rc) {
        b->h.release();
        delete b;
        return rc;
    }
    *out = b;
    return WAVECU_OK;
}

int wavecu_batch_destroy(wavecu_batch *b) {
    if (!b) return WAVECU_OK;
    b->h.release();
    delete b;
    return WAVECU_OK;
}

int wavecu_batch_set_map(wavecu_batch *b, const float *xyzw, size_t n) {
    if (!b || (!xyzw && n)) return WAVECU_ERR_ARG;
    return b->h.set_map(xyzw, n);
}

int wavecu_batch_unique_id(void *id128) {
    if (!id128) return WAVECU_ERR_ARG;
    if (!nccl().load()) return WAVECU_ERR_NCCL;
    Nccl::unique_id id;
    WCU_NCCL(nccl().GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return WAVECU_OK;
}

int wavecu_batch_init_comm(wavecu_batch *b, const void *id128, int rank, int world) {
    if (!b || !id128 || rank < 0 || world < 1 || rank >= world) return WAVECU_ERR_ARG;
    if (!nccl().load()) return WAVECU_ERR_NCCL;
    Nccl::unique_id id;
    std::memcpy(&id, id128, sizeof id);
    WCU_CHECK(cudaSetDevice(b->h.devs[0].device));
    if (b->h.comm) nccl().CommDestroy(b->h.comm);
    b->h.comm = nullptr;
    WCU_NCCL(nccl().CommInitRank(&b->h.comm, world, id, rank));
    b->h.rank = rank;
    b->h.world = world;
    return WAVECU_OK;
}

int wavecu_batch_broadcast_map(wavecu_batch *b, const float *xyzw_on_root, size_t n, int root) {
    if (!b || root < 0 || root >= b->h.world) return WAVECU_ERR_ARG;
    return b->h.broadcast_map(xyzw_on_root, n, root);
}

int wavecu_batch_match(wavecu_batch *b, const float *const *scans, const size_t *n_points, const int *scan_ids,
                       int n_scans, const float *const *targets, const size_t *n_targets, int with_info,
                       wavecu_batch_record *out) {
    if (!b || n_scans < 0 || (n_scans && (!scans || !n_points || !out)) || (targets && !n_targets)) return WAVECU_ERR_ARG;
    return b->h.match(scans, n_points, scan_ids, n_scans, targets, n_targets, with_info, out);
}

int wavecu_batch_allgather(wavecu_batch *b, const wavecu_batch_record *local, int n_local, wavecu_batch_record *all) {
    if (!b || n_local < 0 || (n_local && (!local || !all))) return WAVECU_ERR_ARG;
    return b->h.allgather(local, n_local, all);
}

int wavecu_batch_device_count(wavecu_batch *b) { return b ? (int) b->h.devs.size() : 0; }

}  // extern "C"
