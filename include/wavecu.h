/* wavecu.h - the C ABI of the B200 registration hot path (libwavecu.so).
 *
 * This is the drop-in boundary for wave_matching's ICPMatcher / GICPMatcher / NDTMatcher::match():
 * plain pointers and sizes, no C++/torch types.  The C++ shim under include/wave/matching/ (same
 * class names and semantics as the reference headers) and the Python ctypes mirror in
 * libwave_b200/matching.py both sit on top of exactly these entry points.
 *
 * What each group replaces in the reference (paths relative to wave_matching/):
 *   wavecu_icp_create/destroy      pcl::IterativeClosestPoint member + parameter plumbing,
 *                                  include/wave/matching/icp.hpp:100, src/icp.cpp:32-51
 *   wavecu_icp_set_source/target   ICPMatcher::setRef/setTarget + icp.setInputSource/Target,
 *                                  src/icp.cpp:67-73,124-125
 *   wavecu_icp_match               icp.align + hasConverged + getFinalTransformation,
 *                                  src/icp.cpp:126-128 (and the per-level calls :95-101,:116-119)
 *   wavecu_icp_correspondences     icp.correspondences_, src/icp.cpp:213,
 *                                  src/icp_pcl_functions.cpp:191
 *   wavecu_icp_aligned             the "final" cloud written by align(), icp.hpp:109
 *   wavecu_icp_info                ICPMatcher::estimateLUM / estimateCensi / estimateLUMold,
 *                                  src/icp_pcl_functions.cpp:51-289, src/icp.cpp:167-397
 *   wavecu_nn_*                    pcl::KdTreeFLANN::setInputCloud / nearestKSearch(k=1),
 *                                  src/icp_pcl_functions.cpp:67-80
 *   wavecu_voxel_grid              pcl::VoxelGrid::filter, src/icp.cpp:81-90,106-113,
 *                                  src/gicp.cpp:39-40,49-50
 *   wavecu_gicp_*                  pcl::GeneralizedIterativeClosestPoint member + plumbing,
 *                                  include/wave/matching/gicp.hpp:61, src/gicp.cpp:20-64
 *   wavecu_ndt_*                   pcl::NormalDistributionsTransform member + plumbing,
 *                                  include/wave/matching/ndt.hpp:72, src/ndt.cpp:18-65
 *
 * Conventions: clouds are arrays of pcl::PointXYZ records = 4 floats (x, y, z, pad) = "xyzw";
 * 4x4 transforms are row-major doubles; every function returns 0 on success and a negative
 * wavecu_status on failure (wavecu_last_error() gives the text).  Non-convergence is not an
 * error: it is reported through the `converged` out-flag, as match() returns false
 * (src/icp.cpp:132).  There is no CPU fallback: without a CUDA device every call fails with
 * WAVECU_ERR_CUDA.  A handle is bound to one device and one stream and must not be used from two
 * host threads at once; distinct handles are independent (MultiMatcher gives each worker its own,
 * impl/multi_matcher_impl.hpp:22).
 */
#ifndef WAVECU_H
#define WAVECU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    WAVECU_OK = 0,
    WAVECU_ERR_ARG = -1,
    WAVECU_ERR_CUDA = -2,
    WAVECU_ERR_STATE = -3,
    WAVECU_ERR_NCCL = -4
} wavecu_status;

enum { WAVECU_EST_SVD = 0, WAVECU_EST_POINT_TO_PLANE = 1 };
enum { WAVECU_INFO_LUM = 0, WAVECU_INFO_CENSI = 1, WAVECU_INFO_LUMOLD = 2 };
/* pcl::registration::DefaultConvergenceCriteria::ConvergenceState */
enum {
    WAVECU_CONV_NOT_CONVERGED = 0,
    WAVECU_CONV_ITERATIONS = 1,
    WAVECU_CONV_TRANSFORM = 2,
    WAVECU_CONV_ABS_MSE = 3,
    WAVECU_CONV_REL_MSE = 4,
    WAVECU_CONV_NO_CORRESPONDENCES = 5
};

/* Field-for-field mirror of wave::ICPMatcherParams (icp.hpp:30-65) plus the estimator switch. */
typedef struct {
    double max_corr;        /* icp.hpp:35, default 3 */
    int max_iter;           /* icp.hpp:37, default 100 */
    double t_eps;           /* icp.hpp:41, default 1e-8 */
    double fit_eps;         /* icp.hpp:43, default 1e-2 */
    double lidar_ang_covar; /* icp.hpp:46, default 7.78e-9 */
    double lidar_lin_covar; /* icp.hpp:49, default 2.5e-4 */
    int multiscale_steps;   /* icp.hpp:54, default 3 */
    float res;              /* icp.hpp:59, default 0.1 */
    int covar_estimator;    /* icp.hpp:60-64, default LUM */
    int estimator;          /* WAVECU_EST_*; the reference is always SVD (icp.hpp:100) */
} wavecu_icp_params;

void wavecu_icp_default_params(wavecu_icp_params *p);

typedef struct wavecu_icp wavecu_icp;

/* device: CUDA ordinal.  stream: a cudaStream_t to launch on, or NULL for a stream owned by the
 * handle. */
int wavecu_icp_create(const wavecu_icp_params *params, int device, void *stream, wavecu_icp **out);
int wavecu_icp_destroy(wavecu_icp *h);
int wavecu_icp_set_params(wavecu_icp *h, const wavecu_icp_params *params);

/* Host clouds, copied to the device.  The copy is asynchronous: it overlaps the sort / tree build of
 * the cloud set before it, and at full resolution (res <= 0) the build of this cloud is queued
 * right behind it.  From pageable memory the CUDA runtime stages the data before the call
 * returns; a caller that passes page-locked memory must leave the buffer untouched until the next
 * wavecu_icp_match / wavecu_icp_align on this handle has returned (ICPMatcher holds the caller's
 * cloud pointers over exactly that span, src/icp.cpp:67-73).  A page-locked source is in fact
 * copied behind the target set after it: the first iteration waits for the target's tree, not for
 * the source. */
int wavecu_icp_set_source(wavecu_icp *h, const float *xyzw, size_t n);
int wavecu_icp_set_target(wavecu_icp *h, const float *xyzw, size_t n);
/* Unit normals of the target, same order and stride as the target (point-to-plane only). */
int wavecu_icp_set_target_normals(wavecu_icp *h, const float *nxyzw, size_t n);
/* Device-resident clouds (device pointers on the handle's device; copied device-to-device). */
int wavecu_icp_set_source_device(wavecu_icp *h, const void *d_xyzw, size_t n);
int wavecu_icp_set_target_device(wavecu_icp *h, const void *d_xyzw, size_t n);
int wavecu_icp_set_target_normals_device(wavecu_icp *h, const void *d_nxyzw, size_t n);

/* Scan-to-map batches: many source scans against one target ("map").  The map is uploaded to one
 * handle, its search structure built once (wavecu_icp_build_target), and any number of other handles on
 * the same device then match against it read-only (wavecu_icp_share_target; owner = NULL detaches) -
 * instead of every MultiMatcher worker uploading and indexing the same cloud for every job
 * (impl/multi_matcher_impl.hpp:45-46 calls setTarget per job).  The owner must outlive its sharers and
 * must not change its target while they are matching.  Full resolution (res <= 0) only. */
int wavecu_icp_build_target(wavecu_icp *h);
int wavecu_icp_share_target(wavecu_icp *h, wavecu_icp *owner);

/* Which kernel runs the per-iteration exact 1-NN search (results are identical bit for bit):
 * WAVECU_SEARCH_TREE - one query per lane walking the LBVH (default); WAVECU_SEARCH_TILED - a CTA per
 * tile of Morton-consecutive queries, target runs staged into shared memory with cp.async.bulk, the
 * tree walk only for the queries the tile cannot certify (csrc/tile_nn.cuh).  The environment variable
 * WAVECU_NN=tile|walk sets the default of new handles. */
enum { WAVECU_SEARCH_TREE = 0, WAVECU_SEARCH_TILED = 1 };
int wavecu_icp_set_search(wavecu_icp *h, int mode);

/* One pcl align(): (re)builds the target search structure if the target changed, iterates on
 * the device, returns final_transformation_ (fp32 values widened to double, row major). */
int wavecu_icp_align(wavecu_icp *h, double T_out[16], int *converged, int *iterations, int *state);
/* ICPMatcher::match() (src/icp.cpp:75-133): branches on params.res / multiscale_steps, voxel
 * filtering and level composition included. */
int wavecu_icp_match(wavecu_icp *h, double T_out[16], int *converged, int *iterations);

/* correspondences_ of the last align() in ascending source index; buffers need room for the
 * source size; *n receives the count. */
int wavecu_icp_correspondences(wavecu_icp *h, int *idx_query, int *idx_match, float *dist2, size_t *n);
/* final_transformation_ applied to the (possibly down-sampled) source; *n receives its size.
 * xyzw may be NULL to query the size only. */
int wavecu_icp_aligned(wavecu_icp *h, float *xyzw, size_t *n);
/* Per-iteration trace of the last align(): mse[i], n_corr[i], T_inc[16*i..] (fp32 incremental
 * transform, row major).  Arrays need room for max_iter entries; *n receives the count kept. */
int wavecu_icp_trace(wavecu_icp *h, double *mse, int *n_corr, float *T_inc, int *n);
int wavecu_icp_info(wavecu_icp *h, int method, double info_out[36]);

/* Timing and launch counters of the last align()/match(), measured with CUDA events on the
 * handle's stream when profiling is enabled (adds one event pair per kernel). */
typedef struct {
    double build_ms;        /* search-structure build (sort + tree) */
    double iterate_ms;      /* sum over iterations of the correspondence kernel (transform + 1-NN) */
    double solve_ms;        /* sum of the reduction + estimator/convergence kernels */
    double total_ms;        /* whole align()/match() on the stream */
    long long iterate_launches;
    long long kernel_launches; /* every kernel this library launched in the call */
    long long pairs;        /* sum over iterations of source points queried */
    long long fallback_queries; /* of which the tiled kernel could not certify and finished with the tree walk */
} wavecu_stats;
int wavecu_icp_set_profiling(wavecu_icp *h, int enabled);
int wavecu_icp_stats(wavecu_icp *h, wavecu_stats *out);

/* Stand-alone exact 1-NN (the correspondence kernel without the estimator). */
typedef struct wavecu_nn wavecu_nn;
int wavecu_nn_create(int device, void *stream, wavecu_nn **out);
int wavecu_nn_destroy(wavecu_nn *h);
int wavecu_nn_set_target(wavecu_nn *h, const float *xyzw, size_t n);
/* idx[i] = index of the nearest target point (lowest index among exact fp32 ties), -1 if none
 * within max_dist (max_dist <= 0: unlimited); dist2[i] = fp32 squared distance. */
int wavecu_nn_search(wavecu_nn *h, const float *q_xyzw, size_t nq, double max_dist, int *idx, float *dist2);
/* Same on device-resident queries/outputs; elapsed_ms (may be NULL) receives the device time of
 * `repeats` back-to-back searches measured with CUDA events on the handle's stream. */
int wavecu_nn_search_device(wavecu_nn *h, const void *d_q_xyzw, size_t nq, double max_dist, void *d_idx,
                            void *d_dist2, int repeats, float *elapsed_ms);

/* ---- NDTMatcher (include/wave/matching/ndt.hpp:33-80, src/ndt.cpp:18-65) ---------------------------
 * Field-for-field mirror of wave::NDTMatcherParams (ndt.hpp:37-41); res is clamped to >= 0.05 as the
 * reference constructor does (src/ndt.cpp:23-26). */
typedef struct {
    int step_size;   /* ndt.hpp:37 (an int in the reference), default 3 */
    int max_iter;    /* ndt.hpp:38, default 100 */
    double t_eps;    /* ndt.hpp:39, default 1e-8 */
    float res;       /* ndt.hpp:40, default 5 */
    /* Not a reference field: which PCL release's computeStepLengthMT the match follows.  PCL 1.8
     * initialises `interval_converged = (step_max - step_min) > 0`, which skips the More-Thuente
     * search whenever step_size > t_eps / 2 (always for the reference's values) and leaves a capped
     * Newton step; PCL >= 1.9 tests `< 0` and searches.  Only the latter passes the reference's own
     * smallDisplacement test (tests/ndt_tests.cpp:85-102), so it is the default. */
    int line_search; /* WAVECU_NDT_LS_*, default WAVECU_NDT_LS_MORE_THUENTE */
} wavecu_ndt_params;
enum { WAVECU_NDT_LS_PCL18 = 0, WAVECU_NDT_LS_MORE_THUENTE = 1 };

void wavecu_ndt_default_params(wavecu_ndt_params *p);

typedef struct wavecu_ndt wavecu_ndt;
int wavecu_ndt_create(const wavecu_ndt_params *params, int device, void *stream, wavecu_ndt **out);
int wavecu_ndt_destroy(wavecu_ndt *h);
int wavecu_ndt_set_params(wavecu_ndt *h, const wavecu_ndt_params *params);
/* ndt.setInputSource / setInputTarget (src/ndt.cpp:48-56); the target call marks the voxel grid stale */
int wavecu_ndt_set_source(wavecu_ndt *h, const float *xyzw, size_t n);
int wavecu_ndt_set_target(wavecu_ndt *h, const float *xyzw, size_t n);
int wavecu_ndt_set_source_device(wavecu_ndt *h, const void *d_xyzw, size_t n);
int wavecu_ndt_set_target_device(wavecu_ndt *h, const void *d_xyzw, size_t n);
/* ndt.align + hasConverged + getFinalTransformation (src/ndt.cpp:58-65) */
int wavecu_ndt_match(wavecu_ndt *h, double T_out[16], int *converged, int *iterations);
/* The normal-distribution cells of the target (>= 6 points, ascending voxel index): for parity tests
 * and callers that want the map.  Arrays may be NULL; at most `capacity` cells are written. */
int wavecu_ndt_grid(wavecu_ndt *h, int *n_cells, int *voxel, int *count, float *centroid3, double *mean3,
                    double *icov9, int capacity);
/* One computeDerivatives pass at pose6 = (x, y, z, roll, pitch, yaw) with the source transformed by the
 * fp32 matrix T16: score, gradient (6) and Hessian (6x6, row major). */
int wavecu_ndt_derivatives(wavecu_ndt *h, const double pose6[6], const float T16[16], double *score, double g6[6],
                           double H36[36]);
/* n_cells: the normal-distribution cells (>= 6 points, usable covariance) - what wavecu_ndt_grid lists */
int wavecu_ndt_stats(wavecu_ndt *h, long long *kernel_launches, long long *derivative_passes, int *n_cells);
/* with profiling on, every derivative pass of the following matches is bracketed by CUDA events on the
 * handle's stream; wavecu_ndt_timing returns their summed device time (of the last match) */
int wavecu_ndt_set_profiling(wavecu_ndt *h, int enabled);
int wavecu_ndt_timing(wavecu_ndt *h, double *derivative_kernel_ms);

/* ---- GICPMatcher (include/wave/matching/gicp.hpp:30-65, src/gicp.cpp:20-64) -------------------------
 * Field-for-field mirror of wave::GICPMatcherParams (gicp.hpp:34-38). */
typedef struct {
    int corr_rand;   /* gicp.hpp:34, default 10 (k of the covariance neighbourhoods) */
    int max_iter;    /* gicp.hpp:35, default 100 */
    double r_eps;    /* gicp.hpp:36, default 1e-8 (rotation epsilon) */
    double fit_eps;  /* gicp.hpp:37, default 1e-2 (passed to PCL but unused by GICP) */
    float res;       /* gicp.hpp:38, default 0.1: voxel filter applied in setRef/setTarget; <= 0 none */
} wavecu_gicp_params;

void wavecu_gicp_default_params(wavecu_gicp_params *p);

typedef struct wavecu_gicp wavecu_gicp;
int wavecu_gicp_create(const wavecu_gicp_params *params, int device, void *stream, wavecu_gicp **out);
int wavecu_gicp_destroy(wavecu_gicp *h);
int wavecu_gicp_set_params(wavecu_gicp *h, const wavecu_gicp_params *params);
/* GICPMatcher::setRef / setTarget (src/gicp.cpp:37-55): voxel-filter when res > 0, then
 * gicp.setInputSource / setInputTarget */
int wavecu_gicp_set_source(wavecu_gicp *h, const float *xyzw, size_t n);
int wavecu_gicp_set_target(wavecu_gicp *h, const float *xyzw, size_t n);
int wavecu_gicp_set_source_device(wavecu_gicp *h, const void *d_xyzw, size_t n);
int wavecu_gicp_set_target_device(wavecu_gicp *h, const void *d_xyzw, size_t n);
/* gicp.align + hasConverged + getFinalTransformation (src/gicp.cpp:57-64) */
int wavecu_gicp_match(wavecu_gicp *h, double T_out[16], int *converged, int *iterations);
/* Per-point covariances (which = 0 source, 1 target) in the order of the (filtered) cloud, 9 doubles
 * each; covs9 may be NULL to query the size.  wavecu_gicp_cloud returns that (filtered) cloud. */
int wavecu_gicp_covariances(wavecu_gicp *h, int which, double *covs9, size_t *n);
int wavecu_gicp_cloud(wavecu_gicp *h, int which, float *xyzw, size_t *n);
int wavecu_gicp_stats(wavecu_gicp *h, long long *kernel_launches, long long *evaluations,
                      long long *inner_iterations, size_t *n_corr);
/* with profiling on, every cost / gradient evaluation kernel of the following matches is bracketed by CUDA
 * events on the handle's stream; wavecu_gicp_timing returns their summed device time and count */
int wavecu_gicp_set_profiling(wavecu_gicp *h, int enabled);
int wavecu_gicp_timing(wavecu_gicp *h, double *cost_kernel_ms, long long *cost_kernel_launches);

/* pcl::VoxelGrid<pcl::PointXYZ>::filter (src/icp.cpp:81-90,106-113; src/gicp.cpp:39-40,49-50):
 * one centroid per occupied voxel in ascending voxel index; out_xyzw needs room for n points.
 * *filtered = 0 when the grid would overflow int32 and the input is passed through unchanged. */
int wavecu_voxel_grid(int device, const float *xyzw, size_t n, float leaf, float *out_xyzw, size_t *n_out,
                      int *filtered);

/* ---- batches of independent alignments (wave::MultiMatcher, multi_matcher.hpp:30-96) -------------------
 * One batch object spreads scans over the GPUs of this process (devices[]; n_devices = 0: every visible
 * device), workers_per_device concurrent matches per GPU, each on its own handle and streams.  With a map
 * (wavecu_batch_set_map, or wavecu_batch_broadcast_map when only one rank of a multi-process job holds it)
 * every scan is matched against the same target, uploaded and indexed once per GPU.  Multi-process jobs
 * (one process per GPU, BASELINE.json configs[4]) create a communicator from an id made on one rank and
 * handed to the others by the caller's own means (a file, MPI, torch.distributed ...); the only
 * collectives are the optional map broadcast and ONE all-gather of the result records at the end. */
typedef struct {
    double T[16];      /* Matcher::result, row major */
    double info[36];   /* Matcher::information (estimateInfo: ends as estimateLUMold, src/icp.cpp:135-142) */
    int converged;     /* match() return value */
    int iterations;
    int scan_id;       /* the caller's id of this scan (multi_matcher.hpp:57); -1: unused slot */
    int device;        /* CUDA ordinal that matched it */
} wavecu_batch_record;

typedef struct wavecu_batch wavecu_batch;
int wavecu_batch_create(const wavecu_icp_params *params, const int *devices, int n_devices, int workers_per_device,
                        wavecu_batch **out);
int wavecu_batch_destroy(wavecu_batch *b);
int wavecu_batch_device_count(wavecu_batch *b);
int wavecu_batch_set_map(wavecu_batch *b, const float *xyzw, size_t n);
/* scans[i]: host cloud of n_points[i] points; scan_ids may be NULL (ids 0..n-1); targets NULL: match against
 * the map, else targets[i] / n_targets[i] per scan (MultiMatcher::insert(id, src, target)); with_info != 0
 * also fills the information matrix; out[i] receives the record of scans[i].  Blocks until all are done. */
int wavecu_batch_match(wavecu_batch *b, const float *const *scans, const size_t *n_points, const int *scan_ids,
                       int n_scans, const float *const *targets, const size_t *n_targets, int with_info,
                       wavecu_batch_record *out);
/* multi-process: 128-byte NCCL unique id (make it on one rank), communicator on the batch's first device */
int wavecu_batch_unique_id(void *id128);
int wavecu_batch_init_comm(wavecu_batch *b, const void *id128, int rank, int world);
int wavecu_batch_broadcast_map(wavecu_batch *b, const float *xyzw_on_root, size_t n, int root);
/* every rank passes n_local records (same count everywhere); all receives world * n_local, in rank order */
int wavecu_batch_allgather(wavecu_batch *b, const wavecu_batch_record *local, int n_local, wavecu_batch_record *all);

const char *wavecu_last_error(void);
int wavecu_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* WAVECU_H */
