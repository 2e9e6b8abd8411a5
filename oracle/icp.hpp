// ORACLE - test infrastructure only (see kdtree.hpp header).
//
// CPU restatement of pcl::IterativeClosestPoint<PointXYZ,PointXYZ>::align() as the reference
// drives it from wave_matching/src/icp.cpp:47-50 (parameter plumbing) and :95,116,126 (align),
// following SURVEY.md Appendix A.1-A.5 (PCL 1.8: registration.hpp, icp.hpp,
// correspondence_estimation.hpp, transformation_estimation_svd.hpp,
// default_convergence_criteria.hpp; Eigen 3.3 Umeyama.h).  PCL is not vendored in
// /root/reference and cannot be built here: PARITY UNPINNED at the bit level; the reference's
// own end-to-end tests (tests/icp_tests.cpp, Frobenius < 0.1) are re-stated in tests/.
//
// Two estimator arithmetics are offered:
//   SUM_EXACT   - the repo's estimator spec: every per-pair term is an exact fp64
//                 product of fp32 values, quantised to 2^-k and summed exactly in 128-bit
//                 integers (order independent), then Umeyama / 6x6 LLS in fp64 with a fixed
//                 operation order, result cast to fp32.  The GPU path follows the same spec, so
//                 it can be compared bit for bit.
//   SUM_PCL     - PCL-faithful arithmetic: fp32 sequential sums for the SVD estimator
//                 (Eigen::umeyama with Scalar = float), fp64 sequential sums for point-to-plane.
//                 Used to report the reference's own arithmetic noise floor next to the gap.
// Everything else (fp32 incremental in-place transform, fp32 L2_Simple distances, the
// max-correspondence test in double, the convergence rules and their order, the fp32 4x4
// composition) is identical in both.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "kdtree.hpp"

namespace wo {

enum Estimator { EST_SVD = 0, EST_POINT_TO_PLANE_LLS = 1 };
enum SumMode { SUM_EXACT = 0, SUM_PCL = 1 };
enum ConvState {  // pcl::registration::DefaultConvergenceCriteria::ConvergenceState
    CONV_NOT_CONVERGED = 0,
    CONV_ITERATIONS = 1,
    CONV_TRANSFORM = 2,
    CONV_ABS_MSE = 3,
    CONV_REL_MSE = 4,
    CONV_NO_CORRESPONDENCES = 5
};

struct IcpParams {
    double max_corr = 3;      // icp.hpp:35
    int max_iter = 100;       // icp.hpp:37
    double t_eps = 1e-8;      // icp.hpp:41
    double fit_eps = 1e-2;    // icp.hpp:43
    int estimator = EST_SVD;
    int sum_mode = SUM_EXACT;
};

struct IcpTraceRow {
    double mse;        // mean of the squared correspondence distances seen by this iteration
    int n_corr;
    float T[16];       // incremental fp32 transform estimated in this iteration (row major)
};

struct IcpResult {
    float final_T[16];          // fp32 final_transformation_ (row major)
    bool converged = false;
    int iterations = 0;
    int state = CONV_NOT_CONVERGED;
    std::vector<int> corr_query, corr_match;   // correspondences_ of the last iteration
    std::vector<float> corr_dist;              // squared
    std::vector<float> aligned;                // xyzw, = final_T (x) source, the "final" cloud
    std::vector<IcpTraceRow> trace;
};

struct Correspondences {
    std::vector<int> q, m;
    std::vector<float> d2;
};

// CorrespondenceEstimation::determineCorrespondences (A.3.1): ascending source index, exact 1-NN,
// drop if (double)d2 > max_corr*max_corr, compacted.  `strict` uses d2 < max^2 instead
// (estimateLUMold, src/icp_pcl_functions.cpp:82).
void determine_correspondences(const KdTree &tree, const float *cloud_xyzw, size_t n, double max_corr, bool strict,
                               Correspondences &out, int nthreads = 1);

// target: xyzw; normals: xyzw-stride unit normals of the target (only for EST_POINT_TO_PLANE_LLS)
void icp_align(const float *source, size_t n_src, const float *target, size_t n_tgt, const float *target_normals,
               const IcpParams &prm, IcpResult &res, const KdTree *prebuilt_tree = nullptr, int nn_threads = 1);

// Exponents of the fixed-point spec (shared with the device code through DESIGN.md, not through
// a header): bound M = max|target coord| + min(max_corr, 2 max|source coord| + max|target coord|)
struct FixScales {
    int k_lin, k_quad, k_d2;
};
FixScales fix_scales(const float *source, size_t n_src, const float *target, size_t n_tgt, double max_corr);

}  // namespace wo
