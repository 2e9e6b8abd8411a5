// ORACLE - test infrastructure only (see gicp.cpp for what this restates and cites).
#pragma once
#include <cstddef>
#include <vector>

namespace wo {

class KdTree;

struct GicpParams {  // wave::GICPMatcherParams, gicp.hpp:34-38 (the YAML ctor ignores the file's values,
    int corr_rand = 10;   // src/gicp.cpp:8-13, so these defaults are what always runs)
    int max_iter = 100;
    double r_eps = 1e-8;
    double fit_eps = 1e-2;
    float res = 0.1f;
};

struct GicpResult {
    float final_T[16];
    bool converged = false;
    int iterations = 0;
    size_t n_corr = 0;
    long long inner_iterations = 0, evaluations = 0;
    std::vector<double> delta_trace;
};

// computeCovariances: per point the k-NN covariance re-weighted to singular values (1, 1, eps);
// covs: n x 9 doubles (row major 3x3).  False if the cloud has fewer than k points.
bool gicp_covariances(const float *cloud_xyzw, size_t n, const KdTree &tree, int k, double gicp_epsilon,
                      std::vector<double> &covs);

// GeneralizedIterativeClosestPoint::align on the given clouds (voxel filtering, src/gicp.cpp:37-55,
// is applied by the caller with voxel_grid()).
void gicp_align(const float *source, size_t n_src, const float *target, size_t n_tgt, const GicpParams &prm,
                GicpResult &res);

// k-NN PCA normals oriented towards the origin (point-to-plane extension; see gicp.cpp)
void estimate_normals(const float *cloud_xyzw, size_t n, int k, float *normals_xyzw);

}  // namespace wo
