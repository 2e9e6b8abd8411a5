// ORACLE - test infrastructure only (see kdtree.hpp header and matcher.cpp for citations).
#pragma once
#include <cfloat>
#include <cstddef>
#include <vector>

#include "icp.hpp"

namespace wo {

// pcl::VoxelGrid<PointXYZ>::filter; returns false when the grid would overflow int32 (PCL then
// copies the input through unfiltered).
bool voxel_grid(const float *in_xyzw, size_t n, float leaf, std::vector<float> &out_xyzw);

// pcl::transformPointCloud(in, out, Affine3d): double arithmetic per point, cast to float.
void transform_cloud_affine3d(const float *in_xyzw, size_t n, const double *T16, float *out_xyzw);

struct MatcherParams {  // wave::ICPMatcherParams, icp.hpp:35-59
    double max_corr = 3;
    int max_iter = 100;
    double t_eps = 1e-8;
    double fit_eps = 1e-2;
    int multiscale_steps = 3;
    float res = 0.1f;
    int sum_mode = SUM_EXACT;
};

struct MatchResult {
    double T[16];                       // Matcher::result
    int levels = 0, total_iterations = 0;
    IcpResult last;                     // the last align() (correspondences_, final cloud)
    std::vector<float> ds_ref, ds_tgt;  // clouds handed to the last align()
};

// ICPMatcher::match(), src/icp.cpp:75-133
bool icp_match(const float *ref, size_t n_ref, const float *target, size_t n_tgt, const MatcherParams &mp,
               MatchResult &res, int nn_threads = 1);

// ICPMatcher::estimateLUM (src/icp_pcl_functions.cpp:182-289); false = "unsuccessful" (identity).
// k / k_ss: fixed-point shifts of the exact-sum arithmetic (k_quad - 2, k_quad - 4 of
// the align()'s fix_scales); ignored by SUM_PCL.
bool estimate_lum(const float *aligned, const float *target, const int *corr_q, const int *corr_m, size_t n_corr,
                  int sum_mode, int k, int k_ss, double *info36);
// ICPMatcher::estimateLUMold (src/icp_pcl_functions.cpp:51-179)
bool estimate_lum_old(const float *aligned, size_t n_src, const float *target, size_t n_tgt, double max_corr,
                      int sum_mode, int k, int k_ss, double *info36, int nn_threads = 1);

}  // namespace wo
