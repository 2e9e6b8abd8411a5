// ORACLE - test infrastructure only (see ndt.cpp for what this restates and cites).
#pragma once
#include <cstddef>
#include <vector>

namespace wo {

class KdTree;

struct NdtParams {  // wave::NDTMatcherParams, ndt.hpp:37-41
    int step_size = 3;
    int max_iter = 100;
    double t_eps = 1e-8;
    float res = 5;
    int line_search = 1;  // 1: More-Thuente search runs (PCL >= 1.9); 0: PCL 1.8's skipped search
};

struct NdtLeaf {
    int voxel, n;
    float centroid[3];   // fp32 centroid: what the radius search runs on
    double mean[3];      // fp64 mean: what the score uses
    double icov[9];
};

struct NdtGrid {  // pcl::VoxelGridCovariance with min_points_per_voxel_ = 6, eigenvalue floor 0.01
    std::vector<NdtLeaf> leaves;      // ascending voxel index
    std::vector<float> centroids;     // xyzw
    void build(const float *target_xyzw, size_t n, float res);
};

struct NdtResult {
    float final_T[16];
    double pose[6];
    bool converged = false;
    int iterations = 0;
    int n_voxels = 0;
    double score = 0;
    std::vector<double> step_trace, score_trace;
};

double ndt_derivatives(const NdtGrid &grid, const KdTree &tree, const float *source, const float *trans, size_t n,
                       const double p[6], float res, double d1, double d2, bool with_hessian, double g[6],
                       double H[36]);

void ndt_align(const float *source, size_t n_src, const float *target, size_t n_tgt, const NdtParams &prm,
               NdtResult &res);

}  // namespace wo
