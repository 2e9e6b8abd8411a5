// ORACLE - test infrastructure only.  Flat C entry points for ctypes (oracle/oracle.py).
#include <cstring>
#include <thread>

#include "icp.hpp"
#include "smallmat.hpp"

using namespace wo;

extern "C" {

void *wo_kdtree_create(const float *xyzw, size_t n) { return new KdTree(xyzw, n, 4); }
void wo_kdtree_destroy(void *t) { delete static_cast<KdTree *>(t); }

void wo_kdtree_nn1(void *t, const float *q, size_t nq, int *idx, float *d2, int nthreads) {
    const KdTree &tree = *static_cast<KdTree *>(t);
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) tree.nn1(q + 4 * i, idx + i, d2 + i);
    };
    if (nthreads <= 1) {
        work(0, nq);
        return;
    }
    std::vector<std::thread> th;
    const size_t chunk = (nq + nthreads - 1) / nthreads;
    for (int k = 0; k < nthreads; ++k) {
        const size_t lo = k * chunk, hi = std::min(nq, lo + chunk);
        if (lo < hi) th.emplace_back(work, lo, hi);
    }
    for (auto &x : th) x.join();
}

// k-NN for nq queries; idx/d2 are nq*k, padded with -1 / inf
void wo_kdtree_knn(void *t, const float *q, size_t nq, int k, int *idx, float *d2) {
    const KdTree &tree = *static_cast<KdTree *>(t);
    for (size_t i = 0; i < nq; ++i) {
        const int found = tree.knn(q + 4 * i, k, idx + i * k, d2 + i * k);
        for (int j = found; j < k; ++j) {
            idx[i * k + j] = -1;
            d2[i * k + j] = std::numeric_limits<float>::infinity();
        }
    }
}

// O(n*m) reference for the reference: lowest index among exact fp32 ties
void wo_brute_nn1(const float *tgt, size_t nt, const float *q, size_t nq, int *idx, float *d2) {
    for (size_t i = 0; i < nq; ++i) {
        float best = std::numeric_limits<float>::infinity();
        int bi = -1;
        for (size_t j = 0; j < nt; ++j) {
            const float *p = tgt + 4 * j;
            if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
            const float d = l2_simple(q + 4 * i, p);
            if (d < best) {
                best = d;
                bi = (int) j;
            }
        }
        idx[i] = bi;
        d2[i] = best;
    }
}

struct wo_icp_params_c {
    double max_corr;
    int max_iter;
    double t_eps;
    double fit_eps;
    int estimator;
    int sum_mode;
};

void *wo_icp_run(const float *src, size_t n_src, const float *tgt, size_t n_tgt, const float *normals,
                 const wo_icp_params_c *p, void *prebuilt_tree, int nn_threads) {
    IcpParams prm;
    prm.max_corr = p->max_corr;
    prm.max_iter = p->max_iter;
    prm.t_eps = p->t_eps;
    prm.fit_eps = p->fit_eps;
    prm.estimator = p->estimator;
    prm.sum_mode = p->sum_mode;
    IcpResult *r = new IcpResult();
    icp_align(src, n_src, tgt, n_tgt, normals, prm, *r, static_cast<KdTree *>(prebuilt_tree), nn_threads);
    return r;
}

void wo_icp_result_summary(void *h, float *T16, int *converged, int *iters, int *state, size_t *n_corr,
                           size_t *n_trace, size_t *n_aligned) {
    const IcpResult &r = *static_cast<IcpResult *>(h);
    std::memcpy(T16, r.final_T, sizeof r.final_T);
    *converged = r.converged;
    *iters = r.iterations;
    *state = r.state;
    *n_corr = r.corr_query.size();
    *n_trace = r.trace.size();
    *n_aligned = r.aligned.size() / 4;
}

void wo_icp_result_corr(void *h, int *q, int *m, float *d2) {
    const IcpResult &r = *static_cast<IcpResult *>(h);
    std::memcpy(q, r.corr_query.data(), r.corr_query.size() * sizeof(int));
    std::memcpy(m, r.corr_match.data(), r.corr_match.size() * sizeof(int));
    std::memcpy(d2, r.corr_dist.data(), r.corr_dist.size() * sizeof(float));
}

void wo_icp_result_aligned(void *h, float *xyzw) {
    const IcpResult &r = *static_cast<IcpResult *>(h);
    std::memcpy(xyzw, r.aligned.data(), r.aligned.size() * sizeof(float));
}

void wo_icp_result_trace(void *h, double *mse, int *n_corr, float *T16s) {
    const IcpResult &r = *static_cast<IcpResult *>(h);
    for (size_t i = 0; i < r.trace.size(); ++i) {
        mse[i] = r.trace[i].mse;
        n_corr[i] = r.trace[i].n_corr;
        std::memcpy(T16s + 16 * i, r.trace[i].T, 16 * sizeof(float));
    }
}

void wo_icp_result_free(void *h) { delete static_cast<IcpResult *>(h); }

void wo_fix_scales(const float *src, size_t n_src, const float *tgt, size_t n_tgt, double max_corr, int *k3) {
    const FixScales s = fix_scales(src, n_src, tgt, n_tgt, max_corr);
    k3[0] = s.k_lin;
    k3[1] = s.k_quad;
    k3[2] = s.k_d2;
}

void wo_rotation_from_sigma(const double *S9, double *R9) { rotation_from_sigma(S9, R9); }
// sum of llrint(terms[i] * 2^k) in 128-bit fixed point, converted back (Fix128::value)
double wo_fix128_sum(const double *terms, size_t n, int k) {
    Fix128 f;
    for (size_t i = 0; i < n; ++i) f.add(terms[i], k);
    return f.value(k);
}
int wo_solve6(const double *A36, const double *b6, double *x6) { return solve_pp<6>(A36, b6, x6) ? 1 : 0; }

}  // extern "C"

// ---- matcher level ------------------------------------------------------------------------------
#include "matcher.hpp"

extern "C" {

// returns the number of output points; out must have room for n points
size_t wo_voxel_grid(const float *in, size_t n, float leaf, float *out, int *filtered) {
    std::vector<float> o;
    const bool ok = voxel_grid(in, n, leaf, o);
    if (filtered) *filtered = ok ? 1 : 0;
    std::memcpy(out, o.data(), o.size() * sizeof(float));
    return o.size() / 4;
}

void wo_transform_affine3d(const float *in, size_t n, const double *T16, float *out) {
    transform_cloud_affine3d(in, n, T16, out);
}

struct wo_matcher_params_c {
    double max_corr;
    int max_iter;
    double t_eps;
    double fit_eps;
    int multiscale_steps;
    float res;
    int sum_mode;
};

void *wo_icp_match(const float *ref, size_t n_ref, const float *tgt, size_t n_tgt, const wo_matcher_params_c *p,
                   int nn_threads, int *success) {
    MatcherParams mp;
    mp.max_corr = p->max_corr;
    mp.max_iter = p->max_iter;
    mp.t_eps = p->t_eps;
    mp.fit_eps = p->fit_eps;
    mp.multiscale_steps = p->multiscale_steps;
    mp.res = p->res;
    mp.sum_mode = p->sum_mode;
    MatchResult *r = new MatchResult();
    *success = icp_match(ref, n_ref, tgt, n_tgt, mp, *r, nn_threads) ? 1 : 0;
    return r;
}

void wo_match_summary(void *h, double *T16, int *levels, int *total_iters, size_t *n_ds_ref, size_t *n_ds_tgt) {
    const MatchResult &r = *static_cast<MatchResult *>(h);
    std::memcpy(T16, r.T, sizeof r.T);
    *levels = r.levels;
    *total_iters = r.total_iterations;
    *n_ds_ref = r.ds_ref.size() / 4;
    *n_ds_tgt = r.ds_tgt.size() / 4;
}
void wo_match_clouds(void *h, float *ds_ref, float *ds_tgt) {
    const MatchResult &r = *static_cast<MatchResult *>(h);
    std::memcpy(ds_ref, r.ds_ref.data(), r.ds_ref.size() * sizeof(float));
    std::memcpy(ds_tgt, r.ds_tgt.data(), r.ds_tgt.size() * sizeof(float));
}
// the IcpResult of the last align(): usable with the wo_icp_result_* getters (do not free it)
void *wo_match_last(void *h) { return &static_cast<MatchResult *>(h)->last; }
void wo_match_free(void *h) { delete static_cast<MatchResult *>(h); }

int wo_estimate_lum(const float *aligned, const float *target, const int *q, const int *m, size_t n, int sum_mode,
                    int k, int k_ss, double *info36) {
    return estimate_lum(aligned, target, q, m, n, sum_mode, k, k_ss, info36) ? 1 : 0;
}
int wo_estimate_lum_old(const float *aligned, size_t n_src, const float *target, size_t n_tgt, double max_corr,
                        int sum_mode, int k, int k_ss, double *info36, int nn_threads) {
    return estimate_lum_old(aligned, n_src, target, n_tgt, max_corr, sum_mode, k, k_ss, info36, nn_threads) ? 1 : 0;
}

}  // extern "C"

// ---- NDT ------------------------------------------------------------------------------------------
#include "ndt.hpp"

extern "C" {

struct wo_ndt_params_c {
    int step_size;
    int max_iter;
    double t_eps;
    float res;
    int line_search;
};

// out: T16 (fp32 final transform), pose6, converged, iterations, n_voxels, score; trace arrays need
// room for max_iter + 2 entries
void wo_ndt_align(const float *src, size_t n_src, const float *tgt, size_t n_tgt, const wo_ndt_params_c *p, float *T16,
                  double *pose6, int *converged, int *iterations, int *n_voxels, double *score, double *step_trace,
                  int *n_trace) {
    NdtParams prm;
    prm.step_size = p->step_size;
    prm.max_iter = p->max_iter;
    prm.t_eps = p->t_eps;
    prm.res = p->res;
    prm.line_search = p->line_search;
    NdtResult r;
    ndt_align(src, n_src, tgt, n_tgt, prm, r);
    std::memcpy(T16, r.final_T, sizeof r.final_T);
    std::memcpy(pose6, r.pose, sizeof r.pose);
    *converged = r.converged;
    *iterations = r.iterations;
    *n_voxels = r.n_voxels;
    *score = r.score;
    *n_trace = (int) r.step_trace.size();
    for (size_t i = 0; i < r.step_trace.size(); ++i) step_trace[i] = r.step_trace[i];
}

// voxel statistics of the target grid: returns the number of leaves; arrays sized for n points
size_t wo_ndt_grid(const float *tgt, size_t n, float res, int *voxel, int *count, float *centroid3, double *mean3,
                   double *icov9) {
    NdtGrid g;
    g.build(tgt, n, res);
    for (size_t i = 0; i < g.leaves.size(); ++i) {
        voxel[i] = g.leaves[i].voxel;
        count[i] = g.leaves[i].n;
        for (int d = 0; d < 3; ++d) {
            centroid3[3 * i + d] = g.leaves[i].centroid[d];
            mean3[3 * i + d] = g.leaves[i].mean[d];
        }
        for (int d = 0; d < 9; ++d) icov9[9 * i + d] = g.leaves[i].icov[d];
    }
    return g.leaves.size();
}

// one computeDerivatives pass at pose p (source transformed by the fp32 pose matrix)
double wo_ndt_derivatives(const float *src, size_t n_src, const float *tgt, size_t n_tgt, float res, const double *pose6,
                          const float *T16, double *g6, double *H36) {
    NdtGrid grid;
    grid.build(tgt, n_tgt, res);
    KdTree tree(grid.centroids.data(), grid.leaves.size(), 4);
    std::vector<float> trans(4 * n_src);
    for (size_t i = 0; i < n_src; ++i) {
        const float x = src[4 * i], y = src[4 * i + 1], z = src[4 * i + 2];
        trans[4 * i + 0] = T16[0] * x + T16[1] * y + T16[2] * z + T16[3];
        trans[4 * i + 1] = T16[4] * x + T16[5] * y + T16[6] * z + T16[7];
        trans[4 * i + 2] = T16[8] * x + T16[9] * y + T16[10] * z + T16[11];
        trans[4 * i + 3] = 1.f;
    }
    const double o = 0.55, c1 = 10 * (1 - o), c2 = o / std::pow((double) res, 3), d3 = -std::log(c2);
    const double d1 = -std::log(c1 + c2) - d3;
    const double d2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / d1);
    return ndt_derivatives(grid, tree, src, trans.data(), n_src, pose6, res, d1, d2, true, g6, H36);
}

}  // extern "C"

// ---- GICP -----------------------------------------------------------------------------------------
#include "gicp.hpp"

extern "C" {

void wo_gicp_covariances(const float *cloud, size_t n, int k, double eps, double *covs9) {
    KdTree tree(cloud, n, 4);
    std::vector<double> c;
    if (!gicp_covariances(cloud, n, tree, k, eps, c)) c.assign(9 * n, 0.0);
    std::memcpy(covs9, c.data(), c.size() * sizeof(double));
}

void wo_gicp_align(const float *src, size_t n_src, const float *tgt, size_t n_tgt, int corr_rand, int max_iter,
                   double r_eps, float *T16, int *converged, int *iterations, size_t *n_corr, long long *inner,
                   long long *evals, double *delta_trace, int *n_trace) {
    GicpParams prm;
    prm.corr_rand = corr_rand;
    prm.max_iter = max_iter;
    prm.r_eps = r_eps;
    GicpResult r;
    gicp_align(src, n_src, tgt, n_tgt, prm, r);
    std::memcpy(T16, r.final_T, sizeof r.final_T);
    *converged = r.converged;
    *iterations = r.iterations;
    *n_corr = r.n_corr;
    *inner = r.inner_iterations;
    *evals = r.evaluations;
    *n_trace = (int) r.delta_trace.size();
    for (size_t i = 0; i < r.delta_trace.size(); ++i) delta_trace[i] = r.delta_trace[i];
}

}  // extern "C"

extern "C" void wo_estimate_normals(const float *cloud, size_t n, int k, float *normals_xyzw) {
    estimate_normals(cloud, n, k, normals_xyzw);
}
