// ORACLE - test infrastructure only (see kdtree.hpp header).
//
// CPU restatement of the first-party arithmetic and plumbing of wave::ICPMatcher:
//   * ICPMatcher::match()          wave_matching/src/icp.cpp:75-133 (resolution / multiscale
//                                  branches, level composition order)
//   * pcl::VoxelGrid<PointXYZ>     SURVEY.md Appendix A.6 (PCL 1.8 filters/impl/voxel_grid.hpp);
//                                  used at src/icp.cpp:81-90,106-113 and src/gicp.cpp:39-40,49-50
//   * pcl::transformPointCloud     Appendix A.6 (common/impl/transforms.hpp), src/icp.cpp:84-86
//   * estimateLUM / estimateLUMold wave_matching/src/icp_pcl_functions.cpp:182-289 / 51-179 -
//                                  first-party code, restated expression by expression (float
//                                  sub-expressions, double accumulators, float `ss`)
//   * estimateCensi                wave_matching/src/icp.cpp:167-397
// Deviation (documented): std::sort in VoxelGrid is unstable, so the order in which
// PCL sums the points of one voxel is libstdc++-specific; this oracle (and the GPU path) sum each
// voxel in ascending cloud index, which can differ from PCL in the last ulp of a centroid.
#include "matcher.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "smallmat.hpp"

namespace wo {

bool voxel_grid(const float *in, size_t n, float leaf, std::vector<float> &out) {
    out.clear();
    const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_.array()
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool any = false;
    for (size_t i = 0; i < n; ++i) {
        const float *p = in + 4 * i;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
        any = true;
        for (int d = 0; d < 3; ++d) {
            mn[d] = std::min(mn[d], p[d]);
            mx[d] = std::max(mx[d], p[d]);
        }
    }
    if (!any) return true;
    const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
    const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
    const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
    if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
        out.assign(in, in + 4 * n);  // "Leaf size is too small": output = input
        return false;
    }
    int min_b[3], max_b[3], div_b[3];
    for (int d = 0; d < 3; ++d) {
        min_b[d] = static_cast<int>(std::floor(mn[d] * inv));
        max_b[d] = static_cast<int>(std::floor(mx[d] * inv));
        div_b[d] = max_b[d] - min_b[d] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    std::vector<std::pair<unsigned, unsigned>> iv;  // (voxel idx, cloud index)
    iv.reserve(n);
    for (size_t i = 0; i < n; ++i) {
        const float *p = in + 4 * i;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
        const int ijk0 = static_cast<int>(std::floor(p[0] * inv) - static_cast<float>(min_b[0]));
        const int ijk1 = static_cast<int>(std::floor(p[1] * inv) - static_cast<float>(min_b[1]));
        const int ijk2 = static_cast<int>(std::floor(p[2] * inv) - static_cast<float>(min_b[2]));
        const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
        iv.emplace_back(static_cast<unsigned>(idx), static_cast<unsigned>(i));
    }
    std::stable_sort(iv.begin(), iv.end(),
                     [](const std::pair<unsigned, unsigned> &a, const std::pair<unsigned, unsigned> &b) {
                         return a.first < b.first;
                     });
    size_t index = 0;
    while (index < iv.size()) {
        size_t j = index + 1;
        while (j < iv.size() && iv[j].first == iv[index].first) ++j;
        float s[3] = {0.f, 0.f, 0.f};  // CentroidPoint / AccumulatorXYZ: Vector3f sums
        for (size_t k = index; k < j; ++k)
            for (int d = 0; d < 3; ++d) s[d] += in[4 * (size_t) iv[k].second + d];
        const float cnt = static_cast<float>(j - index);
        out.push_back(s[0] / cnt);
        out.push_back(s[1] / cnt);
        out.push_back(s[2] / cnt);
        out.push_back(1.0f);
        index = j;
    }
    return true;
}

void transform_cloud_affine3d(const float *in, size_t n, const double *T, float *out) {
    for (size_t i = 0; i < n; ++i) {
        const float *p = in + 4 * i;
        const double x = p[0], y = p[1], z = p[2];
        float *o = out + 4 * i;
        const float ox = static_cast<float>(T[0] * x + T[1] * y + T[2] * z + T[3]);
        const float oy = static_cast<float>(T[4] * x + T[5] * y + T[6] * z + T[7]);
        const float oz = static_cast<float>(T[8] * x + T[9] * y + T[10] * z + T[11]);
        o[0] = ox;
        o[1] = oy;
        o[2] = oz;
        o[3] = p[3];
    }
}

static void matmul4d(const double *A, const double *B, double *C) {  // Eigen column-combination order
    double out[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double r = A[4 * i + 0] * B[0 * 4 + j];
            r = A[4 * i + 1] * B[1 * 4 + j] + r;
            r = A[4 * i + 2] * B[2 * 4 + j] + r;
            r = A[4 * i + 3] * B[3 * 4 + j] + r;
            out[4 * i + j] = r;
        }
    std::memcpy(C, out, sizeof out);
}

bool icp_match(const float *ref, size_t n_ref, const float *target, size_t n_tgt, const MatcherParams &mp,
               MatchResult &res, int nn_threads) {
    res = MatchResult();
    for (int i = 0; i < 16; ++i) res.T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    IcpParams ip;
    ip.max_corr = mp.max_corr;
    ip.max_iter = mp.max_iter;
    ip.t_eps = mp.t_eps;
    ip.fit_eps = mp.fit_eps;
    ip.estimator = EST_SVD;
    ip.sum_mode = mp.sum_mode;
    res.levels = 0;
    res.total_iterations = 0;
    if (mp.res > 0) {
        if (mp.multiscale_steps > 0) {
            double running[16];
            for (int i = 0; i < 16; ++i) running[i] = (i % 5 == 0) ? 1.0 : 0.0;
            for (int i = mp.multiscale_steps; i >= 0; --i) {
                const float leaf_size = std::pow(2, i) * mp.res;
                voxel_grid(ref, n_ref, leaf_size, res.ds_ref);
                std::vector<float> moved(res.ds_ref.size());
                transform_cloud_affine3d(res.ds_ref.data(), res.ds_ref.size() / 4, running, moved.data());
                res.ds_ref.swap(moved);
                voxel_grid(target, n_tgt, leaf_size, res.ds_tgt);
                ip.max_corr = std::pow(2, i) * mp.max_corr;
                icp_align(res.ds_ref.data(), res.ds_ref.size() / 4, res.ds_tgt.data(), res.ds_tgt.size() / 4, nullptr,
                          ip, res.last, nullptr, nn_threads);
                ++res.levels;
                res.total_iterations += res.last.iterations;
                if (!res.last.converged) return false;
                double F[16];
                for (int k = 0; k < 16; ++k) F[k] = (double) res.last.final_T[k];
                matmul4d(F, running, running);
            }
            std::memcpy(res.T, running, sizeof running);
            return true;
        }
        voxel_grid(ref, n_ref, mp.res, res.ds_ref);
        voxel_grid(target, n_tgt, mp.res, res.ds_tgt);
        icp_align(res.ds_ref.data(), res.ds_ref.size() / 4, res.ds_tgt.data(), res.ds_tgt.size() / 4, nullptr, ip,
                  res.last, nullptr, nn_threads);
        res.levels = 1;
        res.total_iterations = res.last.iterations;
        if (res.last.converged) {
            for (int k = 0; k < 16; ++k) res.T[k] = (double) res.last.final_T[k];
            return true;
        }
        return false;
    }
    res.ds_ref.assign(ref, ref + 4 * n_ref);
    res.ds_tgt.assign(target, target + 4 * n_tgt);
    icp_align(ref, n_ref, target, n_tgt, nullptr, ip, res.last, nullptr, nn_threads);
    res.levels = 1;
    res.total_iterations = res.last.iterations;
    if (res.last.converged) {
        for (int k = 0; k < 16; ++k) res.T[k] = (double) res.last.final_T[k];
        return true;
    }
    return false;
}

// ---- Lu & Milios information matrix ---------------------------------------------------------------
namespace {

struct LumPair {
    float aver[3], diff[3];
};

// the per-pair sub-expressions exactly as the reference writes them (float arithmetic)
inline LumPair lum_pair(const float *a /*aligned source*/, const float *b /*target*/) {
    LumPair p;
    for (int d = 0; d < 3; ++d) {
        p.aver[d] = 0.5f * (a[d] + b[d]);
        p.diff[d] = a[d] - b[d];
    }
    return p;
}

bool lum_from_pairs(const std::vector<LumPair> &pr, int sum_mode, bool old_variant, int k, int k_ss, double *info) {
    const int numCorr = (int) pr.size();
    double MM[36], MZ[6];
    for (int i = 0; i < 36; ++i) MM[i] = 0;
    for (int i = 0; i < 6; ++i) MZ[i] = 0;
    auto M = [&](int r, int c) -> double & { return MM[6 * r + c]; };
    if (sum_mode == SUM_PCL) {
        for (int ci = 0; ci != numCorr; ++ci) {
            const float *av = pr[ci].aver, *df = pr[ci].diff;
            M(0, 4) -= av[1];
            M(0, 5) += av[2];
            M(1, 3) -= av[2];
            M(1, 4) += av[0];
            M(2, 3) += av[1];
            M(2, 5) -= av[0];
            M(3, 4) -= av[0] * av[2];
            M(3, 5) -= av[0] * av[1];
            M(4, 5) -= av[1] * av[2];
            M(3, 3) += av[1] * av[1] + av[2] * av[2];
            M(4, 4) += av[0] * av[0] + av[1] * av[1];
            M(5, 5) += av[0] * av[0] + av[2] * av[2];
            MZ[0] += df[0];
            MZ[1] += df[1];
            MZ[2] += df[2];
            MZ[3] += av[1] * df[2] - av[2] * df[1];
            MZ[4] += av[0] * df[1] - av[1] * df[0];
            MZ[5] += av[2] * df[0] - av[0] * df[2];
        }
    } else {
        // the repo's estimator spec: the same float terms, summed exactly in 2^-k fixed point
        Fix128 a[3], aa[6], dz[6];
        for (int ci = 0; ci != numCorr; ++ci) {
            const float *av = pr[ci].aver, *df = pr[ci].diff;
            a[0].add(av[0], k);
            a[1].add(av[1], k);
            a[2].add(av[2], k);
            aa[0].add(av[0] * av[2], k);
            aa[1].add(av[0] * av[1], k);
            aa[2].add(av[1] * av[2], k);
            aa[3].add(av[1] * av[1] + av[2] * av[2], k);
            aa[4].add(av[0] * av[0] + av[1] * av[1], k);
            aa[5].add(av[0] * av[0] + av[2] * av[2], k);
            dz[0].add(df[0], k);
            dz[1].add(df[1], k);
            dz[2].add(df[2], k);
            dz[3].add(av[1] * df[2] - av[2] * df[1], k);
            dz[4].add(av[0] * df[1] - av[1] * df[0], k);
            dz[5].add(av[2] * df[0] - av[0] * df[2], k);
        }
        M(0, 4) = -a[1].value(k);
        M(0, 5) = a[2].value(k);
        M(1, 3) = -a[2].value(k);
        M(1, 4) = a[0].value(k);
        M(2, 3) = a[1].value(k);
        M(2, 5) = -a[0].value(k);
        M(3, 4) = -aa[0].value(k);
        M(3, 5) = -aa[1].value(k);
        M(4, 5) = -aa[2].value(k);
        M(3, 3) = aa[3].value(k);
        M(4, 4) = aa[4].value(k);
        M(5, 5) = aa[5].value(k);
        for (int i = 0; i < 6; ++i) MZ[i] = dz[i].value(k);
    }
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

    double D[6], inv[36];
    bool ok = inverse_pp<6>(MM, inv);
    for (int r = 0; r < 6; ++r) {
        double s = 0;
        for (int c = 0; c < 6; ++c) s += inv[6 * r + c] * MZ[c];
        D[r] = ok ? s : std::numeric_limits<double>::quiet_NaN();
    }
    float ss = 0.0f;
    if (sum_mode == SUM_PCL) {
        for (int ci = 0; ci != numCorr; ++ci) {
            const float *av = pr[ci].aver, *df = pr[ci].diff;
            ss += static_cast<float>(std::pow(df[0] - (D[0] + av[2] * D[5] - av[1] * D[4]), 2.0f) +
                                     std::pow(df[1] - (D[1] + av[0] * D[4] - av[2] * D[3]), 2.0f) +
                                     std::pow(df[2] - (D[2] + av[1] * D[3] - av[0] * D[5]), 2.0f));
        }
    } else {
        Fix128 acc;
        for (int ci = 0; ci != numCorr; ++ci) {
            const float *av = pr[ci].aver, *df = pr[ci].diff;
            const double e0 = df[0] - (D[0] + av[2] * D[5] - av[1] * D[4]);
            const double e1 = df[1] - (D[1] + av[0] * D[4] - av[2] * D[3]);
            const double e2 = df[2] - (D[2] + av[1] * D[3] - av[0] * D[5]);
            const float term = static_cast<float>(e0 * e0 + e1 * e1 + e2 * e2);
            if (std::isfinite(term)) acc.add((double) term, k_ss);
            else ss = std::numeric_limits<float>::quiet_NaN();
        }
        if (!std::isnan(ss)) ss = static_cast<float>(acc.value(k_ss));
    }
    bool failed = false;
    if (ss < 0.0000000000001 || !std::isfinite(ss)) {
        failed = true;
        for (int i = 0; i < 36; ++i) info[i] = (i % 7 == 0) ? 1.0 : 0.0;
        if (!old_variant) return false;  // estimateLUM returns here; estimateLUMold falls through
    }
    const float rec = 1.0f / ss;
    for (int i = 0; i < 36; ++i) info[i] = MM[i] * rec;
    return !failed;
}

}  // namespace

bool estimate_lum(const float *aligned, const float *target, const int *corr_q, const int *corr_m, size_t n_corr,
                  int sum_mode, int k, int k_ss, double *info) {
    std::vector<LumPair> pr;
    pr.reserve(n_corr);
    for (size_t i = 0; i < n_corr; ++i) {
        if (corr_m[i] > -1) pr.push_back(lum_pair(aligned + 4 * (size_t) corr_q[i], target + 4 * (size_t) corr_m[i]));
    }
    return lum_from_pairs(pr, sum_mode, false, k, k_ss, info);
}

bool estimate_lum_old(const float *aligned, size_t n_src, const float *target, size_t n_tgt, double max_corr,
                      int sum_mode, int k, int k_ss, double *info, int nn_threads) {
    KdTree tree(target, n_tgt, 4);
    Correspondences c;
    determine_correspondences(tree, aligned, n_src, max_corr, true /* d2 < max^2 */, c, nn_threads);
    std::vector<LumPair> pr;
    pr.reserve(c.q.size());
    for (size_t i = 0; i < c.q.size(); ++i)
        pr.push_back(lum_pair(aligned + 4 * (size_t) c.q[i], target + 4 * (size_t) c.m[i]));
    return lum_from_pairs(pr, sum_mode, true, k, k_ss, info);
}

}  // namespace wo
