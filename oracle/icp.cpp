// ORACLE - test infrastructure only.  See icp.hpp for what this restates and what it cites.
#include "icp.hpp"

#include <cfloat>
#include <cmath>
#include <cstring>
#include <thread>

#include "smallmat.hpp"

namespace wo {

void determine_correspondences(const KdTree &tree, const float *cloud, size_t n, double max_corr, bool strict,
                               Correspondences &out, int nthreads) {
    const double max_dist_sqr = max_corr * max_corr;  // correspondence_estimation.hpp: double
    std::vector<int> idx(n);
    std::vector<float> d2(n);
    auto work = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            const float *p = cloud + 4 * i;
            if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) {
                idx[i] = -1;
                continue;
            }
            tree.nn1(p, &idx[i], &d2[i]);
        }
    };
    if (nthreads <= 1 || n < 4096) {
        work(0, n);
    } else {
        std::vector<std::thread> th;
        const size_t chunk = (n + nthreads - 1) / nthreads;
        for (int t = 0; t < nthreads; ++t) {
            const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
            if (lo < hi) th.emplace_back(work, lo, hi);
        }
        for (auto &t : th) t.join();
    }
    out.q.clear();
    out.m.clear();
    out.d2.clear();
    for (size_t i = 0; i < n; ++i) {
        if (idx[i] < 0) continue;
        const double d = (double) d2[i];
        if (strict ? !(d < max_dist_sqr) : (d > max_dist_sqr)) continue;
        out.q.push_back((int) i);
        out.m.push_back(idx[i]);
        out.d2.push_back(d2[i]);
    }
}

static double max_abs_coord(const float *c, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            const float v = c[4 * i + d];
            if (std::isfinite(v)) m = std::max(m, (double) std::fabs(v));
        }
    return m;
}

FixScales fix_scales(const float *source, size_t n_src, const float *target, size_t n_tgt, double max_corr) {
    const double ms = max_abs_coord(source, n_src), mt = max_abs_coord(target, n_tgt);
    const double reach = std::min(max_corr, 2.0 * ms + mt);
    const double M = std::max(mt + reach, 1.0);
    const int E = pow2_exponent(M);
    const double dmax = std::max(std::min(max_corr * max_corr, 12.0 * M * M), 1e-30);
    FixScales s;
    s.k_lin = 50 - E;
    s.k_quad = 50 - 2 * E;
    s.k_d2 = 50 - pow2_exponent(dmax);
    return s;
}

namespace {

void set_identity4f(float *T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}

// ---- TransformationEstimationSVD -> Eigen::umeyama, Scalar = float (A.4), sequential fp32 ----
bool estimate_svd_pcl(const float *cur, const float *tgt, const Correspondences &c, float *T) {
    const size_t n = c.q.size();
    const float one_over_n = 1.0f / static_cast<float>(n);
    float sm[3] = {0, 0, 0}, dm[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            sm[a] += cur[4 * (size_t) c.q[i] + a];
            dm[a] += tgt[4 * (size_t) c.m[i] + a];
        }
    for (int a = 0; a < 3; ++a) {
        sm[a] = sm[a] * one_over_n;
        dm[a] = dm[a] * one_over_n;
    }
    float sig[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < n; ++i) {
        float sd[3], dd[3];
        for (int a = 0; a < 3; ++a) {
            sd[a] = cur[4 * (size_t) c.q[i] + a] - sm[a];
            dd[a] = tgt[4 * (size_t) c.m[i] + a] - dm[a];
        }
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) sig[3 * a + b] += dd[a] * sd[b];
    }
    double S[9], R[9];
    for (int i = 0; i < 9; ++i) S[i] = (double) (one_over_n * sig[i]);
    rotation_from_sigma(S, R);
    float Rf[9];
    for (int i = 0; i < 9; ++i) Rf[i] = (float) R[i];
    set_identity4f(T);
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) T[4 * a + b] = Rf[3 * a + b];
        float r = Rf[3 * a + 0] * sm[0];
        r = Rf[3 * a + 1] * sm[1] + r;
        r = Rf[3 * a + 2] * sm[2] + r;
        T[4 * a + 3] = dm[a] - r;
    }
    return true;
}

// ---- the repo's estimator spec: exact fixed-point sums + fp64 Umeyama ----
bool estimate_svd_exact(const float *cur, const float *tgt, const Correspondences &c, const FixScales &fs,
                        float *T, double *mse) {
    const size_t n = c.q.size();
    Fix128 Sp[3], Sq[3], Sqp[9], Sd;
    for (size_t i = 0; i < n; ++i) {
        const float *p = cur + 4 * (size_t) c.q[i];
        const float *q = tgt + 4 * (size_t) c.m[i];
        for (int a = 0; a < 3; ++a) {
            Sp[a].add((double) p[a], fs.k_lin);
            Sq[a].add((double) q[a], fs.k_lin);
            for (int b = 0; b < 3; ++b) Sqp[3 * a + b].add((double) q[a] * (double) p[b], fs.k_quad);
        }
        Sd.add((double) c.d2[i], fs.k_d2);
    }
    const double dn = (double) n;
    double sp[3], sq[3], mp[3], mq[3], S[9], R[9];
    for (int a = 0; a < 3; ++a) {
        sp[a] = Sp[a].value(fs.k_lin);
        sq[a] = Sq[a].value(fs.k_lin);
        mp[a] = sp[a] / dn;
        mq[a] = sq[a] / dn;
    }
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S[3 * a + b] = (Sqp[3 * a + b].value(fs.k_quad) - (sq[a] * sp[b]) / dn) / dn;
    rotation_from_sigma(S, R);
    set_identity4f(T);
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) T[4 * a + b] = (float) R[3 * a + b];
        const double r = (R[3 * a + 0] * mp[0] + R[3 * a + 1] * mp[1]) + R[3 * a + 2] * mp[2];
        T[4 * a + 3] = (float) (mq[a] - r);
    }
    *mse = Sd.value(fs.k_d2) / dn;
    return true;
}

// TransformationEstimationPointToPlaneLLS::constructTransformationMatrix
void construct_lls_transform(const double x[6], float *T) {
    const double alpha = x[0], beta = x[1], gamma = x[2];
    set_identity4f(T);
    T[0] = static_cast<float>(std::cos(gamma) * std::cos(beta));
    T[1] = static_cast<float>(-std::sin(gamma) * std::cos(alpha) + std::cos(gamma) * std::sin(beta) * std::sin(alpha));
    T[2] = static_cast<float>(std::sin(gamma) * std::sin(alpha) + std::cos(gamma) * std::sin(beta) * std::cos(alpha));
    T[4] = static_cast<float>(std::sin(gamma) * std::cos(beta));
    T[5] = static_cast<float>(std::cos(gamma) * std::cos(alpha) + std::sin(gamma) * std::sin(beta) * std::sin(alpha));
    T[6] = static_cast<float>(-std::cos(gamma) * std::sin(alpha) + std::sin(gamma) * std::sin(beta) * std::cos(alpha));
    T[8] = static_cast<float>(-std::sin(beta));
    T[9] = static_cast<float>(std::cos(beta) * std::sin(alpha));
    T[10] = static_cast<float>(std::cos(beta) * std::cos(alpha));
    T[3] = static_cast<float>(x[3]);
    T[7] = static_cast<float>(x[4]);
    T[11] = static_cast<float>(x[5]);
}

// per-pair point-to-plane row in PCL's arithmetic: a,b,c,d evaluated in fp32, widened to double
struct PlaneRow {
    double J[6];  // a b c nx ny nz
    double d;
};
inline PlaneRow plane_row(const float *s, const float *t, const float *nrm) {
    const float sx = s[0], sy = s[1], sz = s[2];
    const float dx = t[0], dy = t[1], dz = t[2];
    const float nx = nrm[0], ny = nrm[1], nz = nrm[2];
    PlaneRow r;
    r.J[0] = nz * sy - ny * sz;
    r.J[1] = nx * sz - nz * sx;
    r.J[2] = ny * sx - nx * sy;
    r.J[3] = nx;
    r.J[4] = ny;
    r.J[5] = nz;
    r.d = nx * dx + ny * dy + nz * dz - nx * sx - ny * sy - nz * sz;
    return r;
}

bool estimate_plane(const float *cur, const float *tgt, const float *nrm, const Correspondences &c,
                    const FixScales &fs, int sum_mode, float *T, double *mse_exact) {
    const size_t n = c.q.size();
    double ATA[36], ATb[6];
    if (sum_mode == SUM_EXACT) {
        const int k = fs.k_quad - 3;
        Fix128 A[21], B[6], Sd;
        for (size_t i = 0; i < n; ++i) {
            const float *nn = nrm + 4 * (size_t) c.m[i];
            Sd.add((double) c.d2[i], fs.k_d2);
            if (!(std::isfinite(nn[0]) && std::isfinite(nn[1]) && std::isfinite(nn[2]))) continue;
            const PlaneRow r = plane_row(cur + 4 * (size_t) c.q[i], tgt + 4 * (size_t) c.m[i], nn);
            int u = 0;
            for (int a = 0; a < 6; ++a) {
                for (int b = a; b < 6; ++b) A[u++].add(r.J[a] * r.J[b], k);
                B[a].add(r.J[a] * r.d, k);
            }
        }
        int u = 0;
        for (int a = 0; a < 6; ++a) {
            for (int b = a; b < 6; ++b) {
                ATA[6 * a + b] = ATA[6 * b + a] = A[u].value(k);
                ++u;
            }
            ATb[a] = B[a].value(k);
        }
        *mse_exact = Sd.value(fs.k_d2) / (double) n;
    } else {
        for (int i = 0; i < 36; ++i) ATA[i] = 0;
        for (int i = 0; i < 6; ++i) ATb[i] = 0;
        for (size_t i = 0; i < n; ++i) {
            const float *nn = nrm + 4 * (size_t) c.m[i];
            if (!(std::isfinite(nn[0]) && std::isfinite(nn[1]) && std::isfinite(nn[2]))) continue;
            const PlaneRow r = plane_row(cur + 4 * (size_t) c.q[i], tgt + 4 * (size_t) c.m[i], nn);
            for (int a = 0; a < 6; ++a) {
                for (int b = a; b < 6; ++b) ATA[6 * a + b] += r.J[a] * r.J[b];
                ATb[a] += r.J[a] * r.d;
            }
        }
        for (int a = 0; a < 6; ++a)
            for (int b = a + 1; b < 6; ++b) ATA[6 * b + a] = ATA[6 * a + b];
    }
    double x[6];
    if (sum_mode == SUM_EXACT) {
        if (!solve_pp<6>(ATA, ATb, x)) return false;
    } else {  // x = ATA.inverse() * ATb
        double inv[36];
        if (!inverse_pp<6>(ATA, inv)) return false;
        for (int a = 0; a < 6; ++a) {
            double s = 0;
            for (int b = 0; b < 6; ++b) s += inv[6 * a + b] * ATb[b];
            x[a] = s;
        }
    }
    construct_lls_transform(x, T);
    return true;
}

}  // namespace

void icp_align(const float *source, size_t n_src, const float *target, size_t n_tgt, const float *target_normals,
               const IcpParams &prm, IcpResult &res, const KdTree *prebuilt, int nn_threads) {
    res = IcpResult();
    set_identity4f(res.final_T);
    res.aligned.assign(source, source + 4 * n_src);
    if (n_src == 0 || n_tgt == 0) return;  // Registration::initCompute fails -> converged_ = false
    KdTree *own = nullptr;
    if (!prebuilt) prebuilt = own = new KdTree(target, n_tgt, 4);
    const KdTree &tree = *prebuilt;
    const FixScales fs = fix_scales(source, n_src, target, n_tgt, prm.max_corr);

    std::vector<float> cur(source, source + 4 * n_src);  // input_transformed
    Correspondences corr;
    double prev_mse = std::numeric_limits<double>::max();
    const double rotation_threshold = 1.0 - prm.t_eps, translation_threshold = prm.t_eps;
    int nr_iterations = 0;
    bool converged = false;
    int state = CONV_NOT_CONVERGED;
    do {
        determine_correspondences(tree, cur.data(), n_src, prm.max_corr, false, corr, nn_threads);
        if (corr.q.size() < 3) {  // min_number_correspondences_
            converged = false;
            state = CONV_NO_CORRESPONDENCES;
            break;
        }
        float T[16];
        double mse_exact = 0;
        bool ok;
        if (prm.estimator == EST_SVD) {
            ok = (prm.sum_mode == SUM_EXACT) ? estimate_svd_exact(cur.data(), target, corr, fs, T, &mse_exact)
                                             : estimate_svd_pcl(cur.data(), target, corr, T);
        } else {
            ok = estimate_plane(cur.data(), target, target_normals, corr, fs, prm.sum_mode, T, &mse_exact);
        }
        if (!ok) {
            converged = false;
            state = CONV_NO_CORRESPONDENCES;
            break;
        }
        // transformCloud: in place, incremental (A.3.2)
        for (size_t i = 0; i < n_src; ++i) {
            float *p = cur.data() + 4 * i;
            if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
            float o[3];
            transform_point4f(T, p, o);
            p[0] = o[0];
            p[1] = o[1];
            p[2] = o[2];
        }
        matmul4f(T, res.final_T, res.final_T);
        ++nr_iterations;

        // DefaultConvergenceCriteria::hasConverged (A.5)
        double mse;
        if (prm.sum_mode == SUM_EXACT) {
            mse = mse_exact;
        } else {
            mse = 0;
            for (size_t i = 0; i < corr.d2.size(); ++i) mse += corr.d2[i];
            mse /= double(corr.d2.size());
        }
        IcpTraceRow row;
        row.mse = mse;
        row.n_corr = (int) corr.q.size();
        std::memcpy(row.T, T, sizeof row.T);
        res.trace.push_back(row);

        converged = false;
        state = CONV_NOT_CONVERGED;
        if (nr_iterations >= prm.max_iter) {
            converged = true;
            state = CONV_ITERATIONS;
        } else {
            const double cos_angle = 0.5 * (T[0] + T[5] + T[10] - 1);
            const double translation_sqr = T[3] * T[3] + T[7] * T[7] + T[11] * T[11];
            if (cos_angle >= rotation_threshold && translation_sqr <= translation_threshold) {
                converged = true;  // max_iterations_similar_transforms_ = 0
                state = CONV_TRANSFORM;
            } else if (std::fabs(mse - prev_mse) < 1e-12) {
                converged = true;
                state = CONV_ABS_MSE;
            } else if (std::fabs(mse - prev_mse) / prev_mse < prm.fit_eps) {
                converged = true;
                state = CONV_REL_MSE;
            } else {
                prev_mse = mse;
            }
        }
    } while (!converged);

    res.converged = converged;
    res.iterations = nr_iterations;
    res.state = state;
    res.corr_query = corr.q;
    res.corr_match = corr.m;
    res.corr_dist = corr.d2;
    for (size_t i = 0; i < n_src; ++i) {  // output = final_transformation_ (x) *input_
        const float *p = source + 4 * i;
        float *o = res.aligned.data() + 4 * i;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
        transform_point4f(res.final_T, p, o);
    }
    delete own;
}

}  // namespace wo
