// ORACLE - test infrastructure only (see kdtree.hpp header).
//
// CPU restatement of pcl::NormalDistributionsTransform<PointXYZ,PointXYZ>::align() as the reference
// drives it from wave_matching/src/ndt.cpp:18-65 (setTransformationEpsilon(t_eps), setStepSize
// (step_size), setResolution(res), setMaximumIterations(max_iter); res clamped to >= 0.05), following
// SURVEY.md Appendix A.8 (PCL 1.8 registration/impl/ndt.hpp, filters/impl/voxel_grid_covariance.hpp;
// Magnusson 2009 eq. 6.8-6.21; More & Thuente 1994).  PCL is not vendored: PARITY UNPINNED at the bit
// level; the reference's own tests (tests/ndt_tests.cpp, Frobenius < 0.12) are re-stated in tests/.
//
// Restated faithfully, including two PCL quirks that shape the result:
//   * PCL 1.8's computeStepLengthMT initialises `interval_converged = (step_max - step_min) > 0`, so
//     with step_max > step_min (always, for libwave's step_size = 3, t_eps = 1e-8) the More-Thuente
//     loop never runs and the step is simply clamp(|delta_p|, step_min, step_max) along the Newton
//     direction (NdtParams::line_search = 0).  PCL >= 1.9 corrected the test to `< 0` and the
//     search runs (line_search = 1, the default here): only that behaviour passes the reference's
//     own smallDisplacement case (tests/ndt_tests.cpp:85-102, res 0.3, 0.2 m, bound 0.12) - with the
//     1.8 line the capped Newton steps on the indefinite Hessians of that case wander metres away
//     (tests/test_oracle_ndt.py records both).
//   * updateDerivatives drops a neighbour whose d2 * exp(...) falls outside [0, 1] *after* its
//     score increment was formed (the increment is discarded with it).
// Deviations (documented): voxels whose covariance has a negative eigenvalue are dropped instead of
// being kept with a zero inverse covariance (no effect on gradient or Hessian); the 3x3 symmetric
// eigen-decomposition and the 6x6 SVD solve are Jacobi iterations in fp64 (Eigen uses QR / two-sided
// Jacobi), equal to rounding.
#include "ndt.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <map>

#include "kdtree.hpp"

namespace wo {

namespace {

// cyclic Jacobi eigen-decomposition of a symmetric 3x3 (row major); eigenvalues ascending,
// eigenvectors in the columns of V
void eig_sym3(const double A_in[9], double evals[3], double V[9]) {
    double A[3][3], Q[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = A_in[3 * i + j];
            Q[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double qkp = Q[k][p], qkq = Q[k][q];
                    Q[k][p] = c * qkp - s * qkq;
                    Q[k][q] = s * qkp + c * qkq;
                }
            }
    }
    int order[3] = {0, 1, 2};
    std::sort(order, order + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
    for (int j = 0; j < 3; ++j) {
        evals[j] = A[order[j]][order[j]];
        for (int i = 0; i < 3; ++i) V[3 * i + j] = Q[i][order[j]];
    }
}

bool inv_sym3(const double C[9], double out[9]) {
    const double a = C[0], b = C[1], c = C[2], d = C[4], e = C[5], f = C[8];
    const double det = a * (d * f - e * e) - b * (b * f - e * c) + c * (b * e - d * c);
    if (det == 0.0 || !std::isfinite(det)) return false;
    const double id = 1.0 / det;
    out[0] = (d * f - e * e) * id;
    out[1] = out[3] = (c * e - b * f) * id;
    out[2] = out[6] = (b * e - c * d) * id;
    out[4] = (a * f - c * c) * id;
    out[5] = out[7] = (b * c - a * e) * id;
    out[8] = (a * d - b * b) * id;
    for (int i = 0; i < 9; ++i)
        if (!std::isfinite(out[i])) return false;
    return true;
}

// x = pinv(H) b through a one-sided Jacobi SVD (JacobiSVD<Matrix6d>(H, FullU | FullV).solve(b)):
// singular values <= eps * 6 * sigma_max are treated as zero.
void svd_solve6(const double H[36], const double b[6], double x[6]) {
    double A[6][6], V[6][6];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            A[i][j] = H[6 * i + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < 5; ++p)
            for (int q = p + 1; q < 6; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int k = 0; k < 6; ++k) {
                    alpha += A[k][p] * A[k][p];
                    beta += A[k][q] * A[k][q];
                    gamma += A[k][p] * A[k][q];
                }
                if (gamma == 0.0 || std::fabs(gamma) <= 1e-300) continue;
                if (std::fabs(gamma) <= 2.220446049250313e-16 * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < 6; ++k) {
                    const double ap = A[k][p], aq = A[k][q];
                    A[k][p] = c * ap - s * aq;
                    A[k][q] = s * ap + c * aq;
                    const double vp = V[k][p], vq = V[k][q];
                    V[k][p] = c * vp - s * vq;
                    V[k][q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    double sig[6], smax = 0;
    for (int j = 0; j < 6; ++j) {
        double s = 0;
        for (int k = 0; k < 6; ++k) s += A[k][j] * A[k][j];
        sig[j] = std::sqrt(s);
        smax = std::max(smax, sig[j]);
    }
    const double thr = 2.220446049250313e-16 * 6 * smax;
    for (int i = 0; i < 6; ++i) x[i] = 0;
    for (int j = 0; j < 6; ++j) {
        if (!(sig[j] > thr)) continue;
        // H = U S V^T, A = H V = U S  ->  x += v_j (u_j . b) / s_j  with u_j = A_j / s_j
        double ub = 0;
        for (int k = 0; k < 6; ++k) ub += A[k][j] * b[k];
        const double coef = ub / (sig[j] * sig[j]);
        for (int i = 0; i < 6; ++i) x[i] += V[i][j] * coef;
    }
}

// (Translation(p0..2) * AngleAxis(p3, X) * AngleAxis(p4, Y) * AngleAxis(p5, Z)).matrix(), Scalar = float
void pose_to_matrix4f(const double p[6], float T[16]) {
    const float rx = static_cast<float>(p[3]), ry = static_cast<float>(p[4]), rz = static_cast<float>(p[5]);
    const float cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz),
                sz = std::sin(rz);
    // Rx * Ry * Rz
    const float R[9] = {cy * cz,
                        -cy * sz,
                        sy,
                        sx * sy * cz + cx * sz,
                        -sx * sy * sz + cx * cz,
                        -sx * cy,
                        -cx * sy * cz + sx * sz,
                        cx * sy * sz + sx * cz,
                        cx * cy};
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[4 * r + c] = R[3 * r + c];
    T[3] = static_cast<float>(p[0]);
    T[7] = static_cast<float>(p[1]);
    T[11] = static_cast<float>(p[2]);
}

// pcl::transformPointCloud(in, out, Matrix4f): fp32, left to right
void transform_cloud4f(const float *in, size_t n, const float *T, float *out) {
    for (size_t i = 0; i < n; ++i) {
        const float x = in[4 * i], y = in[4 * i + 1], z = in[4 * i + 2];
        out[4 * i + 0] = T[0] * x + T[1] * y + T[2] * z + T[3];
        out[4 * i + 1] = T[4] * x + T[5] * y + T[6] * z + T[7];
        out[4 * i + 2] = T[8] * x + T[9] * y + T[10] * z + T[11];
        out[4 * i + 3] = in[4 * i + 3];
    }
}

struct AngleTerms {
    double ja[3], jb[3], jc[3], jd[3], je[3], jf[3], jg[3], jh[3];
    double a2[3], a3[3], b2[3], b3[3], c2[3], c3[3], d1[3], d2[3], d3[3], e1[3], e2[3], e3[3], f1[3], f2[3], f3[3];
};

// computeAngleDerivatives (eq. 6.19 / 6.21), with PCL's |angle| < 10e-5 shortcut
void angle_terms(const double p[6], AngleTerms &t) {
    double cx, cy, cz, sx, sy, sz;
    if (std::fabs(p[3]) < 10e-5) { cx = 1.0; sx = 0.0; } else { cx = std::cos(p[3]); sx = std::sin(p[3]); }
    if (std::fabs(p[4]) < 10e-5) { cy = 1.0; sy = 0.0; } else { cy = std::cos(p[4]); sy = std::sin(p[4]); }
    if (std::fabs(p[5]) < 10e-5) { cz = 1.0; sz = 0.0; } else { cz = std::cos(p[5]); sz = std::sin(p[5]); }
    auto set = [](double *v, double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; };
    set(t.ja, (-sx * sz + cx * sy * cz), (-sx * cz - cx * sy * sz), (-cx * cy));
    set(t.jb, (cx * sz + sx * sy * cz), (cx * cz - sx * sy * sz), (-sx * cy));
    set(t.jc, (-sy * cz), sy * sz, cy);
    set(t.jd, sx * cy * cz, (-sx * cy * sz), sx * sy);
    set(t.je, (-cx * cy * cz), cx * cy * sz, (-cx * sy));
    set(t.jf, (-cy * sz), (-cy * cz), 0);
    set(t.jg, (cx * cz - sx * sy * sz), (-cx * sz - sx * sy * cz), 0);
    set(t.jh, (sx * cz + cx * sy * sz), (cx * sy * cz - sx * sz), 0);
    set(t.a2, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), sx * cy);
    set(t.a3, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), (-cx * cy));
    set(t.b2, (cx * cy * cz), (-cx * cy * sz), (cx * sy));
    set(t.b3, (sx * cy * cz), (-sx * cy * sz), (sx * sy));
    set(t.c2, (-sx * cz - cx * sy * sz), (sx * sz - cx * sy * cz), 0);
    set(t.c3, (cx * cz - sx * sy * sz), (-sx * sy * cz - cx * sz), 0);
    set(t.d1, (-cy * cz), (cy * sz), (sy));
    set(t.d2, (-sx * sy * cz), (sx * sy * sz), (sx * cy));
    set(t.d3, (cx * sy * cz), (-cx * sy * sz), (-cx * cy));
    set(t.e1, (sy * sz), (sy * cz), 0);
    set(t.e2, (-sx * cy * sz), (-sx * cy * cz), 0);
    set(t.e3, (cx * cy * sz), (cx * cy * cz), 0);
    set(t.f1, (-cy * cz), (cy * sz), 0);
    set(t.f2, (-cx * sz - sx * sy * cz), (-cx * cz + sx * sy * sz), 0);
    set(t.f3, (-sx * sz + cx * sy * cz), (-cx * sy * sz - sx * cz), 0);
}

inline double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

}  // namespace

void NdtGrid::build(const float *target, size_t n, float res) {
    leaves.clear();
    centroids.clear();
    const float inv = 1.0f / res;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    bool any = false;
    for (size_t i = 0; i < n; ++i) {
        const float *p = target + 4 * i;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
        any = true;
        for (int d = 0; d < 3; ++d) {
            mn[d] = std::min(mn[d], p[d]);
            mx[d] = std::max(mx[d], p[d]);
        }
    }
    if (!any) return;
    const int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
    const int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
    const int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
    if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) return;  // output.clear()
    int min_b[3], div_b[3];
    for (int d = 0; d < 3; ++d) {
        min_b[d] = static_cast<int>(std::floor(mn[d] * inv));
        div_b[d] = static_cast<int>(std::floor(mx[d] * inv)) - min_b[d] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    struct Acc {
        int n = 0;
        double s[3] = {0, 0, 0}, ss[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        float c[3] = {0, 0, 0};
    };
    std::map<size_t, Acc> acc;  // leaves_ is a std::map: ascending voxel index
    for (size_t i = 0; i < n; ++i) {
        const float *p = target + 4 * i;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
        const int i0 = static_cast<int>(std::floor(p[0] * inv) - static_cast<float>(min_b[0]));
        const int i1 = static_cast<int>(std::floor(p[1] * inv) - static_cast<float>(min_b[1]));
        const int i2 = static_cast<int>(std::floor(p[2] * inv) - static_cast<float>(min_b[2]));
        Acc &a = acc[(size_t) (i0 * mul[0] + i1 * mul[1] + i2 * mul[2])];
        const double q[3] = {p[0], p[1], p[2]};
        for (int r = 0; r < 3; ++r) {
            a.s[r] += q[r];
            a.c[r] += p[r];
            for (int c = 0; c < 3; ++c) a.ss[3 * r + c] += q[r] * q[c];
        }
        ++a.n;
    }
    for (auto &kv : acc) {
        Acc &a = kv.second;
        if (a.n < 6) continue;  // min_points_per_voxel_
        NdtLeaf leaf;
        leaf.voxel = (int) kv.first;
        leaf.n = a.n;
        const float fn = static_cast<float>(a.n);
        for (int d = 0; d < 3; ++d) {
            leaf.centroid[d] = a.c[d] / fn;
            leaf.mean[d] = a.s[d] / a.n;
        }
        double cov[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                cov[3 * r + c] = (a.ss[3 * r + c] - 2 * (a.s[r] * leaf.mean[c])) / a.n + leaf.mean[r] * leaf.mean[c];
        for (int k = 0; k < 9; ++k) cov[k] *= (a.n - 1.0) / a.n;
        double ev[3], V[9];
        eig_sym3(cov, ev, V);
        if (ev[0] < 0 || ev[1] < 0 || ev[2] <= 0) continue;
        const double min_ev = 0.01 * ev[2];  // min_covar_eigvalue_mult_
        if (ev[0] < min_ev) {
            ev[0] = min_ev;
            if (ev[1] < min_ev) ev[1] = min_ev;
            // cov = evecs * diag * evecs^-1 (orthogonal: inverse = transpose)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double s = 0;
                    for (int k = 0; k < 3; ++k) s += V[3 * r + k] * ev[k] * V[3 * c + k];
                    cov[3 * r + c] = s;
                }
        }
        if (!inv_sym3(cov, leaf.icov)) continue;
        leaves.push_back(leaf);
    }
    centroids.resize(4 * leaves.size());
    for (size_t i = 0; i < leaves.size(); ++i) {
        for (int d = 0; d < 3; ++d) centroids[4 * i + d] = leaves[i].centroid[d];
        centroids[4 * i + 3] = 1.0f;
    }
}

double ndt_derivatives(const NdtGrid &grid, const KdTree &tree, const float *source, const float *trans, size_t n,
                       const double p[6], float res, double d1, double d2, bool with_hessian, double g[6],
                       double H[36]) {
    AngleTerms at;
    angle_terms(p, at);
    for (int i = 0; i < 6; ++i) g[i] = 0;
    if (with_hessian)
        for (int i = 0; i < 36; ++i) H[i] = 0;
    double score = 0;
    std::vector<std::pair<float, int>> nb;
    const float r2 = static_cast<float>((double) res * (double) res);
    for (size_t idx = 0; idx < n; ++idx) {
        const float *xt = trans + 4 * idx;
        if (!(std::isfinite(xt[0]) && std::isfinite(xt[1]) && std::isfinite(xt[2]))) continue;
        tree.radius(xt, r2, nb);  // sorted by distance; FLANN keeps d < r^2 (strict)
        for (const auto &hit : nb) {
            if (!(hit.first < r2)) continue;
            const NdtLeaf &cell = grid.leaves[(size_t) hit.second];
            const double x[3] = {source[4 * idx], source[4 * idx + 1], source[4 * idx + 2]};
            const double xtr[3] = {xt[0] - cell.mean[0], xt[1] - cell.mean[1], xt[2] - cell.mean[2]};
            // computePointDerivatives: J is 3x6 (first three columns identity), Hp 18x6
            double J[3][6] = {{1, 0, 0, 0, 0, 0}, {0, 1, 0, 0, 0, 0}, {0, 0, 1, 0, 0, 0}};
            J[1][3] = dot3(x, at.ja);
            J[2][3] = dot3(x, at.jb);
            J[0][4] = dot3(x, at.jc);
            J[1][4] = dot3(x, at.jd);
            J[2][4] = dot3(x, at.je);
            J[0][5] = dot3(x, at.jf);
            J[1][5] = dot3(x, at.jg);
            J[2][5] = dot3(x, at.jh);
            double Hp[6][6][3];
            std::memset(Hp, 0, sizeof Hp);
            if (with_hessian) {
                const double a[3] = {0, dot3(x, at.a2), dot3(x, at.a3)}, b[3] = {0, dot3(x, at.b2), dot3(x, at.b3)},
                             c[3] = {0, dot3(x, at.c2), dot3(x, at.c3)},
                             d[3] = {dot3(x, at.d1), dot3(x, at.d2), dot3(x, at.d3)},
                             e[3] = {dot3(x, at.e1), dot3(x, at.e2), dot3(x, at.e3)},
                             f[3] = {dot3(x, at.f1), dot3(x, at.f2), dot3(x, at.f3)};
                for (int k = 0; k < 3; ++k) {
                    Hp[3][3][k] = a[k];
                    Hp[4][3][k] = b[k];
                    Hp[5][3][k] = c[k];
                    Hp[3][4][k] = b[k];
                    Hp[4][4][k] = d[k];
                    Hp[5][4][k] = e[k];
                    Hp[3][5][k] = c[k];
                    Hp[4][5][k] = e[k];
                    Hp[5][5][k] = f[k];
                }
            }
            // updateDerivatives
            const double *C = cell.icov;
            const double Cx[3] = {C[0] * xtr[0] + C[1] * xtr[1] + C[2] * xtr[2],
                                  C[3] * xtr[0] + C[4] * xtr[1] + C[5] * xtr[2],
                                  C[6] * xtr[0] + C[7] * xtr[1] + C[8] * xtr[2]};
            double e_x_cov_x = std::exp(-d2 * dot3(xtr, Cx) / 2);
            const double score_inc = -d1 * e_x_cov_x;
            e_x_cov_x = d2 * e_x_cov_x;
            if (e_x_cov_x > 1 || e_x_cov_x < 0 || e_x_cov_x != e_x_cov_x) continue;
            e_x_cov_x *= d1;
            double cJ[6][3], xcJ[6];
            for (int i = 0; i < 6; ++i) {
                for (int r = 0; r < 3; ++r) cJ[i][r] = C[3 * r] * J[0][i] + C[3 * r + 1] * J[1][i] + C[3 * r + 2] * J[2][i];
                xcJ[i] = dot3(xtr, cJ[i]);
            }
            for (int i = 0; i < 6; ++i) {
                g[i] += xcJ[i] * e_x_cov_x;
                if (with_hessian)
                    for (int j = 0; j < 6; ++j) {
                        const double cH[3] = {C[0] * Hp[i][j][0] + C[1] * Hp[i][j][1] + C[2] * Hp[i][j][2],
                                              C[3] * Hp[i][j][0] + C[4] * Hp[i][j][1] + C[5] * Hp[i][j][2],
                                              C[6] * Hp[i][j][0] + C[7] * Hp[i][j][1] + C[8] * Hp[i][j][2]};
                        const double JjcJi = J[0][j] * cJ[i][0] + J[1][j] * cJ[i][1] + J[2][j] * cJ[i][2];
                        H[6 * i + j] += e_x_cov_x * (-d2 * xcJ[i] * xcJ[j] + dot3(xtr, cH) + JjcJi);
                    }
            }
            score += score_inc;
        }
    }
    return score;
}

namespace {

double psi_mt(double a, double f_a, double f_0, double g_0, double mu) { return f_a - f_0 - mu * g_0 * a; }
double dpsi_mt(double g_a, double g_0, double mu) { return g_a - mu * g_0; }

bool update_interval_mt(double &a_l, double &f_l, double &g_l, double &a_u, double &f_u, double &g_u, double a_t,
                        double f_t, double g_t) {
    if (f_t > f_l) {
        a_u = a_t; f_u = f_t; g_u = g_t;
        return false;
    } else if (g_t * (a_l - a_t) > 0) {
        a_l = a_t; f_l = f_t; g_l = g_t;
        return false;
    } else if (g_t * (a_l - a_t) < 0) {
        a_u = a_l; f_u = f_l; g_u = g_l;
        a_l = a_t; f_l = f_t; g_l = g_t;
        return false;
    }
    return true;
}

double trial_value_mt(double a_l, double f_l, double g_l, double a_u, double f_u, double g_u, double a_t, double f_t,
                      double g_t) {
    if (f_t > f_l) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        const double w = std::sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_q = a_l - 0.5 * (a_l - a_t) * g_l / (g_l - (f_l - f_t) / (a_l - a_t));
        if (std::fabs(a_c - a_l) < std::fabs(a_q - a_l)) return a_c;
        return 0.5 * (a_q + a_c);
    } else if (g_t * g_l < 0) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        const double w = std::sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        if (std::fabs(a_c - a_t) >= std::fabs(a_s - a_t)) return a_c;
        return a_s;
    } else if (std::fabs(g_t) <= std::fabs(g_l)) {
        const double z = 3 * (f_t - f_l) / (a_t - a_l) - g_t - g_l;
        const double w = std::sqrt(z * z - g_t * g_l);
        const double a_c = a_l + (a_t - a_l) * (w - g_l - z) / (g_t - g_l + 2 * w);
        const double a_s = a_l - (a_l - a_t) / (g_l - g_t) * g_l;
        const double a_t_next = (std::fabs(a_c - a_t) < std::fabs(a_s - a_t)) ? a_c : a_s;
        if (a_t > a_l) return std::min(a_t + 0.66 * (a_u - a_t), a_t_next);
        return std::max(a_t + 0.66 * (a_u - a_t), a_t_next);
    }
    const double z = 3 * (f_t - f_u) / (a_t - a_u) - g_t - g_u;
    const double w = std::sqrt(z * z - g_t * g_u);
    return a_u + (a_t - a_u) * (w - g_u - z) / (g_t - g_u + 2 * w);
}

}  // namespace

void ndt_align(const float *source, size_t n_src, const float *target, size_t n_tgt, const NdtParams &prm,
               NdtResult &res) {
    res = NdtResult();
    for (int i = 0; i < 16; ++i) res.final_T[i] = (i % 5 == 0) ? 1.f : 0.f;
    if (n_src == 0 || n_tgt == 0) return;
    float resolution = prm.res;
    if (resolution < 0.05f) resolution = 0.05f;  // NDTMatcher ctor, src/ndt.cpp:23-26 (min_res)
    NdtGrid grid;
    grid.build(target, n_tgt, resolution);
    res.n_voxels = (int) grid.leaves.size();
    KdTree tree(grid.centroids.data(), grid.leaves.size(), 4);

    const double outlier_ratio = 0.55;
    const double c1 = 10 * (1 - outlier_ratio);
    const double c2 = outlier_ratio / std::pow((double) resolution, 3);
    const double d3 = -std::log(c2);
    const double gd1 = -std::log(c1 + c2) - d3;
    const double gd2 = -2 * std::log((-std::log(c1 * std::exp(-0.5) + c2) - d3) / gd1);
    const double step_size = (double) prm.step_size, t_eps = prm.t_eps;

    std::vector<float> trans(source, source + 4 * n_src);  // output = *input_ (guess = identity)
    double p[6] = {0, 0, 0, 0, 0, 0}, g[6], H[36], delta_p[6];
    double score = ndt_derivatives(grid, tree, source, trans.data(), n_src, p, resolution, gd1, gd2, true, g, H);
    int nr_iterations = 0;
    bool converged = false;
    while (!converged) {
        double neg_g[6];
        for (int i = 0; i < 6; ++i) neg_g[i] = -g[i];
        svd_solve6(H, neg_g, delta_p);
        double nrm = 0;
        for (int i = 0; i < 6; ++i) nrm += delta_p[i] * delta_p[i];
        double delta_p_norm = std::sqrt(nrm);
        if (delta_p_norm == 0 || delta_p_norm != delta_p_norm) {
            converged = delta_p_norm == delta_p_norm;
            break;
        }
        for (int i = 0; i < 6; ++i) delta_p[i] /= delta_p_norm;

        // ---- computeStepLengthMT(p, delta_p, delta_p_norm, step_size, t_eps / 2, ...) ----
        double a_t;
        {
            const double step_init = delta_p_norm, step_max = step_size, step_min = t_eps / 2;
            const double phi_0 = -score;
            double d_phi_0 = 0;
            for (int i = 0; i < 6; ++i) d_phi_0 -= g[i] * delta_p[i];
            bool skip = false;
            if (d_phi_0 >= 0) {
                if (d_phi_0 == 0) {
                    a_t = 0;
                    skip = true;
                } else {
                    d_phi_0 *= -1;
                    for (int i = 0; i < 6; ++i) delta_p[i] *= -1;
                }
            }
            if (!skip) {
                const int max_step_iterations = 10;
                int step_iterations = 0;
                const double mu = 1.e-4, nu = 0.9;
                double a_l = 0, a_u = 0;
                double f_l = psi_mt(a_l, phi_0, phi_0, d_phi_0, mu), g_l = dpsi_mt(d_phi_0, d_phi_0, mu);
                double f_u = psi_mt(a_u, phi_0, phi_0, d_phi_0, mu), g_u = dpsi_mt(d_phi_0, d_phi_0, mu);
                // PCL 1.8: `(step_max - step_min) > 0` (sic) - true whenever step_max > step_min, which skips
                // the loop below; PCL >= 1.9: `< 0`, the More-Thuente search runs
                bool interval_converged = prm.line_search ? (step_max - step_min) < 0 : (step_max - step_min) > 0,
                     open_interval = true;
                a_t = std::max(std::min(step_init, step_max), step_min);
                double x_t[6];
                auto evaluate = [&](bool hess) {
                    for (int i = 0; i < 6; ++i) x_t[i] = p[i] + delta_p[i] * a_t;
                    pose_to_matrix4f(x_t, res.final_T);
                    transform_cloud4f(source, n_src, res.final_T, trans.data());
                    score = ndt_derivatives(grid, tree, source, trans.data(), n_src, x_t, resolution, gd1, gd2, hess, g,
                                            H);
                };
                evaluate(true);
                double phi_t = -score, d_phi_t = 0;
                for (int i = 0; i < 6; ++i) d_phi_t -= g[i] * delta_p[i];
                double psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu), d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
                while (!interval_converged && step_iterations < max_step_iterations &&
                       !(psi_t <= 0 && d_phi_t <= -nu * d_phi_0)) {
                    if (open_interval) a_t = trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
                    else a_t = trial_value_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
                    a_t = std::max(std::min(a_t, step_max), step_min);
                    evaluate(false);
                    phi_t = -score;
                    d_phi_t = 0;
                    for (int i = 0; i < 6; ++i) d_phi_t -= g[i] * delta_p[i];
                    psi_t = psi_mt(a_t, phi_t, phi_0, d_phi_0, mu);
                    d_psi_t = dpsi_mt(d_phi_t, d_phi_0, mu);
                    if (open_interval && (psi_t <= 0 && d_psi_t >= 0)) {
                        open_interval = false;
                        f_l = f_l + phi_0 - mu * d_phi_0 * a_l;
                        g_l = g_l + mu * d_phi_0;
                        f_u = f_u + phi_0 - mu * d_phi_0 * a_u;
                        g_u = g_u + mu * d_phi_0;
                    }
                    if (open_interval)
                        interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, psi_t, d_psi_t);
                    else
                        interval_converged = update_interval_mt(a_l, f_l, g_l, a_u, f_u, g_u, a_t, phi_t, d_phi_t);
                    step_iterations++;
                }
                if (step_iterations) {  // computeHessian at the accepted point
                    double gtmp[6];
                    ndt_derivatives(grid, tree, source, trans.data(), n_src, x_t, resolution, gd1, gd2, true, gtmp, H);
                }
            }
        }
        delta_p_norm = a_t;
        for (int i = 0; i < 6; ++i) {
            delta_p[i] *= delta_p_norm;
            p[i] = p[i] + delta_p[i];
        }
        res.step_trace.push_back(delta_p_norm);
        res.score_trace.push_back(score);
        if (nr_iterations > prm.max_iter || (nr_iterations && (std::fabs(delta_p_norm) < t_eps))) converged = true;
        nr_iterations++;
    }
    res.converged = converged;
    res.iterations = nr_iterations;
    res.score = score;
    std::memcpy(res.pose, p, sizeof p);
}

}  // namespace wo
