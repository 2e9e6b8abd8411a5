// ORACLE - test infrastructure only (see kdtree.hpp header).
//
// CPU restatement of pcl::GeneralizedIterativeClosestPoint<PointXYZ,PointXYZ>::align() as the
// reference drives it from wave_matching/src/gicp.cpp:20-64: setCorrespondenceRandomness(corr_rand
// = 10), setMaximumIterations(100), setRotationEpsilon(r_eps = 1e-8), setEuclideanFitnessEpsilon
// (unused by GICP); PCL defaults stay for the rest (transformation_epsilon_ = 5e-4,
// corr_dist_threshold_ = 5, gicp_epsilon_ = 1e-3, max_inner_iterations_ = 20).  Follows SURVEY.md
// Appendix A.7 (PCL 1.8 registration/impl/gicp.hpp, registration/bfgs.h - itself a port of GSL's
// vector_bfgs2 - and Segal et al. 2009).  PCL is not vendored: PARITY UNPINNED at the bit level; the
// reference's own tests (tests/gicp_tests.cpp, Frobenius < 0.1) are re-stated in tests/.
//
// Restated expression by expression, including: covariance products formed in fp32 (pt.x * pt.x)
// and accumulated in fp64; the (1, 1, gicp_epsilon) re-weighting of the singular directions; queries
// = fp32 transformation_ * point; d2 < 25 (strict); M = (R C1 R^T + C2)^-1 in fp64; residuals formed
// in fp32 and widened; gradient through computeRDerivative/matricesInnerProd as written; Fletcher
// line search with cubic/quadratic interpolation; delta scaled by 1/rotation_epsilon (entries of the
// fp32 3x3 block) or 1/transformation_epsilon.  Deviations (documented): the 3x3 SVD of the
// symmetric covariance is a Jacobi eigen-decomposition in fp64 (Eigen: two-sided Jacobi SVD), and
// the 3x3 inverse uses cofactors; both equal Eigen's to rounding.
#include "gicp.hpp"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

#include "kdtree.hpp"
#include "smallmat.hpp"

namespace wo {

namespace {

void eig_sym3_desc(const double A_in[9], double evals[3], double V[9]) {  // eigenvalues descending
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
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
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
    int o0 = 0, o1 = 1, o2 = 2;  // stable sort by descending |eigenvalue| (fixed compare-exchange order)
    if (std::fabs(A[o1][o1]) > std::fabs(A[o0][o0])) std::swap(o0, o1);
    if (std::fabs(A[o2][o2]) > std::fabs(A[o1][o1])) std::swap(o1, o2);
    if (std::fabs(A[o1][o1]) > std::fabs(A[o0][o0])) std::swap(o0, o1);
    const int order[3] = {o0, o1, o2};
    for (int j = 0; j < 3; ++j) {
        evals[j] = A[order[j]][order[j]];
        for (int i = 0; i < 3; ++i) V[3 * i + j] = Q[i][order[j]];
    }
}

bool inv3(const double C[9], double out[9]) {
    const double a = C[0], b = C[1], c = C[2], d = C[3], e = C[4], f = C[5], g = C[6], h = C[7], i = C[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    const double id = 1.0 / det;
    out[0] = (e * i - f * h) * id;
    out[1] = (c * h - b * i) * id;
    out[2] = (b * f - c * e) * id;
    out[3] = (f * g - d * i) * id;
    out[4] = (a * i - c * g) * id;
    out[5] = (c * d - a * f) * id;
    out[6] = (d * h - e * g) * id;
    out[7] = (b * g - a * h) * id;
    out[8] = (a * e - b * d) * id;
    return det != 0.0 && std::isfinite(det);
}

}  // namespace

bool gicp_covariances(const float *cloud, size_t n, const KdTree &tree, int k, double gicp_epsilon,
                      std::vector<double> &covs) {
    if ((size_t) k > n) return false;  // "Number or points in cloud is less than k_correspondences_"
    covs.assign(9 * n, 0.0);
    std::vector<int> idx((size_t) k);
    std::vector<float> d2((size_t) k);
    for (size_t i = 0; i < n; ++i) {
        const int found = tree.knn(cloud + 4 * i, k, idx.data(), d2.data());
        double mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < found; ++j) {
            const float *pt = cloud + 4 * (size_t) idx[(size_t) j];
            mean[0] += pt[0];
            mean[1] += pt[1];
            mean[2] += pt[2];
            cov[0] += pt[0] * pt[0];  // float products, as written in gicp.hpp
            cov[3] += pt[1] * pt[0];
            cov[4] += pt[1] * pt[1];
            cov[6] += pt[2] * pt[0];
            cov[7] += pt[2] * pt[1];
            cov[8] += pt[2] * pt[2];
        }
        for (int d = 0; d < 3; ++d) mean[d] /= static_cast<double>(k);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c <= r; ++c) {
                cov[3 * r + c] /= static_cast<double>(k);
                cov[3 * r + c] -= mean[r] * mean[c];
                cov[3 * c + r] = cov[3 * r + c];
            }
        double ev[3], U[9];
        eig_sym3_desc(cov, ev, U);
        double *out = covs.data() + 9 * i;
        for (int kk = 0; kk < 3; ++kk) {
            const double v = (kk == 2) ? gicp_epsilon : 1.0;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) out[3 * r + c] += v * U[3 * r + kk] * U[3 * c + kk];
        }
    }
    return true;
}

namespace {

// applyState: t.topLeft = Rz(x5) Ry(x4) Rx(x3) * t.topLeft (fp32); t.col(3) += (x0, x1, x2, 0)
void apply_state(float *T /*row major 4x4*/, const double x[6]) {
    const float rx = static_cast<float>(x[3]), ry = static_cast<float>(x[4]), rz = static_cast<float>(x[5]);
    const float cx = std::cos(rx), sx = std::sin(rx), cy = std::cos(ry), sy = std::sin(ry), cz = std::cos(rz),
                sz = std::sin(rz);
    // Rz * Ry * Rx
    const float R[9] = {cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx,
                        sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx,
                        -sy,     cy * sx,                cy * cx};
    float out[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            float acc = R[3 * r + 0] * T[0 * 4 + c];
            acc = R[3 * r + 1] * T[1 * 4 + c] + acc;
            acc = R[3 * r + 2] * T[2 * 4 + c] + acc;
            out[3 * r + c] = acc;
        }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[4 * r + c] = out[3 * r + c];
    T[3] += static_cast<float>(x[0]);
    T[7] += static_cast<float>(x[1]);
    T[11] += static_cast<float>(x[2]);
}

void identity4(float *T) {
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.f : 0.f;
}

// Eigen fixed 4x4 * 4x1 fp32 (column combination), w = 1
inline void xform4(const float *T, const float *p, float *o) {
    for (int r = 0; r < 3; ++r) {
        float t = T[4 * r + 0] * p[0];
        t = T[4 * r + 1] * p[1] + t;
        t = T[4 * r + 2] * p[2] + t;
        t = T[4 * r + 3] + t;
        o[r] = t;
    }
}

struct Problem {  // the OptimizationFunctorWithIndices state
    const float *src;  // `output` cloud (= input, guess is identity)
    const float *tgt;
    const std::vector<int> *is, *it;
    const std::vector<double> *mahal;  // per source index, 9 doubles
    float base[16];                    // base_transformation_ = guess = identity
    long long evals = 0;
    int k = 28;                        // fixed-point exponent of the cost sums (gicp_sum_exponent)
};

// The 13 sums of fdf are accumulated exactly: every term t is rounded once to an integer multiple of
// 2^-k (clamped to +-2^52 units so that absurd trial poses cannot overflow) and added in 128-bit
// integers - the same spec as the ICP estimator sums.  The result does not depend on the
// order of the pairs, which is what lets the device (Morton order, blocks, warps) and this loop (cloud
// order) drive pcl::BFGS through the same trajectory.  PCL itself adds the terms sequentially in fp64;
// the difference is ~1e-12 relative, far below what moves the optimiser, but any difference at all
// can flip a line-search decision - the two implementations must not differ.
// k: terms are bounded by max(|p|, 64 m) * 2^15 (|M| <= 500 from the regularised covariances, residuals
// of trial poses taken as < 64 m).
int gicp_sum_exponent(const float *src, size_t n) {
    double bmax = 0;
    for (size_t i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            const float v = src[4 * i + d];
            if (std::isfinite(v)) bmax = std::max(bmax, (double) std::fabs(v));
        }
    int e;
    std::frexp(std::max(bmax, 64.0) * 32768.0, &e);
    return 50 - e;
}

inline void fix_add(__int128 &acc, double term, double scale) {
    double v = term * scale;
    v = std::fmin(std::fmax(v, -4503599627370496.0), 4503599627370496.0);
    acc += (__int128) std::llrint(v);
}
inline double fix_value(__int128 v, int k) {
    Fix128 f;
    f.v = v;
    return f.value(k);
}

// computeRDerivative as written (matricesInnerProd(m1, m2) = sum_ij m1(j,i) * m2(i,j))
void r_derivative(const double x[6], const double R[9], double g[6]) {
    const double phi = x[3], theta = x[4], psi = x[5];
    const double cphi = std::cos(phi), sphi = std::sin(phi), ctheta = std::cos(theta), stheta = std::sin(theta),
                 cpsi = std::cos(psi), spsi = std::sin(psi);
    double dPhi[9], dTheta[9], dPsi[9];
    auto at = [](double *m, int r, int c) -> double & { return m[3 * r + c]; };
    at(dPhi, 0, 0) = 0.;
    at(dPhi, 1, 0) = 0.;
    at(dPhi, 2, 0) = 0.;
    at(dPhi, 0, 1) = sphi * spsi + cphi * cpsi * stheta;
    at(dPhi, 1, 1) = -cpsi * sphi + cphi * spsi * stheta;
    at(dPhi, 2, 1) = cphi * ctheta;
    at(dPhi, 0, 2) = cphi * spsi - cpsi * sphi * stheta;
    at(dPhi, 1, 2) = -cphi * cpsi - sphi * spsi * stheta;
    at(dPhi, 2, 2) = -ctheta * sphi;
    at(dTheta, 0, 0) = -cpsi * stheta;
    at(dTheta, 1, 0) = -spsi * stheta;
    at(dTheta, 2, 0) = -ctheta;
    at(dTheta, 0, 1) = cpsi * ctheta * sphi;
    at(dTheta, 1, 1) = ctheta * sphi * spsi;
    at(dTheta, 2, 1) = -sphi * stheta;
    at(dTheta, 0, 2) = cphi * cpsi * ctheta;
    at(dTheta, 1, 2) = cphi * ctheta * spsi;
    at(dTheta, 2, 2) = -cphi * stheta;
    at(dPsi, 0, 0) = -ctheta * spsi;
    at(dPsi, 1, 0) = cpsi * ctheta;
    at(dPsi, 2, 0) = 0.;
    at(dPsi, 0, 1) = -cphi * cpsi - sphi * spsi * stheta;
    at(dPsi, 1, 1) = -cphi * spsi + cpsi * sphi * stheta;
    at(dPsi, 2, 1) = 0.;
    at(dPsi, 0, 2) = cpsi * sphi - cphi * spsi * stheta;
    at(dPsi, 1, 2) = sphi * spsi + cphi * cpsi * stheta;
    at(dPsi, 2, 2) = 0.;
    auto inner = [](const double *m1, const double *m2) {
        double r = 0.;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r += m1[3 * j + i] * m2[3 * i + j];
        return r;
    };
    g[3] = inner(dPhi, R);
    g[4] = inner(dTheta, R);
    g[5] = inner(dPsi, R);
}

void fdf(Problem &P, const double x[6], double *f_out, double *g /*nullable*/) {
    ++P.evals;
    float T[16];
    std::memcpy(T, P.base, sizeof T);
    apply_state(T, x);
    __int128 sf = 0, sg[3] = {0, 0, 0}, sR[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const double scale = std::ldexp(1.0, P.k);
    const int m = (int) P.is->size();
    for (int i = 0; i < m; ++i) {
        const float *ps = P.src + 4 * (size_t) (*P.is)[(size_t) i];
        const float *pt = P.tgt + 4 * (size_t) (*P.it)[(size_t) i];
        float pp[3];
        xform4(T, ps, pp);
        const double res[3] = {pp[0] - pt[0], pp[1] - pt[1], pp[2] - pt[2]};  // fp32 differences
        const double *M = P.mahal->data() + 9 * (size_t) (*P.is)[(size_t) i];
        const double temp[3] = {M[0] * res[0] + M[1] * res[1] + M[2] * res[2], M[3] * res[0] + M[4] * res[1] + M[5] * res[2],
                                M[6] * res[0] + M[7] * res[1] + M[8] * res[2]};
        fix_add(sf, res[0] * temp[0] + res[1] * temp[1] + res[2] * temp[2], scale);
        if (g) {
            for (int d = 0; d < 3; ++d) fix_add(sg[d], temp[d], scale);
            float pb[3];
            xform4(P.base, ps, pb);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) fix_add(sR[3 * r + c], (double) pb[r] * temp[c], scale);
        }
    }
    const double f = fix_value(sf, P.k);
    double gt[3], R[9];
    for (int d = 0; d < 3; ++d) gt[d] = fix_value(sg[d], P.k);
    for (int q = 0; q < 9; ++q) R[q] = fix_value(sR[q], P.k);
    if (f_out) *f_out = f / m;
    if (g) {
        for (int d = 0; d < 3; ++d) g[d] = gt[d] * (2.0 / m);
        for (int q = 0; q < 9; ++q) R[q] *= 2.0 / m;
        r_derivative(x, R, g);
    }
}

// ---- pcl/registration/bfgs.h (port of GSL vector_bfgs2) ------------------------------------------------
enum BfgsStatus { kNegativeGradientEpsilon = -3, kNotStarted = -2, kRunning = -1, kSuccess = 0, kNoProgress = 1 };

double poly_eval(const double *c, int n, double x) {  // Eigen::poly_eval
    if (x * x <= 1.0) {
        double val = c[n - 1];
        for (int i = n - 2; i >= 0; --i) val = val * x + c[i];
        return val;
    }
    double val = c[0];
    const double inv_x = 1.0 / x;
    for (int i = 1; i < n; ++i) val = val * inv_x + c[i];
    return std::pow(x, (double) (n - 1)) * val;
}

struct Bfgs {
    Problem &P;
    double rho = 0.01, sigma = 0.01, tau1 = 9, tau2 = 0.05, tau3 = 0.5, step_size = 1;
    int order = 3, bracket_iters = 100, section_iters = 100;
    double f = 0, delta_f = 0, fp0 = 0, pnorm = 0, g0norm = 0;
    double x_cache_key = 0, f_cache_key = 0, g_cache_key = 0, df_cache_key = 0, f_alpha = 0, df_alpha = 0;
    double gradient[6], x0[6], g0[6], p[6], x_alpha[6], g_alpha[6], dx[6];
    explicit Bfgs(Problem &prob) : P(prob) {}

    static double dot(const double *a, const double *b) {
        double s = 0;
        for (int i = 0; i < 6; ++i) s += a[i] * b[i];
        return s;
    }
    static double norm(const double *a) { return std::sqrt(dot(a, a)); }

    void move_to(double alpha) {
        for (int i = 0; i < 6; ++i) x_alpha[i] = x0[i] + alpha * p[i];
        x_cache_key = alpha;
    }
    double slope() { return dot(g_alpha, p); }
    double apply_f(double alpha) {
        if (alpha == f_cache_key) return f_alpha;
        move_to(alpha);
        fdf(P, x_alpha, &f_alpha, nullptr);
        f_cache_key = alpha;
        return f_alpha;
    }
    double apply_df(double alpha) {
        if (alpha == df_cache_key) return df_alpha;
        move_to(alpha);
        if (alpha != g_cache_key) {
            fdf(P, x_alpha, nullptr, g_alpha);
            g_cache_key = alpha;
        }
        df_alpha = slope();
        df_cache_key = alpha;
        return df_alpha;
    }
    void apply_fdf(double alpha, double &fo, double &dfo) {
        if (alpha == f_cache_key && alpha == df_cache_key) {
            fo = f_alpha;
            dfo = df_alpha;
            return;
        }
        if (alpha == f_cache_key || alpha == df_cache_key) {
            fo = apply_f(alpha);
            dfo = apply_df(alpha);
            return;
        }
        move_to(alpha);
        fdf(P, x_alpha, &f_alpha, g_alpha);
        f_cache_key = alpha;
        g_cache_key = alpha;
        df_alpha = slope();
        df_cache_key = alpha;
        fo = f_alpha;
        dfo = df_alpha;
    }
    void change_direction() {
        std::memcpy(x_alpha, x0, sizeof x0);
        x_cache_key = 0.0;
        f_cache_key = 0.0;
        std::memcpy(g_alpha, g0, sizeof g0);
        g_cache_key = 0.0;
        df_alpha = slope();
        df_cache_key = 0.0;
    }

    void init(const double x[6]) {
        delta_f = 0;
        for (int i = 0; i < 6; ++i) dx[i] = 0;
        fdf(P, x, &f, gradient);
        std::memcpy(x0, x, sizeof x0);
        std::memcpy(g0, gradient, sizeof g0);
        g0norm = norm(g0);
        for (int i = 0; i < 6; ++i) p[i] = gradient[i] * -1 / g0norm;
        pnorm = norm(p);
        fp0 = -g0norm;
        std::memcpy(x_alpha, x0, sizeof x0);
        x_cache_key = 0;
        f_alpha = f;
        f_cache_key = 0;
        std::memcpy(g_alpha, g0, sizeof g0);
        g_cache_key = 0;
        df_alpha = slope();
        df_cache_key = 0;
    }

    double interpolate(double a, double fa, double fpa, double b, double fb, double fpb, double xmin, double xmax,
                       int ord) {
        double y, ymin = (xmin - a) / (b - a), ymax = (xmax - a) / (b - a), fmin;
        if (ymin > ymax) std::swap(ymin, ymax);
        if (ord > 2 && !(fpb != fpb) && fpb != std::numeric_limits<double>::infinity()) {
            fpa = fpa * (b - a);
            fpb = fpb * (b - a);
            const double eta = 3 * (fb - fa) - 2 * fpa - fpb, xi = fpa + fpb - 2 * (fb - fa);
            const double c[4] = {fa, fpa, eta, xi};
            y = ymin;
            fmin = poly_eval(c, 4, ymin);
            auto check = [&](double xx) {
                const double yy = poly_eval(c, 4, xx);
                if (yy < fmin) {
                    y = xx;
                    fmin = yy;
                }
            };
            check(ymax);
            // roots of c1 + 2 c2 y + 3 c3 y^2 (PolynomialSolver<Scalar, 2> of bfgs.h: closed form)
            const double q0 = c[1], q1 = 2 * c[2], q2 = 3 * c[3];
            const double a2 = 2 * q2, disc = q1 * q1 - 4 * q0 * q2;
            if (0 < disc) {
                const double sq = std::sqrt(disc);
                double y0 = (-q1 - sq) / a2, y1 = (-q1 + sq) / a2;
                if (y0 > y1) std::swap(y0, y1);
                if (y0 > ymin && y0 < ymax) check(y0);
                if (y1 > ymin && y1 < ymax) check(y1);
            } else if (0 == disc) {
                const double y0 = -q1 / a2;
                if (y0 > ymin && y0 < ymax) check(y0);
            }
        } else {
            fpa = fpa * (b - a);
            const double fl = fa + ymin * (fpa + ymin * (fb - fa - fpa));
            const double fh = fa + ymax * (fpa + ymax * (fb - fa - fpa));
            const double c = 2 * (fb - fa - fpa);
            y = ymin;
            fmin = fl;
            if (fh < fmin) {
                y = ymax;
                fmin = fh;
            }
            if (c > a) {  // sic: compared against a, as in bfgs.h
                const double z = -fpa / c;
                if (z > ymin && z < ymax) {
                    const double fz = fa + z * (fpa + z * (fb - fa - fpa));
                    if (fz < fmin) {
                        y = z;
                        fmin = fz;
                    }
                }
            }
        }
        return a + y * (b - a);
    }

    int line_search(double alpha1, double &alpha_new) {
        double f0, fp0l, falpha, falpha_prev, fpalpha, fpalpha_prev, delta, alpha_next;
        double alpha = alpha1, alpha_prev = 0.0, a, b, fa, fb, fpa, fpb;
        int i = 0;
        apply_fdf(0.0, f0, fp0l);
        falpha_prev = f0;
        fpalpha_prev = fp0l;
        a = 0.0;
        b = alpha;
        fa = f0;
        fb = 0.0;
        fpa = fp0l;
        fpb = 0.0;
        while (i++ < bracket_iters) {
            falpha = apply_f(alpha);
            if (falpha > f0 + alpha * rho * fp0l || falpha >= falpha_prev) {
                a = alpha_prev;
                fa = falpha_prev;
                fpa = fpalpha_prev;
                b = alpha;
                fb = falpha;
                fpb = std::numeric_limits<double>::quiet_NaN();
                break;
            }
            fpalpha = apply_df(alpha);
            if (std::fabs(fpalpha) <= -sigma * fp0l) {
                alpha_new = alpha;
                return kSuccess;
            }
            if (fpalpha >= 0) {
                a = alpha;
                fa = falpha;
                fpa = fpalpha;
                b = alpha_prev;
                fb = falpha_prev;
                fpb = fpalpha_prev;
                break;
            }
            delta = alpha - alpha_prev;
            alpha_next = interpolate(alpha_prev, falpha_prev, fpalpha_prev, alpha, falpha, fpalpha, alpha + delta,
                                     alpha + tau1 * delta, order);
            alpha_prev = alpha;
            falpha_prev = falpha;
            fpalpha_prev = fpalpha;
            alpha = alpha_next;
        }
        while (i++ < section_iters) {
            delta = b - a;
            alpha = interpolate(a, fa, fpa, b, fb, fpb, a + tau2 * delta, b - tau3 * delta, order);
            falpha = apply_f(alpha);
            if ((a - alpha) * fpa <= std::numeric_limits<double>::epsilon()) return kNoProgress;
            if (falpha > f0 + rho * alpha * fp0l || falpha >= fa) {
                b = alpha;
                fb = falpha;
                fpb = std::numeric_limits<double>::quiet_NaN();
            } else {
                fpalpha = apply_df(alpha);
                if (std::fabs(fpalpha) <= -sigma * fp0l) {
                    alpha_new = alpha;
                    return kSuccess;
                }
                if (((b - a) >= 0 && fpalpha >= 0) || ((b - a) <= 0 && fpalpha <= 0)) {
                    b = a;
                    fb = fa;
                    fpb = fpa;
                    a = alpha;
                    fa = falpha;
                    fpa = fpalpha;
                } else {
                    a = alpha;
                    fa = falpha;
                    fpa = fpalpha;
                }
            }
        }
        return kSuccess;
    }

    int one_step(double x[6]) {
        double alpha = 0.0, alpha1;
        const double f0 = f;
        if (pnorm == 0.0 || g0norm == 0.0 || fp0 == 0) {
            for (int i = 0; i < 6; ++i) dx[i] = 0;
            return kNoProgress;
        }
        if (delta_f < 0) {
            const double del = std::max(-delta_f, 10 * std::numeric_limits<double>::epsilon() * std::fabs(f0));
            alpha1 = std::min(1.0, 2.0 * del / (-fp0));
        } else {
            alpha1 = std::fabs(step_size);
        }
        const int status = line_search(alpha1, alpha);
        if (status != kSuccess) return status;
        {  // updatePosition
            double fa, dfa;
            apply_fdf(alpha, fa, dfa);
            f = f_alpha;
            std::memcpy(x, x_alpha, sizeof x_alpha);
            std::memcpy(gradient, g_alpha, sizeof g_alpha);
        }
        delta_f = f - f0;
        double dx0[6], dg0[6];
        for (int i = 0; i < 6; ++i) {
            dx0[i] = x[i] - x0[i];
            dx[i] = dx0[i];
            dg0[i] = gradient[i] - g0[i];
        }
        const double dxg = dot(dx0, gradient), dgg = dot(dg0, gradient), dxdg = dot(dx0, dg0), dgnorm = norm(dg0);
        double A, B;
        if (dxdg != 0) {
            B = dxg / dxdg;
            A = -(1.0 + dgnorm * dgnorm / dxdg) * B + dgg / dxdg;
        } else {
            B = 0;
            A = 0;
        }
        for (int i = 0; i < 6; ++i) {
            p[i] = -A * dx0[i];
            p[i] += gradient[i];
            p[i] += -B * dg0[i];
        }
        std::memcpy(g0, gradient, sizeof g0);
        std::memcpy(x0, x, sizeof x0);
        g0norm = norm(g0);
        pnorm = norm(p);
        const double dir = (dot(p, gradient) > 0) ? -1.0 : 1.0;
        for (int i = 0; i < 6; ++i) p[i] *= dir / pnorm;
        pnorm = norm(p);
        fp0 = dot(p, g0);
        change_direction();
        return kSuccess;
    }
    int test_gradient(double eps) {
        if (eps < 0) return kNegativeGradientEpsilon;
        return norm(gradient) < eps ? kSuccess : kRunning;
    }
};

}  // namespace

void gicp_align(const float *source, size_t n_src, const float *target, size_t n_tgt, const GicpParams &prm,
                GicpResult &res) {
    res = GicpResult();
    identity4(res.final_T);
    if (n_src == 0 || n_tgt == 0) return;
    KdTree tree_tgt(target, n_tgt, 4), tree_src(source, n_src, 4);
    std::vector<double> cov_tgt, cov_src;
    const bool ok_t = gicp_covariances(target, n_tgt, tree_tgt, prm.corr_rand, 1e-3, cov_tgt);
    const bool ok_s = gicp_covariances(source, n_src, tree_src, prm.corr_rand, 1e-3, cov_src);
    if (!ok_t) cov_tgt.assign(9 * n_tgt, 0.0);  // PCL logs an error and carries on with what it has
    if (!ok_s) cov_src.assign(9 * n_src, 0.0);

    float transformation[16], previous[16];
    identity4(transformation);
    identity4(previous);
    std::vector<double> mahal(9 * n_src, 0.0);
    for (size_t i = 0; i < n_src; ++i) mahal[9 * i] = mahal[9 * i + 4] = mahal[9 * i + 8] = 1.0;
    const double dist_threshold = 5.0 * 5.0;  // corr_dist_threshold_
    const double rotation_epsilon = prm.r_eps, transformation_epsilon = 5e-4;
    const int max_inner = 20;
    int nr_iterations = 0;
    bool converged = false;
    while (!converged) {
        std::vector<int> is, it;
        double R[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R[3 * r + c] = (double) transformation[4 * r + c];  // transform_R = t * I
        for (size_t i = 0; i < n_src; ++i) {
            const float *p = source + 4 * i;
            if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) continue;
            float q[4] = {0, 0, 0, 1};
            xform4(transformation, p, q);  // guess (identity) first, then transformation_
            int ni;
            float nd;
            tree_tgt.nn1(q, &ni, &nd);
            if (ni < 0) return;
            if (nd < dist_threshold) {
                const double *C1 = cov_src.data() + 9 * i, *C2 = cov_tgt.data() + 9 * (size_t) ni;
                double M[9], temp[9];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c) M[3 * r + c] = R[3 * r] * C1[c] + R[3 * r + 1] * C1[3 + c] + R[3 * r + 2] * C1[6 + c];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                        temp[3 * r + c] = (M[3 * r] * R[3 * c] + M[3 * r + 1] * R[3 * c + 1] + M[3 * r + 2] * R[3 * c + 2]) + C2[3 * r + c];
                inv3(temp, mahal.data() + 9 * i);
                is.push_back((int) i);
                it.push_back(ni);
            }
        }
        std::memcpy(previous, transformation, sizeof previous);
        res.n_corr = is.size();
        if (is.size() < 4) break;  // NotEnoughPointsException -> caught -> break, converged_ stays false
        // estimateRigidTransformationBFGS
        double x[6];
        x[0] = transformation[3];
        x[1] = transformation[7];
        x[2] = transformation[11];
        x[3] = std::atan2(transformation[9], transformation[10]);
        x[4] = std::asin(-transformation[8]);
        x[5] = std::atan2(transformation[4], transformation[0]);
        Problem P;
        P.src = source;
        P.tgt = target;
        P.is = &is;
        P.it = &it;
        P.mahal = &mahal;
        identity4(P.base);
        P.k = gicp_sum_exponent(source, n_src);
        Bfgs bfgs(P);
        bfgs.init(x);
        int inner = 0, result = kRunning;
        do {
            ++inner;
            result = bfgs.one_step(x);
            if (result) break;
            result = bfgs.test_gradient(1e-2);
        } while (result == kRunning && inner < max_inner);
        res.inner_iterations += inner;
        res.evaluations += P.evals;
        if (!(result == kNoProgress || result == kSuccess || inner == max_inner)) break;  // SolverDidntConverge
        identity4(transformation);
        apply_state(transformation, x);
        double delta = 0.;
        for (int k = 0; k < 4; k++)
            for (int l = 0; l < 4; l++) {
                const double ratio = (k < 3 && l < 3) ? 1. / rotation_epsilon : 1. / transformation_epsilon;
                const double c_delta = ratio * std::fabs(previous[4 * k + l] - transformation[4 * k + l]);
                if (c_delta > delta) delta = c_delta;
            }
        nr_iterations++;
        res.delta_trace.push_back(delta);
        if (nr_iterations >= prm.max_iter || delta < 1) {
            converged = true;
            std::memcpy(previous, transformation, sizeof previous);
        }
    }
    res.converged = converged;
    res.iterations = nr_iterations;
    std::memcpy(res.final_T, previous, sizeof previous);  // previous_transformation_ * guess(= I)
}

}  // namespace wo

// ---- surface normals for the point-to-plane extension --------------------------
// Not in the reference (its ICP is point-to-point); the repo's definition, shared with the device
// code through DESIGN.md: k nearest neighbours (the point itself included), fp64 covariance,
// direction of least variance, flipped towards the sensor origin.
namespace wo {

void estimate_normals(const float *cloud, size_t n, int k, float *normals_xyzw) {
    KdTree tree(cloud, n, 4);
    std::vector<int> idx((size_t) k);
    std::vector<float> d2((size_t) k);
    for (size_t i = 0; i < n; ++i) {
        float *o = normals_xyzw + 4 * i;
        o[3] = 0.f;
        const float *q = cloud + 4 * i;
        if (!(std::isfinite(q[0]) && std::isfinite(q[1]) && std::isfinite(q[2]))) {
            o[0] = o[1] = o[2] = std::numeric_limits<float>::quiet_NaN();
            continue;
        }
        const int m = tree.knn(q, k, idx.data(), d2.data());
        if (m < 3) {
            o[0] = o[1] = o[2] = std::numeric_limits<float>::quiet_NaN();
            continue;
        }
        double mean[3] = {0, 0, 0}, cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int j = 0; j < m; ++j) {
            const float *p = cloud + 4 * (size_t) idx[(size_t) j];
            const double x = p[0], y = p[1], z = p[2];
            mean[0] += x; mean[1] += y; mean[2] += z;
            cov[0] += x * x; cov[1] += x * y; cov[2] += x * z; cov[4] += y * y; cov[5] += y * z; cov[8] += z * z;
        }
        for (int d = 0; d < 3; ++d) mean[d] /= (double) m;
        cov[0] = cov[0] / m - mean[0] * mean[0];
        cov[1] = cov[1] / m - mean[0] * mean[1];
        cov[2] = cov[2] / m - mean[0] * mean[2];
        cov[4] = cov[4] / m - mean[1] * mean[1];
        cov[5] = cov[5] / m - mean[1] * mean[2];
        cov[8] = cov[8] / m - mean[2] * mean[2];
        cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
        double ev[3], U[9];
        eig_sym3_desc(cov, ev, U);
        double nx = U[2], ny = U[5], nz = U[8];
        if (nx * (double) q[0] + ny * (double) q[1] + nz * (double) q[2] > 0) {
            nx = -nx; ny = -ny; nz = -nz;
        }
        o[0] = (float) nx; o[1] = (float) ny; o[2] = (float) nz;
    }
}

}  // namespace wo
