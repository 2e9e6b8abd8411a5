// ORACLE - test infrastructure only (see kdtree.hpp header).
//
// ICPMatcher::estimateCensi (wave_matching/src/icp.cpp:167-397) restated from its mathematics rather
// than from its expanded expressions.  The reference evaluates, for every correspondence
// (a = matched target point, b = reference point, both fp32), the closed-form second derivatives of
//     J(X, Z) = | t + R(roll, pitch, yaw) a - b |^2 ,     R = Rz(yaw) Ry(pitch) Rx(roll),
// at X = (translation, eulerAngles(0,1,2) of the match result):
//     d2J_dX2  += d2J/dX2                      (icp.cpp:262-317, upper triangle only)
//     middle   += D cov_Z D^T                  (icp.cpp:319-389), D(i, j) = d2J / dZ_i dX_j,
//     cov_Z     = j diag(lin, ang, ang, lin, ang, ang) j^T   (icp.cpp:219-246)
// and returns information = (H^-1 middle H^-1)^-1 with H = the symmetric completion of d2J_dX2
// (icp.cpp:391-393).  Here the derivatives are assembled from R and its first and second
// partial derivatives, built numerically from the six trigonometric values:
//     e = t + R a - b,  g_k = R_k a
//     H(t, t) = 2 I,  H(t_i, k) = 2 g_k[i],  H(k, l) = 2 g_k.g_l + 2 e.(R_kl a)
//     D(a_i, t_j) = 2 R(j, i),  D(b_i, t_j) = -2 delta_ij,
//     D(a_i, k) = 2 R(:, i).g_k + 2 e.R_k(:, i),  D(b_i, k) = -2 g_k[i]
// Behaviour kept as the reference has it (tests/golden/censi_fixture.npz holds the outputs of the
// reference's own statements, tests/golden/make_censi_fixture.py):
//  * the Euler angles come from Eigen's eulerAngles(0, 1, 2) (an X-Y-Z factorisation whose first
//    angle is folded into [0, pi]) but are used as (roll, pitch, yaw) of a Z-Y-X product;
//  * the spherical-to-Cartesian Jacobian j uses range / bearing / elevation computed in fp32 (sqrtf,
//    atan2f, atanf) and the sine of the elevation where a cosine would be expected (icp.cpp:222-246);
//  * D is indexed (Z, X) but multiplied as D cov_Z D^T, i.e. cov_Z meets D's X index.
#include <cmath>
#include <cstddef>

#include "smallmat.hpp"

namespace wo {

namespace {

void mat3_mul(const double *A, const double *B, double *C) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = (A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j]) + A[3 * i + 2] * B[6 + j];
}

// Eigen 3.3 MatrixBase::eulerAngles(0, 1, 2) (Geometry/EulerAngles.h)
void euler_angles_012(const double *R, double *out) {
    const int i = 0, j = 1, k = 2;  // a0 = 0, a1 = 1: the "even" permutation
    auto at = [&](int r, int c) { return R[3 * r + c]; };
    double r0 = std::atan2(at(j, k), at(k, k)), r1;
    const double c2 = std::sqrt(at(i, i) * at(i, i) + at(i, j) * at(i, j));
    if (r0 > 0.0) {
        r0 -= M_PI;
        r1 = std::atan2(-at(i, k), -c2);
    } else {
        r1 = std::atan2(-at(i, k), c2);
    }
    const double s1 = std::sin(r0), c1 = std::cos(r0);
    const double r2 = std::atan2(s1 * at(k, i) - c1 * at(j, i), c1 * at(j, j) - s1 * at(k, j));
    out[0] = -r0;
    out[1] = -r1;
    out[2] = -r2;
}

// R = Rz(y) Ry(p) Rx(r) and its partial derivatives: dR[0..2] = d/dr, d/dp, d/dy;
// ddR[0..5] = rr, rp, ry, pp, py, yy
void rotation_derivatives(const double *eul, double *R, double dR[3][9], double ddR[6][9]) {
    const double cr = std::cos(eul[0]), sr = std::sin(eul[0]), cp = std::cos(eul[1]), sp = std::sin(eul[1]),
                 cy = std::cos(eul[2]), sy = std::sin(eul[2]);
    // value, first and second derivative of each elementary rotation
    const double X[3][9] = {{1, 0, 0, 0, cr, -sr, 0, sr, cr}, {0, 0, 0, 0, -sr, -cr, 0, cr, -sr}, {0, 0, 0, 0, -cr, sr, 0, -sr, -cr}};
    const double Y[3][9] = {{cp, 0, sp, 0, 1, 0, -sp, 0, cp}, {-sp, 0, cp, 0, 0, 0, -cp, 0, -sp}, {-cp, 0, -sp, 0, 0, 0, sp, 0, -cp}};
    const double Z[3][9] = {{cy, -sy, 0, sy, cy, 0, 0, 0, 1}, {-sy, -cy, 0, cy, -sy, 0, 0, 0, 0}, {-cy, sy, 0, -sy, -cy, 0, 0, 0, 0}};
    auto prod = [&](int dz, int dy, int dx, double *out) {
        double t[9];
        mat3_mul(Z[dz], Y[dy], t);
        mat3_mul(t, X[dx], out);
    };
    prod(0, 0, 0, R);
    prod(0, 0, 1, dR[0]);
    prod(0, 1, 0, dR[1]);
    prod(1, 0, 0, dR[2]);
    prod(0, 0, 2, ddR[0]);  // rr
    prod(0, 1, 1, ddR[1]);  // rp
    prod(1, 0, 1, ddR[2]);  // ry
    prod(0, 2, 0, ddR[3]);  // pp
    prod(1, 1, 0, ddR[4]);  // py
    prod(2, 0, 0, ddR[5]);  // yy
}

// cov = j diag(lin, ang, ang) j^T for one fp32 point, j as icp.cpp:222-233 builds it
void point_cov(float x, float y, float z, double lin, double ang, double *cov /*3x3*/) {
    const double rg = std::sqrt(x * x + y * y + z * z);      // float expression, float sqrt
    const double br = std::atan2(y, x);                      // atan2f
    const double az = std::atan(z / std::sqrt(x * x + y * y));  // float division, atanf
    const double cb = std::cos(br), sb = std::sin(br), ca = std::cos(az), sa = std::sin(az);
    const double j[9] = {cb * sa, -rg * sb * sa, rg * cb * ca,   //
                         sb * sa, rg * cb * sa,  rg * ca * sb,   //
                         ca,      0.0,           -rg * sa};
    const double s[3] = {lin, ang, ang};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) cov[3 * r + c] = (j[3 * r] * s[0] * j[3 * c] + j[3 * r + 1] * s[1] * j[3 * c + 1]) + j[3 * r + 2] * s[2] * j[3 * c + 2];
}

}  // namespace

// ref / target: xyzw fp32 clouds as ICPMatcher::estimateCensi reads them (the clouds handed to the
// last align()); corr_q / corr_m: icp.correspondences_; T16: Matcher::result (row major).
// H36 / middle36 (may be null) receive the two accumulated matrices.  false: singular.
bool estimate_censi(const float *ref, const float *target, const int *corr_q, const int *corr_m, size_t n_corr,
                    const double *T16, double lin_covar, double ang_covar, double *H36, double *middle36,
                    double *info36) {
    double L[9], Rp[9], eul[3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) L[3 * r + c] = T16[4 * r + c];
    rotation_from_sigma(L, Rp);  // Transform::rotation(): the polar factor of the linear part
    euler_angles_012(Rp, eul);
    double R[9], dR[3][9], ddR[6][9];
    rotation_derivatives(eul, R, dR, ddR);
    const double t[3] = {T16[3], T16[7], T16[11]};
    static const int kSecond[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};

    double H[36], M[36];
    for (int i = 0; i < 36; ++i) H[i] = M[i] = 0.0;
    for (size_t n = 0; n < n_corr; ++n) {
        const float *pa = target + 4 * (size_t) corr_m[n], *pb = ref + 4 * (size_t) corr_q[n];
        const double a[3] = {pa[0], pa[1], pa[2]}, b[3] = {pb[0], pb[1], pb[2]};
        double e[3], g[3][3];
        for (int i = 0; i < 3; ++i) e[i] = t[i] + ((R[3 * i] * a[0] + R[3 * i + 1] * a[1]) + R[3 * i + 2] * a[2]) - b[i];
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < 3; ++i) g[k][i] = (dR[k][3 * i] * a[0] + dR[k][3 * i + 1] * a[1]) + dR[k][3 * i + 2] * a[2];
        // d2J/dX2, upper triangle
        for (int i = 0; i < 3; ++i) {
            H[6 * i + i] += 2.0;
            for (int k = 0; k < 3; ++k) H[6 * i + 3 + k] += 2.0 * g[k][i];
        }
        for (int k = 0; k < 3; ++k)
            for (int l = k; l < 3; ++l) {
                const double *S = ddR[kSecond[k][l]];
                double ea = 0.0;
                for (int i = 0; i < 3; ++i) ea += e[i] * ((S[3 * i] * a[0] + S[3 * i + 1] * a[1]) + S[3 * i + 2] * a[2]);
                H[6 * (3 + k) + 3 + l] += 2.0 * ((g[k][0] * g[l][0] + g[k][1] * g[l][1]) + g[k][2] * g[l][2]) + 2.0 * ea;
            }
        // D(z, x) = d2J / dZ_z dX_x
        double D[36];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                D[6 * i + j] = 2.0 * R[3 * j + i];
                D[6 * (3 + i) + j] = (i == j) ? -2.0 : 0.0;
            }
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < 3; ++i) {
                const double rg = (R[i] * g[k][0] + R[3 + i] * g[k][1]) + R[6 + i] * g[k][2];
                const double er = (e[0] * dR[k][i] + e[1] * dR[k][3 + i]) + e[2] * dR[k][6 + i];
                D[6 * i + 3 + k] = 2.0 * rg + 2.0 * er;
                D[6 * (3 + i) + 3 + k] = -2.0 * g[k][i];
            }
        double ca[9], cb[9];
        point_cov(pa[0], pa[1], pa[2], lin_covar, ang_covar, ca);
        point_cov(pb[0], pb[1], pb[2], lin_covar, ang_covar, cb);
        // middle += D cov_Z D^T, cov_Z = blockdiag(ca, cb)
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k)
                    for (int l = 0; l < 3; ++l)
                        s += D[6 * i + k] * ca[3 * k + l] * D[6 * j + l] + D[6 * i + 3 + k] * cb[3 * k + l] * D[6 * j + 3 + l];
                M[6 * i + j] += s;
            }
    }
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < r; ++c) H[6 * r + c] = H[6 * c + r];
    if (H36)
        for (int i = 0; i < 36; ++i) H36[i] = H[i];
    if (middle36)
        for (int i = 0; i < 36; ++i) middle36[i] = M[i];
    double Hi[36], A[36], B[36];
    if (!inverse_pp<6>(H, Hi)) return false;
    matmul<6>(Hi, M, A);
    matmul<6>(A, Hi, B);
    return inverse_pp<6>(B, info36);
}

}  // namespace wo

extern "C" int wo_estimate_censi(const float *ref, const float *target, const int *q, const int *m, size_t n,
                                 const double *T16, double lin, double ang, double *H36, double *middle36,
                                 double *info36) {
    return wo::estimate_censi(ref, target, q, m, n, T16, lin, ang, H36, middle36, info36) ? 1 : 0;
}
