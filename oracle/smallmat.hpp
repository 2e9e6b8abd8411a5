// ORACLE - test infrastructure only (see kdtree.hpp header).  Tiny dense linear algebra in IEEE
// fp64 with a fixed operation order (built with -ffp-contract=off), used by the estimator
// restatements.  Independent of the device code in libwave_b200/csrc; both follow the written
// spec in DESIGN.md ("Estimator arithmetic") so that results can be compared bit for bit.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdint>

namespace wo {

// ---- order-independent accumulation: 128-bit fixed point ------------------------------------
// term -> llrint(term * 2^k) (round-to-nearest-even), summed exactly in a signed 128-bit integer,
// converted back sign-magnitude: |v| = hi * 2^64 + lo as ldexp((double)hi, 64 - k) + ldexp((double)lo, -k),
// then the sign (the two's-complement halves of a negative total would lose its low bits).
struct Fix128 {
    __int128 v = 0;
    inline void add(double term, int k) { v += (__int128) std::llrint(std::ldexp(term, k)); }
    inline double value(int k) const {
        const bool neg = v < 0;
        const unsigned __int128 m = neg ? (unsigned __int128) 0 - (unsigned __int128) v : (unsigned __int128) v;
        const uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t) m;
        const double mag = std::ldexp((double) hi, 64 - k) + std::ldexp((double) lo, -k);
        return neg ? -mag : mag;
    }
};

// smallest e with 2^e >= x (x > 0), as frexp reports it
static inline int pow2_exponent(double x) {
    int e;
    std::frexp(x, &e);
    return e;
}

// ---- 3x3 one-sided Jacobi (Hestenes) -----------------------------------------------------------
// Rotation that maximises trace(R^T S) for a 3x3 S (Umeyama / Kabsch with the reflection fix):
//   R = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T  for the two dominant singular pairs of S,
// which equals U diag(1,1,det(U)det(V)) V^T (Eigen::umeyama) without needing the third pair.
// Fixed 12 sweeps over the column pairs (0,1),(0,2),(1,2); a pair is skipped when its columns are
// already exactly orthogonal (gamma == 0).
static inline void rotation_from_sigma(const double S[9] /*row major*/, double R[9]) {
    double A[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = S[3 * i + j];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    static const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
    for (int sweep = 0; sweep < 12; ++sweep) {
        for (int pr = 0; pr < 3; ++pr) {
            const int p = P[pr], q = Q[pr];
            const double alpha = (A[0][p] * A[0][p] + A[1][p] * A[1][p]) + A[2][p] * A[2][p];
            const double beta = (A[0][q] * A[0][q] + A[1][q] * A[1][q]) + A[2][q] * A[2][q];
            const double gamma = (A[0][p] * A[0][q] + A[1][p] * A[1][q]) + A[2][p] * A[2][q];
            if (gamma == 0.0) continue;
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / std::sqrt(1.0 + t * t);
            const double s = c * t;
            for (int i = 0; i < 3; ++i) {
                const double ap = A[i][p], aq = A[i][q];
                A[i][p] = c * ap - s * aq;
                A[i][q] = s * ap + c * aq;
                const double vp = V[i][p], vq = V[i][q];
                V[i][p] = c * vp - s * vq;
                V[i][q] = s * vp + c * vq;
            }
        }
    }
    double nrm2[3];
    for (int j = 0; j < 3; ++j) nrm2[j] = (A[0][j] * A[0][j] + A[1][j] * A[1][j]) + A[2][j] * A[2][j];
    // indices of the two largest column norms (ties -> lower index first)
    int i1 = 0;
    if (nrm2[1] > nrm2[i1]) i1 = 1;
    if (nrm2[2] > nrm2[i1]) i1 = 2;
    int i2 = -1;
    for (int j = 0; j < 3; ++j) {
        if (j == i1) continue;
        if (i2 < 0 || nrm2[j] > nrm2[i2]) i2 = j;
    }
    const double s1 = std::sqrt(nrm2[i1]), s2 = std::sqrt(nrm2[i2]);
    if (!(s1 > 0.0) || !(s2 > 0.0)) {  // rank < 2: rotation undetermined -> identity
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
        return;
    }
    double u1[3], u2[3], v1[3], v2[3], u3[3], v3[3];
    for (int i = 0; i < 3; ++i) {
        u1[i] = A[i][i1] / s1;
        u2[i] = A[i][i2] / s2;
        v1[i] = V[i][i1];
        v2[i] = V[i][i2];
    }
    u3[0] = u1[1] * u2[2] - u1[2] * u2[1];
    u3[1] = u1[2] * u2[0] - u1[0] * u2[2];
    u3[2] = u1[0] * u2[1] - u1[1] * u2[0];
    v3[0] = v1[1] * v2[2] - v1[2] * v2[1];
    v3[1] = v1[2] * v2[0] - v1[0] * v2[2];
    v3[2] = v1[0] * v2[1] - v1[1] * v2[0];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = (u1[i] * v1[j] + u2[i] * v2[j]) + u3[i] * v3[j];
}

// ---- N x N solve, Gaussian elimination with partial pivoting, fixed order -----------------------
// Returns false if a pivot is exactly zero (singular).
template <int N>
static inline bool solve_pp(const double A_in[N * N] /*row major*/, const double b_in[N], double x[N]) {
    double A[N][N + 1];
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) A[i][j] = A_in[N * i + j];
        A[i][N] = b_in[i];
    }
    for (int c = 0; c < N; ++c) {
        int piv = c;
        double best = std::fabs(A[c][c]);
        for (int r = c + 1; r < N; ++r)
            if (std::fabs(A[r][c]) > best) {
                best = std::fabs(A[r][c]);
                piv = r;
            }
        if (best == 0.0 || !std::isfinite(best)) return false;
        if (piv != c)
            for (int j = 0; j <= N; ++j) {
                const double t = A[c][j];
                A[c][j] = A[piv][j];
                A[piv][j] = t;
            }
        for (int r = c + 1; r < N; ++r) {
            const double f = A[r][c] / A[c][c];
            for (int j = c; j <= N; ++j) A[r][j] = A[r][j] - f * A[c][j];
        }
    }
    for (int r = N - 1; r >= 0; --r) {
        double s = A[r][N];
        for (int j = r + 1; j < N; ++j) s = s - A[r][j] * x[j];
        x[r] = s / A[r][r];
    }
    return true;
}

// N x N inverse via N solves (column by column); false if singular.
template <int N>
static inline bool inverse_pp(const double A[N * N], double Ainv[N * N]) {
    for (int c = 0; c < N; ++c) {
        double e[N], x[N];
        for (int i = 0; i < N; ++i) e[i] = (i == c) ? 1.0 : 0.0;
        if (!solve_pp<N>(A, e, x)) return false;
        for (int i = 0; i < N; ++i) Ainv[N * i + c] = x[i];
    }
    return true;
}

template <int N>
static inline void matmul(const double *A, const double *B, double *C) {  // C = A B, row major
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += A[N * i + k] * B[N * k + j];
            C[N * i + j] = s;
        }
}

// fp32 4x4 product in Eigen's evaluation order for fixed 4x4 * 4x4 (column combination):
//   C(:,j) = ((A(:,0)*B(0,j) + A(:,1)*B(1,j)) + A(:,2)*B(2,j)) + A(:,3)*B(3,j)
static inline void matmul4f(const float *A, const float *B, float *C) {  // row major storage
    float out[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float r = A[4 * i + 0] * B[0 * 4 + j];
            r = A[4 * i + 1] * B[1 * 4 + j] + r;
            r = A[4 * i + 2] * B[2 * 4 + j] + r;
            r = A[4 * i + 3] * B[3 * 4 + j] + r;
            out[4 * i + j] = r;
        }
    std::memcpy(C, out, sizeof out);
}

// pt' = T (x,y,z,1) in fp32, Eigen fixed 4x4 * 4x1 order:
//   t = m_i0*x; t = m_i1*y + t; t = m_i2*z + t; t = m_i3*1 + t
static inline void transform_point4f(const float *T /*row major 4x4*/, const float *p, float *out) {
    const float x = p[0], y = p[1], z = p[2];
    for (int i = 0; i < 3; ++i) {
        float t = T[4 * i + 0] * x;
        t = T[4 * i + 1] * y + t;
        t = T[4 * i + 2] * z + t;
        t = T[4 * i + 3] + t;
        out[i] = t;
    }
}

}  // namespace wo
