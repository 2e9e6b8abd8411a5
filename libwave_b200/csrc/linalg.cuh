// Small fp64 linear algebra shared by the device kernels.
#pragma once
#include "common.cuh"

namespace wavecu {

// symmetric 3x3 Jacobi, eigenvalues by descending magnitude (= singular values of a symmetric matrix)
__device__ inline void eig_sym3_desc(const double A_in[9], double V[9]) {
    double A[3][3], Q[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            A[i][j] = A_in[3 * i + j];
            Q[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
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
    int o0 = 0, o1 = 1, o2 = 2;  // stable sort by descending |eigenvalue|
    if (fabs(A[o1][o1]) > fabs(A[o0][o0])) { const int t = o0; o0 = o1; o1 = t; }
    if (fabs(A[o2][o2]) > fabs(A[o1][o1])) { const int t = o1; o1 = o2; o2 = t; }
    if (fabs(A[o1][o1]) > fabs(A[o0][o0])) { const int t = o0; o0 = o1; o1 = t; }
    const int order[3] = {o0, o1, o2};
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i) V[3 * i + j] = Q[i][order[j]];
}

}  // namespace wavecu
