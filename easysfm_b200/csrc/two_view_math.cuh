// two_view_math.cuh -- float64 building blocks of the two-view geometric verification (two_view.cu), SURVEY 8f rank 1:
// the reference's MotionEstimator::estimate2D2D_E5P_RANSAC (cpp_code/src/estimate_motion.cpp:27-97: cv::findEssentialMat(RANSAC) +
// cv::recoverPose) and getDepthFast (:234-283: cv::triangulatePoints + mean point norm).
//
// Everything is __host__ __device__ so that a host harness (tests/host/two_view_host.cpp, built by tests/two_view_util.py) can compare these functions, without a GPU, with the
// numpy oracle (oracle/two_view_oracle.py) that restates the same published algorithms through LAPACK:
//   * sample_indices      the counter-based 5-subset generator shared with the oracle (OpenCV's cv::RNG sequence cannot be restated);
//   * five_point          Nister's minimal problem by the action-matrix method: null space of the 5 x 9 epipolar system (Gauss-Jordan with
//                         full pivoting + Gram-Schmidt), the ten cubic constraints expanded with small polynomial products, Gauss-Jordan on
//                         the 10 x 20 coefficient matrix, real eigenvalues of the 10 x 10 multiplication matrix (Hessenberg reduction +
//                         Francis double-shift QR, EISPACK hqr), eigenvectors by elimination with complete pivoting;
//   * sampson_error       OpenCV's model error for essential matrices;
//   * decompose_essential, triangulate_dlt (smallest eigenvector of A^T A by cyclic Jacobi), the pieces of recoverPose.
#pragma once
#include <cmath>
#include <cstdint>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace esfm {
namespace tv {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ inline uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// five distinct indices in [0, m), m >= 5 (oracle.two_view_oracle.sample_indices)
__host__ __device__ inline void sample_indices(uint64_t seed, uint64_t pair, uint64_t hyp, uint32_t m, int (&idx)[5]) {
    uint64_t key = splitmix64(seed ^ splitmix64(pair));
    key = splitmix64(key ^ (hyp * 0xD6E8FEB86659FD93ull));
    int n = 0;
    for (uint64_t c = 0; n < 5; ++c) {
        const int i = (int)mulhi64(splitmix64(key + c), (uint64_t)m);
        bool dup = false;
        for (int k = 0; k < n; ++k) dup |= idx[k] == i;
        if (!dup) idx[n++] = i;
    }
}

// ---- polynomials in (x, y, z): linear [x, y, z, 1]; quadratic [x2, xy, xz, y2, yz, z2, x, y, z, 1]; cubic in the order of the 10 x 20
//      system: [x3, x2y, x2z, xy2, xyz, xz2, y3, y2z, yz2, z3 | x2, xy, xz, y2, yz, z2, x, y, z, 1] ----
__host__ __device__ inline void mul_ll(const double* a, const double* b, double* q) {
    q[0] = a[0] * b[0];
    q[1] = a[0] * b[1] + a[1] * b[0];
    q[2] = a[0] * b[2] + a[2] * b[0];
    q[3] = a[1] * b[1];
    q[4] = a[1] * b[2] + a[2] * b[1];
    q[5] = a[2] * b[2];
    q[6] = a[0] * b[3] + a[3] * b[0];
    q[7] = a[1] * b[3] + a[3] * b[1];
    q[8] = a[2] * b[3] + a[3] * b[2];
    q[9] = a[3] * b[3];
}
// c += s * (quadratic q) * (linear l)
__host__ __device__ inline void fma_ql(const double* q, const double* l, double s, double* c) {
    const double lx = s * l[0], ly = s * l[1], lz = s * l[2], l1 = s * l[3];
    c[0] += q[0] * lx;                                  // x3
    c[1] += q[0] * ly + q[1] * lx;                      // x2y
    c[2] += q[0] * lz + q[2] * lx;                      // x2z
    c[3] += q[1] * ly + q[3] * lx;                      // xy2
    c[4] += q[1] * lz + q[2] * ly + q[4] * lx;          // xyz
    c[5] += q[2] * lz + q[5] * lx;                      // xz2
    c[6] += q[3] * ly;                                  // y3
    c[7] += q[3] * lz + q[4] * ly;                      // y2z
    c[8] += q[4] * lz + q[5] * ly;                      // yz2
    c[9] += q[5] * lz;                                  // z3
    c[10] += q[0] * l1 + q[6] * lx;                     // x2
    c[11] += q[1] * l1 + q[6] * ly + q[7] * lx;         // xy
    c[12] += q[2] * l1 + q[6] * lz + q[8] * lx;         // xz
    c[13] += q[3] * l1 + q[7] * ly;                     // y2
    c[14] += q[4] * l1 + q[7] * lz + q[8] * ly;         // yz
    c[15] += q[5] * l1 + q[8] * lz;                     // z2
    c[16] += q[6] * l1 + q[9] * lx;                     // x
    c[17] += q[7] * l1 + q[9] * ly;                     // y
    c[18] += q[8] * l1 + q[9] * lz;                     // z
    c[19] += q[9] * l1;                                 // 1
}

// ---- real eigenvalues of a 10 x 10 matrix: elimination to Hessenberg form + Francis double-shift QR (EISPACK elmhes / hqr, 1-based) ----
constexpr int kEig = 10;
__host__ __device__ inline double sign_of(double a, double b) { return b >= 0.0 ? fabs(a) : -fabs(a); }

__host__ __device__ inline void elmhes(double (&a)[kEig + 1][kEig + 1]) {
    const int n = kEig;
    for (int m = 2; m < n; ++m) {
        double x = 0.0;
        int i = m;
        for (int j = m; j <= n; ++j)
            if (fabs(a[j][m - 1]) > fabs(x)) { x = a[j][m - 1]; i = j; }
        if (i != m) {
            for (int j = m - 1; j <= n; ++j) { const double t = a[i][j]; a[i][j] = a[m][j]; a[m][j] = t; }
            for (int j = 1; j <= n; ++j) { const double t = a[j][i]; a[j][i] = a[j][m]; a[j][m] = t; }
        }
        if (x != 0.0) {
            for (int ii = m + 1; ii <= n; ++ii) {
                double y = a[ii][m - 1];
                if (y != 0.0) {
                    y /= x;
                    a[ii][m - 1] = y;
                    for (int j = m; j <= n; ++j) a[ii][j] -= y * a[m][j];
                    for (int j = 1; j <= n; ++j) a[j][m] += y * a[j][ii];
                }
            }
        }
    }
    for (int i = 3; i <= n; ++i)
        for (int j = 1; j <= i - 2; ++j) a[i][j] = 0.0;        // (the multipliers elmhes leaves below the subdiagonal)
}

// returns false if an eigenvalue did not converge in 30 iterations
__host__ __device__ inline bool hqr(double (&a)[kEig + 1][kEig + 1], double (&wr)[kEig + 1], double (&wi)[kEig + 1]) {
    const int n = kEig;
    int nn, m, l, k, j, its, i, mmin;
    double z, y, x, w, v, u, t, s, r = 0.0, q = 0.0, p = 0.0, anorm = 0.0;
    for (i = 1; i <= n; ++i)
        for (j = (i - 1 > 1 ? i - 1 : 1); j <= n; ++j) anorm += fabs(a[i][j]);
    nn = n;
    t = 0.0;
    while (nn >= 1) {
        its = 0;
        do {
            for (l = nn; l >= 2; --l) {
                s = fabs(a[l - 1][l - 1]) + fabs(a[l][l]);
                if (s == 0.0) s = anorm;
                if (fabs(a[l][l - 1]) + s == s) { a[l][l - 1] = 0.0; break; }
            }
            x = a[nn][nn];
            if (l == nn) {
                wr[nn] = x + t;
                wi[nn--] = 0.0;
            } else {
                y = a[nn - 1][nn - 1];
                w = a[nn][nn - 1] * a[nn - 1][nn];
                if (l == nn - 1) {
                    p = 0.5 * (y - x);
                    q = p * p + w;
                    z = sqrt(fabs(q));
                    x += t;
                    if (q >= 0.0) {
                        z = p + sign_of(z, p);
                        wr[nn - 1] = wr[nn] = x + z;
                        if (z != 0.0) wr[nn] = x - w / z;
                        wi[nn - 1] = wi[nn] = 0.0;
                    } else {
                        wr[nn - 1] = wr[nn] = x + p;
                        wi[nn - 1] = -(wi[nn] = z);
                    }
                    nn -= 2;
                } else {
                    if (its == 30) return false;
                    if (its == 10 || its == 20) {
                        t += x;
                        for (i = 1; i <= nn; ++i) a[i][i] -= x;
                        s = fabs(a[nn][nn - 1]) + fabs(a[nn - 1][nn - 2]);
                        y = x = 0.75 * s;
                        w = -0.4375 * s * s;
                    }
                    ++its;
                    for (m = nn - 2; m >= l; --m) {
                        z = a[m][m];
                        r = x - z;
                        s = y - z;
                        p = (r * s - w) / a[m + 1][m] + a[m][m + 1];
                        q = a[m + 1][m + 1] - z - r - s;
                        r = a[m + 2][m + 1];
                        s = fabs(p) + fabs(q) + fabs(r);
                        p /= s; q /= s; r /= s;
                        if (m == l) break;
                        u = fabs(a[m][m - 1]) * (fabs(q) + fabs(r));
                        v = fabs(p) * (fabs(a[m - 1][m - 1]) + fabs(z) + fabs(a[m + 1][m + 1]));
                        if (u + v == v) break;
                    }
                    for (i = m + 2; i <= nn; ++i) {
                        a[i][i - 2] = 0.0;
                        if (i != m + 2) a[i][i - 3] = 0.0;
                    }
                    for (k = m; k <= nn - 1; ++k) {
                        if (k != m) {
                            p = a[k][k - 1];
                            q = a[k + 1][k - 1];
                            r = 0.0;
                            if (k != nn - 1) r = a[k + 2][k - 1];
                            if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0) { p /= x; q /= x; r /= x; }
                        }
                        if ((s = sign_of(sqrt(p * p + q * q + r * r), p)) != 0.0) {
                            if (k == m) {
                                if (l != m) a[k][k - 1] = -a[k][k - 1];
                            } else
                                a[k][k - 1] = -s * x;
                            p += s;
                            x = p / s; y = q / s; z = r / s;
                            q /= p; r /= p;
                            for (j = k; j <= nn; ++j) {
                                p = a[k][j] + q * a[k + 1][j];
                                if (k != nn - 1) { p += r * a[k + 2][j]; a[k + 2][j] -= p * z; }
                                a[k + 1][j] -= p * y;
                                a[k][j] -= p * x;
                            }
                            mmin = nn < k + 3 ? nn : k + 3;
                            for (i = l; i <= mmin; ++i) {
                                p = x * a[i][k] + y * a[i][k + 1];
                                if (k != nn - 1) { p += z * a[i][k + 2]; a[i][k + 2] -= p * r; }
                                a[i][k + 1] -= p * q;
                                a[i][k] -= p;
                            }
                        }
                    }
                }
            }
        } while (l < nn - 1);
    }
    return true;
}

// null vector of the (numerically singular) 10 x 10 matrix a: elimination with complete pivoting, the last pivot is taken as zero
__host__ __device__ inline void null_vector10(double (&a)[10][10], double (&v)[10]) {
    int perm[10];
    for (int i = 0; i < 10; ++i) perm[i] = i;
    for (int s = 0; s < 9; ++s) {
        int pi = s, pj = s;
        double best = -1.0;
        for (int i = s; i < 10; ++i)
            for (int j = s; j < 10; ++j)
                if (fabs(a[i][j]) > best) { best = fabs(a[i][j]); pi = i; pj = j; }
        if (pi != s) for (int j = 0; j < 10; ++j) { const double t = a[pi][j]; a[pi][j] = a[s][j]; a[s][j] = t; }
        if (pj != s) {
            for (int i = 0; i < 10; ++i) { const double t = a[i][pj]; a[i][pj] = a[i][s]; a[i][s] = t; }
            const int t = perm[pj]; perm[pj] = perm[s]; perm[s] = t;
        }
        const double d = a[s][s];
        if (d == 0.0) continue;
        for (int i = s + 1; i < 10; ++i) {
            const double f = a[i][s] / d;
            if (f != 0.0) for (int j = s; j < 10; ++j) a[i][j] -= f * a[s][j];
        }
    }
    double y[10];
    y[9] = 1.0;
    for (int s = 8; s >= 0; --s) {
        double acc = 0.0;
        for (int j = s + 1; j < 10; ++j) acc += a[s][j] * y[j];
        y[s] = a[s][s] != 0.0 ? -acc / a[s][s] : 0.0;
    }
    for (int s = 0; s < 10; ++s) v[perm[s]] = y[s];
}

// All real essential matrices through five correspondences in normalised coordinates (q2_h^T E q1_h = 0), each scaled to unit Frobenius
// norm, in ascending order of the leading 2 x 2 minor (canonical: the oracle uses another null-space basis).  E row-major.  Returns the
// number of solutions (<= 10).
__host__ __device__ inline int five_point(const double (&q1)[5][2], const double (&q2)[5][2], double (&Es)[10][9]) {
    // ---- null space of the 5 x 9 system ----
    double A[5][9];
    for (int i = 0; i < 5; ++i) {
        const double a[3] = {q2[i][0], q2[i][1], 1.0}, b[3] = {q1[i][0], q1[i][1], 1.0};
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) A[i][3 * r + c] = a[r] * b[c];
    }
    int perm[9];
    for (int j = 0; j < 9; ++j) perm[j] = j;
    for (int s = 0; s < 5; ++s) {
        int pi = s, pj = s;
        double best = -1.0;
        for (int i = s; i < 5; ++i)
            for (int j = s; j < 9; ++j)
                if (fabs(A[i][j]) > best) { best = fabs(A[i][j]); pi = i; pj = j; }
        if (best <= 0.0) return 0;
        if (pi != s) for (int j = 0; j < 9; ++j) { const double t = A[pi][j]; A[pi][j] = A[s][j]; A[s][j] = t; }
        if (pj != s) {
            for (int i = 0; i < 5; ++i) { const double t = A[i][pj]; A[i][pj] = A[i][s]; A[i][s] = t; }
            const int t = perm[pj]; perm[pj] = perm[s]; perm[s] = t;
        }
        const double d = 1.0 / A[s][s];
        for (int j = 0; j < 9; ++j) A[s][j] *= d;
        for (int i = 0; i < 5; ++i) {
            if (i == s) continue;
            const double f = A[i][s];
            if (f != 0.0) for (int j = 0; j < 9; ++j) A[i][j] -= f * A[s][j];
        }
    }
    double B[4][9];                                   // basis vector k: free column 5 + k = 1, pivot columns = - reduced entries
    for (int k = 0; k < 4; ++k) {
        for (int j = 0; j < 9; ++j) B[k][j] = 0.0;
        B[k][perm[5 + k]] = 1.0;
        for (int s = 0; s < 5; ++s) B[k][perm[s]] = -A[s][5 + k];
    }
    for (int rep = 0; rep < 2; ++rep)                 // modified Gram-Schmidt, twice
        for (int k = 0; k < 4; ++k) {
            for (int p = 0; p < k; ++p) {
                double d = 0.0;
                for (int j = 0; j < 9; ++j) d += B[k][j] * B[p][j];
                for (int j = 0; j < 9; ++j) B[k][j] -= d * B[p][j];
            }
            double nrm = 0.0;
            for (int j = 0; j < 9; ++j) nrm += B[k][j] * B[k][j];
            nrm = 1.0 / sqrt(nrm);
            for (int j = 0; j < 9; ++j) B[k][j] *= nrm;
        }
    // ---- the ten cubic constraints: E(x, y, z) = x B0 + y B1 + z B2 + B3, entries as linear polynomials [x, y, z, 1] ----
    double El[9][4];
    for (int e = 0; e < 9; ++e)
        for (int k = 0; k < 4; ++k) El[e][k] = B[k][e];
    double EEt[6][10];                                // (0,0) (0,1) (0,2) (1,1) (1,2) (2,2)
    {
        int idx = 0;
        for (int r = 0; r < 3; ++r)
            for (int c = r; c < 3; ++c, ++idx) {
                double acc[10], t[10];
                for (int k = 0; k < 10; ++k) acc[k] = 0.0;
                for (int k = 0; k < 3; ++k) {
                    mul_ll(El[3 * r + k], El[3 * c + k], t);
                    for (int u = 0; u < 10; ++u) acc[u] += t[u];
                }
                for (int u = 0; u < 10; ++u) EEt[idx][u] = acc[u];
            }
    }
    const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    double tr[10];
    for (int u = 0; u < 10; ++u) tr[u] = EEt[0][u] + EEt[3][u] + EEt[5][u];
    double M[10][20];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double* row = M[3 * r + c];
            for (int u = 0; u < 20; ++u) row[u] = 0.0;
            for (int k = 0; k < 3; ++k) fma_ql(EEt[sym[r][k]], El[3 * k + c], 2.0, row);
            fma_ql(tr, El[3 * r + c], -1.0, row);
        }
    {
        double* row = M[9];
        for (int u = 0; u < 20; ++u) row[u] = 0.0;
        double m0[10], m1[10], m2[10], t[10];
        mul_ll(El[4], El[8], m0); mul_ll(El[5], El[7], t); for (int u = 0; u < 10; ++u) m0[u] -= t[u];     // E11 E22 - E12 E21
        mul_ll(El[3], El[8], m1); mul_ll(El[5], El[6], t); for (int u = 0; u < 10; ++u) m1[u] -= t[u];     // E10 E22 - E12 E20
        mul_ll(El[3], El[7], m2); mul_ll(El[4], El[6], t); for (int u = 0; u < 10; ++u) m2[u] -= t[u];     // E10 E21 - E11 E20
        fma_ql(m0, El[0], 1.0, row);
        fma_ql(m1, El[1], -1.0, row);
        fma_ql(m2, El[2], 1.0, row);
    }
    // ---- Gauss-Jordan on the cubic block (partial pivoting): cubic_i = - sum_j M[i][10 + j] basis_j ----
    for (int s = 0; s < 10; ++s) {
        int pi = s;
        double best = fabs(M[s][s]);
        for (int i = s + 1; i < 10; ++i)
            if (fabs(M[i][s]) > best) { best = fabs(M[i][s]); pi = i; }
        if (best == 0.0) return 0;
        if (pi != s) for (int j = 0; j < 20; ++j) { const double t = M[pi][j]; M[pi][j] = M[s][j]; M[s][j] = t; }
        const double d = 1.0 / M[s][s];
        for (int j = s; j < 20; ++j) M[s][j] *= d;
        for (int i = 0; i < 10; ++i) {
            if (i == s) continue;
            const double f = M[i][s];
            if (f != 0.0) for (int j = s; j < 20; ++j) M[i][j] -= f * M[s][j];
        }
    }
    // ---- multiplication by x in the basis (x2, xy, xz, y2, yz, z2, x, y, z, 1) ----
    double Ax[10][10];
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 10; ++j) Ax[i][j] = 0.0;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 10; ++j) Ax[i][j] = -M[i][10 + j];
    Ax[6][0] = Ax[7][1] = Ax[8][2] = Ax[9][6] = 1.0;
    double H[kEig + 1][kEig + 1], wr[kEig + 1], wi[kEig + 1];
    for (int i = 0; i < 10; ++i)
        for (int j = 0; j < 10; ++j) H[i + 1][j + 1] = Ax[i][j];
    elmhes(H);
    if (!hqr(H, wr, wi)) return 0;
    // real eigenvalues, ascending
    double lam[10];
    int nl = 0;
    for (int k = 1; k <= 10; ++k) {
        if (fabs(wi[k]) > 1e-9 * fmax(1.0, fabs(wr[k]))) continue;
        int p = nl++;
        while (p > 0 && lam[p - 1] > wr[k]) { lam[p] = lam[p - 1]; --p; }
        lam[p] = wr[k];
    }
    int ns = 0;
    for (int k = 0; k < nl; ++k) {
        double S[10][10], v[10];
        for (int i = 0; i < 10; ++i)
            for (int j = 0; j < 10; ++j) S[i][j] = Ax[i][j] - (i == j ? lam[k] : 0.0);
        null_vector10(S, v);
        if (fabs(v[9]) < 1e-14 * fmax(fmax(fabs(v[6]), fabs(v[7])), fmax(fabs(v[8]), 1e-300))) continue;
        const double x = v[6] / v[9], y = v[7] / v[9], z = v[8] / v[9];
        double nrm = 0.0;
        for (int e = 0; e < 9; ++e) {
            Es[ns][e] = x * B[0][e] + y * B[1][e] + z * B[2][e] + B[3][e];
            nrm += Es[ns][e] * Es[ns][e];
        }
        if (!(nrm > 0.0) || !(nrm < 1e300)) continue;
        nrm = 1.0 / sqrt(nrm);
        for (int e = 0; e < 9; ++e) Es[ns][e] *= nrm;
        ++ns;
    }
    // canonical order, independent of the null-space basis and of the sign of E: ascending leading 2 x 2 minor (insertion sort)
    for (int a = 1; a < ns; ++a) {
        double cur[9];
        for (int e = 0; e < 9; ++e) cur[e] = Es[a][e];
        const double key = cur[0] * cur[4] - cur[1] * cur[3];
        int b = a;
        while (b > 0 && Es[b - 1][0] * Es[b - 1][4] - Es[b - 1][1] * Es[b - 1][3] > key) {
            for (int e = 0; e < 9; ++e) Es[b][e] = Es[b - 1][e];
            --b;
        }
        for (int e = 0; e < 9; ++e) Es[b][e] = cur[e];
    }
    return ns;
}

// OpenCV's EMEstimatorCallback::computeError: (x2^T E x1)^2 / (|E x1|_xy^2 + |E^T x2|_xy^2), normalised coordinates
__host__ __device__ inline double sampson_error(const double* E, double x1, double y1, double x2, double y2) {
    const double a0 = E[0] * x1 + E[1] * y1 + E[2], a1 = E[3] * x1 + E[4] * y1 + E[5], a2 = E[6] * x1 + E[7] * y1 + E[8];
    const double b0 = E[0] * x2 + E[3] * y2 + E[6], b1 = E[1] * x2 + E[4] * y2 + E[7];
    const double d = x2 * a0 + y2 * a1 + a2;
    return d * d / (a0 * a0 + a1 * a1 + b0 * b0 + b1 * b1);
}

// OpenCV's RANSACUpdateNumIters
__host__ __device__ inline int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.0), 1.0);
    ep = fmin(fmax(ep, 0.0), 1.0);
    double num = fmax(1.0 - p, 2.2250738585072014e-308);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num);
    denom = log(denom);
    if (denom >= 0.0 || -num >= max_iters * (-denom)) return max_iters;
    return (int)floor(num / denom + 0.5);
}

// ---- symmetric eigenproblems by cyclic Jacobi (N = 3, 4): a destroyed, eigenvalues in d, eigenvectors in the COLUMNS of v ----
template <int N>
__host__ __device__ inline void jacobi_sym(double (&a)[N][N], double (&d)[N], double (&v)[N][N]) {
    for (int i = 0; i < N; ++i) {
        for (int j = 0; j < N; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; ++i) {
            diag += fabs(a[i][i]);
            for (int j = i + 1; j < N; ++j) off += fabs(a[i][j]);
        }
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                if (a[p][q] == 0.0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) {       // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {       // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < N; ++i) d[i] = a[i][i];
}

// cv::decomposeEssentialMat: E = U diag(s, s, 0) V^T with det U = det V = +1; R1 = U W V^T, R2 = U W^T V^T, t = third column of U
__host__ __device__ inline void decompose_essential(const double* E, double* R1, double* R2, double* t) {
    double ata[3][3], d[3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) ata[i][j] = E[0 + i] * E[0 + j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    jacobi_sym<3>(ata, d, V);
    int o[3] = {0, 1, 2};                              // descending eigenvalues
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (d[o[j]] > d[o[i]]) { const int s = o[i]; o[i] = o[j]; o[j] = s; }
    double v[3][3], u[3][3];                           // v[k] = k-th right singular vector, u[k] = k-th left one
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) v[k][i] = V[i][o[k]];
    for (int k = 0; k < 2; ++k) {
        for (int i = 0; i < 3; ++i) u[k][i] = E[3 * i] * v[k][0] + E[3 * i + 1] * v[k][1] + E[3 * i + 2] * v[k][2];
        if (k == 1) {
            const double dp = u[1][0] * u[0][0] + u[1][1] * u[0][1] + u[1][2] * u[0][2];
            for (int i = 0; i < 3; ++i) u[1][i] -= dp * u[0][i];
        }
        const double n = 1.0 / sqrt(u[k][0] * u[k][0] + u[k][1] * u[k][1] + u[k][2] * u[k][2]);
        for (int i = 0; i < 3; ++i) u[k][i] *= n;
    }
    u[2][0] = u[0][1] * u[1][2] - u[0][2] * u[1][1];
    u[2][1] = u[0][2] * u[1][0] - u[0][0] * u[1][2];
    u[2][2] = u[0][0] * u[1][1] - u[0][1] * u[1][0];
    // v3 = v1 x v2 makes det V = +1 (its sign is free: the third singular value is zero)
    v[2][0] = v[0][1] * v[1][2] - v[0][2] * v[1][1];
    v[2][1] = v[0][2] * v[1][0] - v[0][0] * v[1][2];
    v[2][2] = v[0][0] * v[1][1] - v[0][1] * v[1][0];
    // U W = [u2, -u1, u3] (columns), U W^T = [-u2, u1, u3]
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            R1[3 * i + j] = u[1][i] * v[0][j] - u[0][i] * v[1][j] + u[2][i] * v[2][j];
            R2[3 * i + j] = -u[1][i] * v[0][j] + u[0][i] * v[1][j] + u[2][i] * v[2][j];
        }
    for (int i = 0; i < 3; ++i) t[i] = u[2][i];
}

// cv::triangulatePoints for one correspondence with P1 = [I | 0], P2 = [R | t] (normalised coordinates): the smallest right singular
// vector of the 4 x 4 DLT system, as the smallest eigenvector of A^T A.  X = homogeneous point.
__host__ __device__ inline void triangulate_dlt(const double* R, const double* t, double ax, double ay, double bx, double by, double (&X)[4]) {
    double A[4][4] = {{-1.0, 0.0, ax, 0.0},
                      {0.0, -1.0, ay, 0.0},
                      {bx * R[6] - R[0], bx * R[7] - R[1], bx * R[8] - R[2], bx * t[2] - t[0]},
                      {by * R[6] - R[3], by * R[7] - R[4], by * R[8] - R[5], by * t[2] - t[1]}};
    double ata[4][4], d[4], V[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) ata[i][j] = A[0][i] * A[0][j] + A[1][i] * A[1][j] + A[2][i] * A[2][j] + A[3][i] * A[3][j];
    jacobi_sym<4>(ata, d, V);
    int k = 0;
    for (int i = 1; i < 4; ++i)
        if (d[i] < d[k]) k = i;
    for (int i = 0; i < 4; ++i) X[i] = V[i][k];
}

// the per-point test of cv::recoverPose for one (R, t) candidate: in front of both cameras and nearer than `dist`
__host__ __device__ inline bool cheirality_ok(const double* R, const double* t, double ax, double ay, double bx, double by, double dist) {
    double Q[4];
    triangulate_dlt(R, t, ax, ay, bx, by, Q);
    if (!(Q[2] * Q[3] > 0.0)) return false;
    const double x = Q[0] / Q[3], y = Q[1] / Q[3], z = Q[2] / Q[3];
    if (!(z < dist)) return false;
    const double z2 = R[6] * x + R[7] * y + R[8] * z + t[2];
    return z2 > 0.0 && z2 < dist;
}

}  // namespace tv
}  // namespace esfm
