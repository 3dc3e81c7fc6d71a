// TEST INFRASTRUCTURE -- part of the CPU oracle (see oracle/README.md).  Never linked into the product.
//
// Minimal dependency-free dense FP64 linear algebra that restates the Eigen operations the
// reference's hot path calls (Eigen itself is not available offline):
//   * Mat                 : column-major dynamic matrix (Eigen::MatrixXd layout, reference
//                           tests/dwbc_test_util.h:23-26 dumps .data() directly)
//   * llt_solve_identity  : A.llt().solve(I)                      reference src/dwbc.cpp:307
//   * lu_inverse          : MatrixXd::inverse() = PartialPivLU    reference src/wbd.cpp:115,128,210
//   * cod_pinv            : CompleteOrthogonalDecomposition with setThreshold(COD_THRESHOLD),
//                           pseudoInverse(), rank(), householderQ()   reference src/wbd.cpp:5-53
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
    static Mat identity(int n) {
        Mat m(n, n);
        for (int i = 0; i < n; i++) m(i, i) = 1.0;
        return m;
    }
    void set_zero(int r_, int c_) {
        r = r_; c = c_;
        d.assign((size_t)r_ * c_, 0.0);
    }
    inline double &operator()(int i, int j) { return d[(size_t)j * r + i]; }
    inline double operator()(int i, int j) const { return d[(size_t)j * r + i]; }
    inline double &operator[](int i) { return d[i]; }
    inline double operator[](int i) const { return d[i]; }
    int size() const { return r * c; }
    Mat block(int i0, int j0, int nr, int nc) const {
        Mat m(nr, nc);
        for (int j = 0; j < nc; j++)
            for (int i = 0; i < nr; i++) m(i, j) = (*this)(i0 + i, j0 + j);
        return m;
    }
    void set_block(int i0, int j0, const Mat &b) {
        for (int j = 0; j < b.c; j++)
            for (int i = 0; i < b.r; i++) (*this)(i0 + i, j0 + j) = b(i, j);
    }
    Mat t() const {
        Mat m(c, r);
        for (int j = 0; j < c; j++)
            for (int i = 0; i < r; i++) m(j, i) = (*this)(i, j);
        return m;
    }
    Mat right_cols(int n) const { return block(0, c - n, r, n); }
    Mat left_cols(int n) const { return block(0, 0, r, n); }
    Mat top_rows(int n) const { return block(0, 0, n, c); }
    Mat bottom_rows(int n) const { return block(r - n, 0, n, c); }
    double norm() const {
        double s = 0;
        for (double v : d) s += v * v;
        return std::sqrt(s);
    }
};

inline Mat operator*(const Mat &a, const Mat &b) {
    assert(a.c == b.r);
    Mat m(a.r, b.c);
    for (int j = 0; j < b.c; j++)
        for (int k = 0; k < a.c; k++) {
            double bkj = b(k, j);
            if (bkj == 0.0) continue;
            const double *ak = &a.d[(size_t)k * a.r];
            double *mj = &m.d[(size_t)j * m.r];
            for (int i = 0; i < a.r; i++) mj[i] += ak[i] * bkj;
        }
    return m;
}
inline Mat operator+(const Mat &a, const Mat &b) {
    assert(a.r == b.r && a.c == b.c);
    Mat m(a.r, a.c);
    for (size_t i = 0; i < m.d.size(); i++) m.d[i] = a.d[i] + b.d[i];
    return m;
}
inline Mat operator-(const Mat &a, const Mat &b) {
    assert(a.r == b.r && a.c == b.c);
    Mat m(a.r, a.c);
    for (size_t i = 0; i < m.d.size(); i++) m.d[i] = a.d[i] - b.d[i];
    return m;
}
inline Mat operator*(double s, const Mat &a) {
    Mat m(a.r, a.c);
    for (size_t i = 0; i < m.d.size(); i++) m.d[i] = s * a.d[i];
    return m;
}
inline Mat operator-(const Mat &a) { return (-1.0) * a; }

// A.llt().solve(Identity): Cholesky A = L L^T, then forward/back substitution per column.
inline bool llt_solve_identity(const Mat &A, Mat &Ainv) {
    int n = A.r;
    Mat L = A;
    for (int j = 0; j < n; j++) {
        double s = L(j, j);
        for (int k = 0; k < j; k++) s -= L(j, k) * L(j, k);
        if (!(s > 0)) return false;
        double ljj = std::sqrt(s);
        L(j, j) = ljj;
        for (int i = j + 1; i < n; i++) {
            double v = L(i, j);
            for (int k = 0; k < j; k++) v -= L(i, k) * L(j, k);
            L(i, j) = v / ljj;
        }
    }
    Ainv.set_zero(n, n);
    std::vector<double> y(n);
    for (int col = 0; col < n; col++) {
        for (int i = 0; i < n; i++) {
            double v = (i == col) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) v -= L(i, k) * y[k];
            y[i] = v / L(i, i);
        }
        for (int i = n - 1; i >= 0; i--) {
            double v = y[i];
            for (int k = i + 1; k < n; k++) v -= L(k, i) * Ainv(k, col);
            Ainv(i, col) = v / L(i, i);
        }
    }
    return true;
}

// MatrixXd::inverse() on a dynamic matrix = PartialPivLU(A).inverse().
inline Mat lu_inverse(const Mat &A) {
    int n = A.r;
    assert(A.c == n);
    Mat LU = A;
    std::vector<int> piv(n);
    for (int i = 0; i < n; i++) piv[i] = i;
    for (int k = 0; k < n; k++) {
        int p = k;
        double best = std::fabs(LU(k, k));
        for (int i = k + 1; i < n; i++)
            if (std::fabs(LU(i, k)) > best) best = std::fabs(LU(i, k)), p = i;
        if (p != k) {
            for (int j = 0; j < n; j++) std::swap(LU(k, j), LU(p, j));
            std::swap(piv[k], piv[p]);
        }
        double d = LU(k, k);
        for (int i = k + 1; i < n; i++) {
            LU(i, k) /= d;
            double l = LU(i, k);
            if (l != 0.0)
                for (int j = k + 1; j < n; j++) LU(i, j) -= l * LU(k, j);
        }
    }
    Mat inv(n, n);
    std::vector<double> y(n);
    for (int col = 0; col < n; col++) {
        for (int i = 0; i < n; i++) {
            double v = (piv[i] == col) ? 1.0 : 0.0;
            for (int k = 0; k < i; k++) v -= LU(i, k) * y[k];
            y[i] = v;
        }
        for (int i = n - 1; i >= 0; i--) {
            double v = y[i];
            for (int k = i + 1; k < n; k++) v -= LU(i, k) * inv(k, col);
            inv(i, col) = v / LU(i, i);
        }
    }
    return inv;
}

// Householder reflector for x (length n): H = I - tau v v^T with v[0] = 1, H x = beta e0.
inline void make_householder(double *x, int n, double &tau, double &beta) {
    double tail = 0;
    for (int i = 1; i < n; i++) tail += x[i] * x[i];
    double c0 = x[0];
    if (tail <= 1e-300) {
        tau = 0;
        beta = c0;
        for (int i = 1; i < n; i++) x[i] = 0;
        return;
    }
    beta = std::sqrt(c0 * c0 + tail);
    if (c0 >= 0) beta = -beta;
    for (int i = 1; i < n; i++) x[i] /= (c0 - beta);
    tau = (beta - c0) / beta;
}

struct Cod {
    int rows = 0, cols = 0, rank = 0;
    Mat Q;                  // rows x rows, householderQ()
    Mat R;                  // rows x cols upper trapezoidal, A P = Q R
    std::vector<int> perm;  // column k of A*P is column perm[k] of A
    Mat pinv;               // cols x rows
};

// Column-pivoted Householder QR + rank decision |R_ii| > threshold * max|R_ii| (Eigen
// ColPivHouseholderQR::rank()), then the complete orthogonal decomposition of the leading
// `rank` rows to get the minimum-norm pseudo-inverse.  `force_rank` > 0 overrides the rank
// used for the returned null-space rows only in the caller (reference src/wbd.cpp:39-51);
// the pseudo-inverse always uses the detected rank, as Eigen's pseudoInverse() does.
inline Cod cod_compute(const Mat &A, double threshold) {
    Cod f;
    int m = A.r, n = A.c;
    f.rows = m; f.cols = n;
    Mat R = A;
    f.perm.resize(n);
    for (int j = 0; j < n; j++) f.perm[j] = j;
    int steps = std::min(m, n);
    std::vector<double> taus(steps, 0.0);
    std::vector<double> w(std::max(m, n));
    for (int k = 0; k < steps; k++) {
        int p = k;
        double best = -1;
        for (int j = k; j < n; j++) {
            double s = 0;
            for (int i = k; i < m; i++) s += R(i, j) * R(i, j);
            if (s > best) best = s, p = j;
        }
        if (p != k) {
            for (int i = 0; i < m; i++) std::swap(R(i, k), R(i, p));
            std::swap(f.perm[k], f.perm[p]);
        }
        double tau, beta;
        make_householder(&R(k, k), m - k, tau, beta);
        taus[k] = tau;
        // apply H to the trailing columns
        for (int j = k + 1; j < n; j++) {
            double s = R(k, j);
            for (int i = k + 1; i < m; i++) s += R(i, k) * R(i, j);
            s *= tau;
            R(k, j) -= s;
            for (int i = k + 1; i < m; i++) R(i, j) -= s * R(i, k);
        }
        R(k, k) = beta;
    }
    // Q = H_0 H_1 ... H_{steps-1}
    Mat Q = Mat::identity(m);
    for (int k = steps - 1; k >= 0; k--) {
        double tau = taus[k];
        if (tau == 0) continue;
        for (int j = k; j < m; j++) {
            double s = Q(k, j);
            for (int i = k + 1; i < m; i++) s += R(i, k) * Q(i, j);
            s *= tau;
            Q(k, j) -= s;
            for (int i = k + 1; i < m; i++) Q(i, j) -= s * R(i, k);
        }
    }
    for (int k = 0; k < steps; k++)
        for (int i = k + 1; i < m; i++) R(i, k) = 0;
    double maxpiv = 0;
    for (int k = 0; k < steps; k++) maxpiv = std::max(maxpiv, std::fabs(R(k, k)));
    int rank = 0;
    for (int k = 0; k < steps; k++)
        if (std::fabs(R(k, k)) > threshold * maxpiv) rank++;
    f.rank = rank;
    f.Q = Q;
    f.R = R;
    // minimum-norm pseudo-inverse: Rt = R[0:rank, :] (rank x n) = U^T Z1^T with Z1 (n x rank)
    // orthonormal and U upper triangular (QR of Rt^T); pinv = P Z1 U^{-T} Q1^T.
    f.pinv.set_zero(n, m);
    if (rank == 0) return f;
    Mat T(n, rank);
    for (int i = 0; i < rank; i++)
        for (int j = 0; j < n; j++) T(j, i) = R(i, j);
    std::vector<double> tz(rank, 0.0);
    for (int k = 0; k < rank; k++) {
        double tau, beta;
        make_householder(&T(k, k), n - k, tau, beta);
        tz[k] = tau;
        for (int j = k + 1; j < rank; j++) {
            double s = T(k, j);
            for (int i = k + 1; i < n; i++) s += T(i, k) * T(i, j);
            s *= tau;
            T(k, j) -= s;
            for (int i = k + 1; i < n; i++) T(i, j) -= s * T(i, k);
        }
        T(k, k) = beta;
    }
    // Y = U^{-T} Q1^T  (rank x m): solve U^T Y = Q1^T, U^T lower triangular
    Mat Y(rank, m);
    for (int col = 0; col < m; col++) {
        for (int i = 0; i < rank; i++) {
            double v = Q(col, i);
            for (int k = 0; k < i; k++) v -= T(k, i) * Y(k, col);
            Y(i, col) = v / T(i, i);
        }
    }
    // X = Z1 Y = H_0 ... H_{rank-1} [Y; 0]
    Mat X(n, m);
    for (int col = 0; col < m; col++)
        for (int i = 0; i < rank; i++) X(i, col) = Y(i, col);
    for (int k = rank - 1; k >= 0; k--) {
        double tau = tz[k];
        if (tau == 0) continue;
        for (int col = 0; col < m; col++) {
            double s = X(k, col);
            for (int i = k + 1; i < n; i++) s += T(i, k) * X(i, col);
            s *= tau;
            X(k, col) -= s;
            for (int i = k + 1; i < n; i++) X(i, col) -= s * T(i, k);
        }
    }
    for (int k = 0; k < n; k++)
        for (int col = 0; col < m; col++) f.pinv(f.perm[k], col) = X(k, col);
    return f;
}

}  // namespace orc
