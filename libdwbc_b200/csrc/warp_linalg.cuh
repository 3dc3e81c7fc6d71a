// Warp-cooperative dense FP64 linear algebra on small column-major matrices (n <= ~50) held in
// shared memory.  One warp owns one robot state, so every routine here is called by all 32 lanes
// of a warp with identical arguments and only uses full-mask warp primitives.
//
// Layout rules that keep shared-memory accesses conflict-free for 64-bit words:
//   * leading dimensions are ODD (a lane-per-column sweep then strides by an odd number of doubles,
//     which touches 16 distinct bank pairs per half-warp),
//   * loops over the contiguous index are ALIGNED across lanes (every lane visits the same row at the
//     same time, predicated if it does not need it), so the shared operand is a broadcast.
#pragma once
#include "cuda_compat.h"

namespace dwbc {

__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31); }
__device__ __forceinline__ int odd_ld(int n) { return n | 1; }

__device__ __forceinline__ double warp_sum(double v) {
    DWBC_UNROLL(5)
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DWBC_FULL, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
    DWBC_UNROLL(5)
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(DWBC_FULL, v, o);
        v = w > v ? w : v;
    }
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
    DWBC_UNROLL(5)
    for (int o = 16; o > 0; o >>= 1) {
        double w = __shfl_xor_sync(DWBC_FULL, v, o);
        v = w < v ? w : v;
    }
    return v;
}

// Cooperative copy of a column-major rows x cols block (global or shared on either side).
__device__ __forceinline__ void wcopy(int rows, int cols, const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd) {
    const int lane = lane_id();
    if (lds == rows && ldd == rows) {
        for (int e = lane; e < rows * cols; e += 32) dst[e] = src[e];
    } else {
        for (int e = lane; e < rows * cols; e += 32) {
            int j = e / rows, i = e - j * rows;
            dst[i + j * ldd] = src[i + j * lds];
        }
    }
}
__device__ __forceinline__ void wfill(int count, double *dst, double v) {
    for (int e = lane_id(); e < count; e += 32) dst[e] = v;
}

// C(i,j) <- epi(i, j, sum_k opA(i,k) * opB(k,j)) for an M x N result.
//   opA(i,k) = TA ? A[k + i*lda] : A[i + k*lda],  opB(k,j) = TB ? B[j + k*ldb] : B[k + j*ldb]
// Every lane accumulates TR x TC outputs in registers.  Tiles are CYCLIC (rows ti + r*tr, columns
// tj + c*tc) so that, for a fixed register index, consecutive lanes touch consecutive rows.
template <int TR, int TC, bool TA, bool TB, class Epi>
__device__ __forceinline__ void wgemm(int M, int N, int K, const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb,
                                      Epi epi) {
    const int lane = lane_id();
    const int tr = (M + TR - 1) / TR, tc = (N + TC - 1) / TC;
    for (int tile = lane; tile < tr * tc; tile += 32) {
        const int tj = tile / tr, ti = tile - tj * tr;
        int ia[TR], jb[TC];
        DWBC_UNROLL(TR)
        for (int r = 0; r < TR; r++) {
            int i = ti + r * tr;
            ia[r] = (i < M ? i : ti) * (TA ? lda : 1);
        }
        DWBC_UNROLL(TC)
        for (int c = 0; c < TC; c++) {
            int j = tj + c * tc;
            jb[c] = (j < N ? j : tj) * (TB ? 1 : ldb);
        }
        double acc[TR][TC];
        DWBC_UNROLL(TR)
        for (int r = 0; r < TR; r++) {
            DWBC_UNROLL(TC)
            for (int c = 0; c < TC; c++) acc[r][c] = 0.0;
        }
        const int sa = TA ? 1 : lda, sb = TB ? ldb : 1;
        DWBC_UNROLL(2)
        for (int k = 0; k < K; k++) {
            double a[TR], b[TC];
            DWBC_UNROLL(TR)
            for (int r = 0; r < TR; r++) a[r] = A[ia[r] + k * sa];
            DWBC_UNROLL(TC)
            for (int c = 0; c < TC; c++) b[c] = B[jb[c] + k * sb];
            DWBC_UNROLL(TR)
            for (int r = 0; r < TR; r++) {
                DWBC_UNROLL(TC)
                for (int c = 0; c < TC; c++) acc[r][c] = fma(a[r], b[c], acc[r][c]);
            }
        }
        DWBC_UNROLL(TR)
        for (int r = 0; r < TR; r++) {
            DWBC_UNROLL(TC)
            for (int c = 0; c < TC; c++) {
                int i = ti + r * tr, j = tj + c * tc;
                if (i < M && j < N) epi(i, j, acc[r][c]);
            }
        }
    }
}

// y(i) <- epi(i, sum_k opA(i,k) x(k)), one row per lane (M may exceed 32).
template <bool TA, class Epi>
__device__ __forceinline__ void wgemv(int M, int K, const double *__restrict__ A, int lda, const double *__restrict__ x, Epi epi) {
    for (int i = lane_id(); i < M; i += 32) {
        double a0 = 0, a1 = 0;
        int k = 0;
        for (; k + 1 < K; k += 2) {
            a0 = fma(TA ? A[k + i * lda] : A[i + k * lda], x[k], a0);
            a1 = fma(TA ? A[k + 1 + i * lda] : A[i + (k + 1) * lda], x[k + 1], a1);
        }
        if (k < K) a0 = fma(TA ? A[k + i * lda] : A[i + k * lda], x[k], a0);
        epi(i, a0 + a1);
    }
}

// In-place lower Cholesky A = L L^T of the SPD matrix in the lower triangle of A (n x n, ld odd).
// rdiag[i] receives 1 / L(i,i).  Returns the smallest pivot seen (<= 0 means "not SPD"; the
// factorisation then continues with a clamped pivot so that no NaN is produced).
__device__ __forceinline__ double wchol_lower(int n, double *A, int ld, double *rdiag) {
    const int lane = lane_id();
    double minpiv = 1e300;
    for (int k = 0; k < n; k++) {
        double akk = A[k + k * ld];
        minpiv = akk < minpiv ? akk : minpiv;
        if (!(akk > 1e-300)) akk = 1e-300;
        const double lkk = sqrt(akk), r = 1.0 / lkk;
        __syncwarp();
        for (int i = k + 1 + lane; i < n; i += 32) A[i + k * ld] *= r;
        if (lane == 0) {
            A[k + k * ld] = lkk;
            rdiag[k] = r;
        }
        __syncwarp();
        // trailing update, one column per lane, rows aligned across lanes (broadcast of column k)
        for (int j0 = k + 1; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            const bool act = j < n;
            const double ljk = act ? A[j + k * ld] : 0.0;
            double *cj = A + (act ? j : k + 1) * ld;
            for (int i = j0; i < n; i++) {
                const double lik = A[i + k * ld];
                if (act && i >= j) cj[i] = fma(-lik, ljk, cj[i]);
            }
        }
        __syncwarp();
    }
    return minpiv;
}

// X <- A^{-1} (full, symmetric) from the Cholesky factor L (lower triangle of L, ld odd) and rdiag.
// One column per lane: forward then backward substitution on a unit vector, all lanes walking the
// same rows so that the L operand is a broadcast.
__device__ __forceinline__ void wchol_inverse(int n, const double *L, int ld, const double *rdiag, double *X, int ldx) {
    const int lane = lane_id();
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const bool act = j < n;
        double *xj = X + (act ? j : j0) * ldx;  // idle lanes shadow column j0 without storing
        // forward: y = L^{-1} e_j ; y_k = 0 for k < j comes out of the arithmetic (no predicate needed)
        for (int i = j0; i < n; i++) {
            double a0 = (i == j) ? 1.0 : 0.0, a1 = 0.0;
            int k = j0;
            for (; k + 1 < i; k += 2) {
                a0 = fma(-L[i + k * ld], xj[k], a0);
                a1 = fma(-L[i + (k + 1) * ld], xj[k + 1], a1);
            }
            if (k < i) a0 = fma(-L[i + k * ld], xj[k], a0);
            const double v = (a0 + a1) * rdiag[i];
            if (act) xj[i] = v;
        }
        // backward: x = L^{-T} y, rows n-1 .. j0
        for (int i = n - 1; i >= j0; i--) {
            double a0 = xj[i], a1 = 0.0;
            int k = i + 1;
            for (; k + 1 < n; k += 2) {
                a0 = fma(-L[k + i * ld], xj[k], a0);
                a1 = fma(-L[k + 1 + i * ld], xj[k + 1], a1);
            }
            if (k < n) a0 = fma(-L[k + i * ld], xj[k], a0);
            const double v = (a0 + a1) * rdiag[i];
            if (act) xj[i] = v;
        }
        __syncwarp();
    }
    // rows above the block diagonal of later column blocks come from symmetry
    for (int j0 = 32; j0 < n; j0 += 32) {
        for (int j = j0; j < n; j++)
            for (int i = lane; i < j0; i += 32) X[i + j * ldx] = X[j + i * ldx];
    }
    __syncwarp();
}

// Gauss-Jordan with partial pivoting on the augmented matrix [M | R] (n x (n + nr), column-major,
// ld odd, n + nr <= 32 columns, one column per lane): on return R holds M^{-1} R.  Returns the
// smallest pivot magnitude relative to the largest (0 => singular).
__device__ __forceinline__ double wgauss_jordan(int n, int nr, double *Aug, int ld) {
    const int lane = lane_id();
    const int cols = n + nr;
    const bool act = lane < cols;
    double *col = Aug + (act ? lane : 0) * ld;
    double pmax = 0.0, pmin = 1e300;
    for (int k = 0; k < n; k++) {
        // pivot search in column k (all lanes redundantly; broadcast reads)
        int p = k;
        double best = fabs(Aug[k + k * ld]);
        for (int i = k + 1; i < n; i++) {
            double v = fabs(Aug[i + k * ld]);
            if (v > best) best = v, p = i;
        }
        pmax = best > pmax ? best : pmax;
        pmin = best < pmin ? best : pmin;
        const double piv = Aug[p + k * ld];
        const double rp = 1.0 / (piv != 0.0 ? piv : 1e-300);
        __syncwarp();
        // every lane handles its own column: swap rows k and p, scale row k, eliminate
        if (act) {
            double t = col[p];
            col[p] = col[k];
            col[k] = t;
        }
        __syncwarp();
        if (act && lane > k) {
            const double rk = col[k] * rp;
            col[k] = rk;
            for (int i = 0; i < n; i++)
                if (i != k) col[i] = fma(-Aug[i + k * ld], rk, col[i]);  // column k itself is read-only here
        }
        __syncwarp();
    }
    return pmax > 0.0 ? pmin / pmax : 0.0;
}

}  // namespace dwbc
