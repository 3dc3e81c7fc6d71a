// Team-cooperative dense FP64 linear algebra: one CTA (the "team", 64-256 threads) owns one robot state
// and works on small column-major matrices (n <= ~50) held in shared memory.
//
// Why a CTA and not a warp per state: a state's working set is 30-45 KB of shared memory, so only 5-7
// states fit on an SM.  With one warp each the SM would run 5-7 warps (measured: 4-6 % warp occupancy, FP64
// pipe 2-7 % busy); with 4 warps each it runs 20-28 warps and the matrix products are spread over 128 threads.
//
// Conventions
//   * every routine is called by ALL threads of the CTA with identical arguments;
//   * routines do NOT synchronise on entry or exit unless their comment says so: the caller places
//     tsync() between a producer and its consumers;
//   * leading dimensions are odd (conflict-free column sweeps for 64-bit words).
#pragma once
#include "cuda_compat.h"
#include "warp_linalg.cuh"

namespace dwbc {

// A CTA may hold several teams (lean_cycle_kernel: LEAN_GANG states per CTA, kernels_lean.cuh): thread index within the team and team
// size.  For the one-team kernels (blockDim.x <= TEAM) these are threadIdx.x and blockDim.x.
#ifdef DWBC_CPU_EMUL
__device__ __forceinline__ int t_id() { return (int)threadIdx.x; }
__device__ __forceinline__ int t_n() { return (int)blockDim.x; }
#else
static_assert((TEAM & (TEAM - 1)) == 0, "TEAM must be a power of two");
__device__ __forceinline__ int t_id() { return (int)(threadIdx.x & (TEAM - 1)); }
__device__ __forceinline__ int t_n() { return (int)blockDim.x < TEAM ? (int)blockDim.x : TEAM; }
#endif
__device__ __forceinline__ void tsync() { __syncthreads(); }

// Cooperative copy of a column-major rows x cols block (global or shared on either side).
// (tid, nt): the threads taking part, numbered 0..nt-1 -- the whole team by default, or a sub-team of whole warps.
__device__ __forceinline__ void tcopy(int rows, int cols, const double *__restrict__ src, int lds, double *__restrict__ dst, int ldd, int tid = t_id(),
                                      int nt = t_n()) {
    if (lds == rows && ldd == rows) {
        for (int e = tid; e < rows * cols; e += nt) dst[e] = src[e];
    } else {
        for (int e = tid; e < rows * cols; e += nt) {
            int j = e / rows, i = e - j * rows;
            dst[i + j * ldd] = src[i + j * lds];
        }
    }
}
// Copy into a zero-padded destination block: dst is drows x dcols (leading dimension drows), src is rows x cols (lds).
__device__ __forceinline__ void tcopy_pad(int rows, int cols, const double *__restrict__ src, int lds, int drows, int dcols, double *__restrict__ dst) {
    for (int e = t_id(); e < drows * dcols; e += t_n()) {
        int j = e / drows, i = e - j * drows;
        dst[e] = (i < rows && j < cols) ? src[i + j * lds] : 0.0;
    }
}
__device__ __forceinline__ void tfill(int count, double *dst, double v) {
    for (int e = t_id(); e < count; e += t_n()) dst[e] = v;
}

// C(i,j) <- epi(i, j, sum_k opA(i,k) * opB(k,j)) for an M x N result; same operand conventions as wgemm.
// Tiles are cyclic (rows ti + r*tr, columns tj + c*tc); the tile index runs over the whole team.
template <int TR, int TC, bool TA, bool TB, class Epi>
__device__ __forceinline__ void tgemm(int M, int N, int K, const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, Epi epi) {
    const int tr = (M + TR - 1) / TR, tc = (N + TC - 1) / TC;
    for (int tile = t_id(); tile < tr * tc; tile += t_n()) {
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

// ---- FP64 tensor-core products (DMMA) ---------------------------------------------------------------------
// One mma.sync.m8n8k4.f64 does 256 FMAs for ONE issue slot and two 8-byte operand loads per lane; the register-tiled
// tgemm above needs 8 issue slots and 8-16 shared-memory wavefronts for the same work, and ncu shows the dense stages
// limited by exactly those (LSU wavefronts 45-87 %, DFMA 10 % of the issued instructions; profiles/r01_ncu_main_lqpbuild.md).
// Measured on this pool's B200 (tools/dmma_probe.cu, profiles/r01_dmma_probe.json): DMMA chain 37.2 TFLOP/s vs DFMA chain
// 34.1; a 40x40x40 product out of shared memory with 128 threads: 21.5 TFLOP/s with DMMA vs 7.9 with 2x2 register tiles.
// Fragment layout (PTX ISA, m8n8k4 .f64): lane = 4 g + t;  a = A(g, t);  b = B(t, g);  d0, d1 = D(g, 2t), D(g, 2t + 1).
#ifdef DWBC_CPU_EMUL
__device__ inline void dmma884(double &d0, double &d1, double a, double b) {
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    double s0 = d0, s1 = d1;
    for (int k = 0; k < 4; k++) {
        const double ak = __shfl_sync(DWBC_FULL, a, g * 4 + k);
        const double b0 = __shfl_sync(DWBC_FULL, b, (2 * t) * 4 + k);
        const double b1 = __shfl_sync(DWBC_FULL, b, (2 * t + 1) * 4 + k);
        s0 = fma(ak, b0, s0);
        s1 = fma(ak, b1, s1);
    }
    d0 = s0;
    d1 = s1;
}
#else
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

// Same contract as tgemm (C(i,j) <- epi(i, j, sum_k opA(i,k) opB(k,j)), any M, N, K; operands in shared or global memory),
// computed on the FP64 tensor cores.  A work item is one 8-row band times NT adjacent 8-column tiles (the A fragment is
// reused NT times); items are dealt round-robin to the warps of the team.  Ragged edges are zero-filled by predicated
// loads.  Requires blockDim.x to be a multiple of 32; every warp runs the same item count modulo one.
// KTRI: opA(i,k) = 0 for k < i and opB(k,j) = 0 for k < j (the X^T X product of a lower-triangular X): the k loop of an
// item starts at its first row / column.
// woff rotates the item -> warp assignment, so that several small products issued in one phase do not all land on warp 0.
template <int NT, bool TA, bool TB, bool KTRI = false, class Epi>
__device__ __forceinline__ void tgemm_mma(int M, int N, int K, const double *A, int lda, const double *B, int ldb, Epi epi, int woff = 0) {
    // (A and B are deliberately not __restrict__: an operand may have been written to global memory earlier in the same kernel
    // -- Null_{l-1} in task_space_kernel -- and must not be fetched through the non-coherent read-only path.)
    const int lane = lane_id(), nw = t_n() >> 5, warp = ((t_id() >> 5) + nw - (woff % nw)) % nw;
    const int g = lane >> 2, t = lane & 3;
    const int tr = (M + 7) >> 3, tc = (N + 7) >> 3, gc = (tc + NT - 1) / NT;
    const int sa = TA ? 1 : lda, sb = TB ? ldb : 1;
    const int Kfull = K & ~3;
    for (int item = warp; item < tr * gc; item += nw) {
        const int gj = item / tr, bi = item - gj * tr;
        const int i = bi * 8 + g;
        const bool iv = i < M;
        // Rows / columns beyond the edge read row / column 0 instead of being zero-filled: row i of the product only depends on
        // row i of opA and column j only on column j of opB, so the garbage stays in outputs that are never stored.  Only the
        // k tail (K not a multiple of 4) has to be zeroed.
        const int kbeg = KTRI ? 8 * (bi > gj * NT ? bi : gj * NT) : 0;
        const double *pa = A + (iv ? i : 0) * (TA ? lda : 1) + (kbeg + t) * sa;
        const double *pb[NT];
        double c[NT][2];
        DWBC_UNROLL(NT)
        for (int jj = 0; jj < NT; jj++) {
            const int j = (gj * NT + jj) * 8 + g;
            pb[jj] = B + (j < N ? j : 0) * (TB ? 1 : ldb) + (kbeg + t) * sb;
            c[jj][0] = 0.0;
            c[jj][1] = 0.0;
        }
        const int ntile = tc - gj * NT < NT ? tc - gj * NT : NT;  // tiles of this item that exist (warp-uniform)
        DWBC_UNROLL(2)
        for (int k0 = kbeg; k0 < Kfull; k0 += 4) {
            const double a = *pa;
            pa += 4 * sa;
            DWBC_UNROLL(NT)
            for (int jj = 0; jj < NT; jj++) {
                if (jj < ntile) {
                    const double b = *pb[jj];
                    dmma884(c[jj][0], c[jj][1], a, b);
                }
                pb[jj] += 4 * sb;
            }
        }
        if (Kfull < K && Kfull >= kbeg) {
            const bool kv = Kfull + t < K;
            const double a = kv ? *pa : 0.0;
            DWBC_UNROLL(NT)
            for (int jj = 0; jj < NT; jj++) {
                if (jj < ntile) {
                    const double b = kv ? *pb[jj] : 0.0;
                    dmma884(c[jj][0], c[jj][1], a, b);
                }
            }
        }
        DWBC_UNROLL(NT)
        for (int jj = 0; jj < NT; jj++) {
            const int j = (gj * NT + jj) * 8 + 2 * t;
            if (iv && j < N) epi(i, j, c[jj][0]);
            if (iv && j + 1 < N) epi(i, j + 1, c[jj][1]);
        }
    }
}

// y(i) <- epi(i, sum_k opA(i,k) x(k)).  Four adjacent lanes share one row (each takes every 4th k) and combine with
// two shuffles, so a 33..39-row product keeps the whole team busy.  Requires blockDim.x to be a multiple of 32.
// (tid, nt) as in tcopy: a sub-team of whole warps gives the same sums (the split of k over the four lanes does not change).
template <bool TA, class Epi>
__device__ __forceinline__ void tgemv(int M, int K, const double *__restrict__ A, int lda, const double *__restrict__ x, Epi epi, int tid = t_id(),
                                      int nt = t_n()) {
    const int rows_per_pass = nt >> 2, part = tid & 3;
    for (int r0 = 0; r0 < M; r0 += rows_per_pass) {
        const int i = r0 + (tid >> 2);
        const bool act = i < M;
        const int ii = act ? i : 0;
        double a0 = 0.0;
        DWBC_UNROLL(2)
        for (int k = part; k < K; k += 4) a0 = fma(TA ? A[k + ii * lda] : A[ii + k * lda], x[k], a0);
        a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
        a0 += __shfl_xor_sync(DWBC_FULL, a0, 2);
        if (act && part == 0) epi(i, a0);
    }
}

// ---- Cholesky with register-resident ownership ------------------------------------------------------------
// The lower triangle (n(n+1)/2 entries, packed column by column) is dealt cyclically to the threads; every thread keeps
// its entries in registers for the whole factorisation.  Step k: the owners of column k publish their (fully updated,
// still unscaled) values to shared memory, ONE barrier, then every thread applies  a_ij -= a_ik a_jk / a_kk  to its
// entries with j > k.  A final pass scales column k by 1/sqrt(a_kk).  One barrier per column instead of two, no
// shared-memory read-modify-write, and the trailing update is spread over the whole team.
constexpr int CHOL_MAX_OWN = 8;  // ceil(40*41/2 / 128) = 7 for n <= 40 with >= 128 threads (13 for 64 threads -> see static checks at call sites)

// In-place lower Cholesky of the SPD matrix in the lower triangle of A (n x n, leading dimension ld).  On return the
// lower triangle holds L, rdiag[i] = 1/L(i,i).  Returns the smallest pivot (<= 0: not SPD; the factorisation continues
// with a clamped pivot so no NaN appears).  Synchronises on entry and on exit.  n(n+1)/2 <= MAXOWN * blockDim.x.
template <int MAXOWN>
__device__ __noinline__ double tchol_lower(int n, double *A, int ld, double *rdiag) {
    const int tid = t_id(), nt = t_n();
    // Ownership without a square root: fold the triangle into a ((n+1)/2) x (n+1) rectangle -- fold row a holds column a
    // (n - a entries) followed by column n-1-a (a + 1 entries).
    const int w = n + 1, slots = ((n + 1) >> 1) * w;
    int ei[MAXOWN], ej[MAXOWN];
    double v[MAXOWN];
    tsync();
    DWBC_UNROLL(MAXOWN)
    for (int s = 0; s < MAXOWN; s++) {
        const int e = tid + s * nt;
        int i = 0, j = n;  // j = n marks "no entry"
        if (e < slots) {
            const int a = e / w, b = e - a * w;
            if (b < n - a) {
                j = a;
                i = a + b;
            } else if (n - 1 - a != a) {
                j = n - 1 - a;
                i = j + (b - (n - a));
            }
        }
        ei[s] = i;
        ej[s] = j;
        v[s] = (j < n) ? A[i + j * ld] : 0.0;
    }
    double minpiv = 1e300;
    for (int k = 0; k < n; k++) {
        DWBC_UNROLL(MAXOWN)
        for (int s = 0; s < MAXOWN; s++)
            if (ej[s] == k) A[ei[s] + k * ld] = v[s];
        tsync();
        double akk = A[k + k * ld];
        minpiv = akk < minpiv ? akk : minpiv;
        if (!(akk > 1e-300)) akk = 1e-300;
        const double inv = 1.0 / akk;
        DWBC_UNROLL(MAXOWN)
        for (int s = 0; s < MAXOWN; s++)
            if (ej[s] > k && ej[s] < n) v[s] = fma(-A[ei[s] + k * ld] * inv, A[ej[s] + k * ld], v[s]);
    }
    tsync();
    // rdiag[k] = 1/sqrt(a_kk); the diagonal itself becomes sqrt(a_kk) = a_kk * rdiag[k]
    for (int k = tid; k < n; k += nt) {
        double akk = A[k + k * ld];
        if (!(akk > 1e-300)) akk = 1e-300;
        const double r = rsqrt(akk);
        rdiag[k] = r;
        A[k + k * ld] = akk * r;
    }
    tsync();
    DWBC_UNROLL(MAXOWN)
    for (int s = 0; s < MAXOWN; s++)
        if (ej[s] < n && ei[s] != ej[s]) A[ei[s] + ej[s] * ld] *= rdiag[ej[s]];
    tsync();
    return minpiv;
}

// Xi <- L^{-1} (lower triangular, full n x n storage with zeros above the diagonal) by forward substitution on unit vectors.
// Four adjacent lanes share a column: lane `part` accumulates the terms k = j + part (mod 4) of every row's dot product and
// two shuffles combine them.  A group owns column g and then column n-1-g, so every group runs exactly n + 1 row steps
// (uniform trip count: the shuffles are full-warp) and the long columns are not left to a few threads -- with one column per
// thread 39 of 128 threads worked and a third of ainv_kernel's warp time was spent waiting for them (profiles/r01_ncu_dmma.md).
// Entry (i, j) is written by the lane with part == (i - j) mod 4, which is also the only lane that reads it back.
// L and Xi must not overlap.  No synchronisation (Xi is complete after the caller's next barrier).
__device__ __noinline__ void ttri_inverse(int n, const double *L, int ld, const double *rdiag, double *Xi, int ldx) {
    const int tid = t_id(), nt = t_n(), part = tid & 3;
    for (int e = tid; e < n * n; e += nt) {
        const int j = e / n, i = e - j * n;
        if (i < j) Xi[i + j * ldx] = 0.0;
    }
    const int ng = (n + 1) >> 1;
    for (int g0 = 0; g0 < ng; g0 += nt >> 2) {
        const int g = g0 + (tid >> 2);
        const bool act = g < ng;
        DWBC_UNROLL(1)
        for (int it = 0; it < n + 1; it++) {
            int j, i;
            bool valid = act;
            if (it < n - g) {
                j = g;
                i = g + it;
            } else {
                j = n - 1 - g;
                i = j + (it - (n - g));
                valid = valid && j != g;
            }
            double s0 = 0.0;
            if (valid) {
                const double *lr = L + i, *xc = Xi + j * ldx;
                DWBC_UNROLL(2)
                for (int k = j + part; k < i; k += 4) s0 = fma(lr[k * ld], xc[k], s0);
            }
            s0 += __shfl_xor_sync(DWBC_FULL, s0, 1);
            s0 += __shfl_xor_sync(DWBC_FULL, s0, 2);
            if (valid && ((i - j) & 3) == part) Xi[i + j * ldx] = (i == j) ? rdiag[j] : -s0 * rdiag[i];
        }
    }
}

// SPD inverse by ONE warp, n <= 32, in place (full symmetric result in S); X (n x n, ldx) is scratch.  Only __syncwarp inside:
// different warps of a team can invert different matrices at the same time (task_space_kernel: one hierarchy level per warp)
// instead of the whole team going through n + 5 CTA barriers per matrix.  Returns min pivot / max diagonal (<= 0: not SPD).
__device__ __noinline__ double warp_spd_inverse(int n, double *S, int ld, double *X, int ldx) {
    const int lane = lane_id();
    double dmax = 0.0, minpiv = 1e300;
    for (int i = 0; i < n; i++) dmax = fmax(dmax, S[i + i * ld]);
    double lik = 0.0;
    for (int k = 0; k < n; k++) {
        __syncwarp();
        double akk = S[k + k * ld];
        minpiv = akk < minpiv ? akk : minpiv;
        if (!(akk > 1e-300)) akk = 1e-300;
        const double r = rsqrt(akk);
        __syncwarp();
        if (lane > k && lane < n) {
            lik = S[lane + k * ld] * r;
            S[lane + k * ld] = lik;
        }
        if (lane == k) S[k + k * ld] = akk * r;
        __syncwarp();
        if (lane > k && lane < n)
            for (int j = k + 1; j <= lane; j++) S[lane + j * ld] = fma(-lik, S[j + k * ld], S[lane + j * ld]);
    }
    __syncwarp();
    if (lane < n) {  // X[:, lane] = L^{-1} e_lane
        double *xj = X + lane * ldx;
        for (int i = 0; i < lane; i++) xj[i] = 0.0;
        xj[lane] = 1.0 / S[lane + lane * ld];
        for (int i = lane + 1; i < n; i++) {
            double a0 = 0.0;
            for (int k = lane; k < i; k++) a0 = fma(-S[i + k * ld], xj[k], a0);
            xj[i] = a0 / S[i + i * ld];
        }
    }
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) {  // S = X^T X, lower triangle mirrored
        const int j = e / n, i = e - j * n;
        if (i < j) continue;
        double a0 = 0.0;
        for (int k = i; k < n; k++) a0 = fma(X[k + i * ldx], X[k + j * ldx], a0);
        S[i + j * ld] = a0;
        S[j + i * ld] = a0;
    }
    __syncwarp();
    return dmax > 0.0 ? minpiv / dmax : 0.0;
}

// Out(i,j) <- epi(i, j, (A^{-1})(i,j)) with A^{-1} = Xi^T Xi for the lower-triangular Xi = L^{-1} (full storage, zeros above
// the diagonal): only k >= max(i,j) contributes.  Computed for the full square on the tensor cores (the register-tiled
// triangle version kept 55 of 128 threads busy for n = 39 and was a third of ainv_kernel's stall samples).  No synchronisation.
template <class Epi>
__device__ __forceinline__ void tsyrk_inverse(int n, const double *Xi, int ldx, Epi epi) {
    tgemm_mma<2, true, false, true>(n, n, n, Xi, ldx, Xi, ldx, epi);
}

}  // namespace dwbc
