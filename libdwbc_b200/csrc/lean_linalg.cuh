// Compile-time-specialised dense FP64 linear algebra for ONE robot state per CTA ("team" of TEAM = 128 threads, four warps).
//
// Why this layer exists (profiles/r01_ncu_final_d4eaa6f.md): the runtime-dimension routines of team_linalg.cuh spent ~90 % of
// their issued instructions on index arithmetic, edge predicates and loop control, and the SPD inverses went through one CTA
// barrier per matrix column.  Here every dimension and leading dimension is a template parameter, the k loops are fully
// unrolled (operand addresses become immediates), and the factorisations are BLOCKED by 8 so that all O(n^3) work is
// mma.sync.m8n8k4.f64 (SASS DMMA) and only an 8-wide panel is factored by one warp with shuffles:
//
//   chol_blocked   A = L L^T      panel (one warp, rows in registers) + DMMA trailing update; 2 barriers per 8 columns
//   trinv_blocked  X = L^{-1}     8x8 diagonal inverses (one warp each), then one warp per block column with DMMA
//   syrk_inverse   A^{-1} = X^T X lower tiles on DMMA, mirrored
//
// Serial sections rotate over the warps of the team (warp w of every resident CTA lives on the same SM sub-partition; without
// the rotation all panels of all CTAs would queue on one scheduler).
//
// Shared-memory layout rule: leading dimensions = 4 (mod 8) doubles make the 8-byte DMMA fragment loads (lane = 4 g + t reads
// element (g, t)) hit 16 distinct bank pairs per half-warp for both the plain and the transposed operand.
// Ragged edges: row / column indices beyond the matrix are clamped to the last valid one (the garbage lands in outputs that are
// never stored); only the k tail is zero-filled, by predicated loads in the last k step.
#pragma once
#include <type_traits>

#include "cuda_compat.h"
#include "team_linalg.cuh"

namespace dwbc {
namespace lean {

constexpr int NW = TEAM / 32;

__device__ __forceinline__ int warp_id() { return (int)((threadIdx.x >> 5) & (NW - 1)); }  // warp within the team

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// C(i,j) <- epi(i, j, sum_k opA(i,k) opB(k,j)),  i < M, j < N, k < K, all compile-time.
//   opA(i,k) = TA ? A[k + i*LDA] : A[i + k*LDA],   opB(k,j) = TB ? B[j + k*LDB] : B[k + j*LDB]
// A work item is one 8-row band times NT adjacent 8-column tiles; items are dealt round-robin to the warps, starting with warp
// `wrot` (rotate it between the small products of one phase).  No synchronisation.
template <int M, int N, int K, int LDA, int LDB, bool TA, bool TB, int NT, class Epi>
__device__ __forceinline__ void mma_gemm(const double *A, const double *B, Epi epi, int wrot = 0) {
    constexpr int TR = (M + 7) / 8, TC = (N + 7) / 8, GC = (TC + NT - 1) / NT, KS = (K + 3) / 4, KR = K & 3;
    constexpr int SA = TA ? 1 : LDA, SB = TB ? LDB : 1;  // stride of k
    constexpr int RA = TA ? LDA : 1, RB = TB ? 1 : LDB;  // stride of i (A) / j (B)
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    const int w = (warp_id() + NW - (wrot % NW)) % NW;
    for (int item = w; item < TR * GC; item += NW) {
        const int gj = item / TR, bi = item - gj * TR;
        const int i = bi * 8 + g;
        const int ic = (M % 8 == 0 || i < M) ? i : M - 1;
        const double *pa = A + ic * RA + t * SA;
        const double *pb[NT];
        double c[NT][2];
        DWBC_UNROLL(NT)
        for (int jj = 0; jj < NT; jj++) {
            const int j = (gj * NT + jj) * 8 + g;
            const int jc = j < N ? j : N - 1;
            pb[jj] = B + jc * RB + t * SB;
            c[jj][0] = 0.0;
            c[jj][1] = 0.0;
        }
#ifdef DWBC_LEAN_ROLLK
        // experiment: rolled k loop (smaller code, run-time addresses) -- see tools/gpu_variants.sh
        DWBC_UNROLL(1)
        for (int ks = 0; ks < KS - (KR != 0 ? 1 : 0); ks++) {
            const double a = pa[ks * 4 * SA];
            DWBC_UNROLL(NT)
            for (int jj = 0; jj < NT; jj++) {
                if ((TC % NT == 0) || gj * NT + jj < TC) {
                    const double b = pb[jj][ks * 4 * SB];
                    dmma884(c[jj][0], c[jj][1], a, b);
                }
            }
        }
        if constexpr (KR != 0) {
            constexpr int ks = KS - 1;
            const double a = t < KR ? pa[ks * 4 * SA] : 0.0;
            DWBC_UNROLL(NT)
            for (int jj = 0; jj < NT; jj++) {
                if ((TC % NT == 0) || gj * NT + jj < TC) {
                    const double b = t < KR ? pb[jj][ks * 4 * SB] : 0.0;
                    dmma884(c[jj][0], c[jj][1], a, b);
                }
            }
        }
#else
        static_for<0, KS>([&](auto ks_) {
            constexpr int ks = decltype(ks_)::value;
            constexpr bool tail = (ks == KS - 1) && KR != 0;
            double a;
            if constexpr (tail) a = t < KR ? pa[ks * 4 * SA] : 0.0;
            else a = pa[ks * 4 * SA];
            DWBC_UNROLL(NT)
            for (int jj = 0; jj < NT; jj++) {
                if ((TC % NT == 0) || gj * NT + jj < TC) {
                    double b;
                    if constexpr (tail) b = t < KR ? pb[jj][ks * 4 * SB] : 0.0;
                    else b = pb[jj][ks * 4 * SB];
                    dmma884(c[jj][0], c[jj][1], a, b);
                }
            }
        });
#endif
        DWBC_UNROLL(NT)
        for (int jj = 0; jj < NT; jj++) {
            const int j = (gj * NT + jj) * 8 + 2 * t;
            if ((M % 8 == 0 || i < M) && j < N) epi(i, j, c[jj][0]);
            if ((M % 8 == 0 || i < M) && j + 1 < N) epi(i, j + 1, c[jj][1]);
        }
    }
}

// Symmetric result: C = opA opB with C known to be symmetric (e.g. X X^T, V V^T): only the lower tiles are computed, the
// epilogue is called for (i, j) with i's tile >= j's tile and must store both (i, j) and (j, i).  M = N.
template <int N, int K, int LDA, int LDB, bool TA, bool TB, class Epi>
__device__ __forceinline__ void mma_gemm_sym(const double *A, const double *B, Epi epi, int wrot = 0) {
    constexpr int T = (N + 7) / 8, NTILES = T * (T + 1) / 2, KS = (K + 3) / 4, KR = K & 3;
    constexpr int SA = TA ? 1 : LDA, SB = TB ? LDB : 1;
    constexpr int RA = TA ? LDA : 1, RB = TB ? 1 : LDB;
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    const int w = (warp_id() + NW - (wrot % NW)) % NW;
    for (int p = w; p < NTILES; p += NW) {
        int bj = 0, rem = p;
        while (rem >= T - bj) {
            rem -= T - bj;
            bj++;
        }
        const int bi = bj + rem;
        const int i = bi * 8 + g, jg = bj * 8 + g;
        const int ic = i < N ? i : N - 1, jc = jg < N ? jg : N - 1;
        const double *pa = A + ic * RA + t * SA, *pb = B + jc * RB + t * SB;
        double c0 = 0.0, c1 = 0.0;
        static_for<0, KS>([&](auto ks_) {
            constexpr int ks = decltype(ks_)::value;
            constexpr bool tail = (ks == KS - 1) && KR != 0;
            double a, b;
            if constexpr (tail) {
                a = t < KR ? pa[ks * 4 * SA] : 0.0;
                b = t < KR ? pb[ks * 4 * SB] : 0.0;
            } else {
                a = pa[ks * 4 * SA];
                b = pb[ks * 4 * SB];
            }
            dmma884(c0, c1, a, b);
        });
        const int j = bj * 8 + 2 * t;
        if (i < N && j < N && (bi > bj || i >= j)) epi(i, j, c0);
        if (i < N && j + 1 < N && (bi > bj || i >= j + 1)) epi(i, j + 1, c1);
    }
}

// y(i) <- epi(i, sum_k opA(i,k) x(k)), LPR lanes per row (1, 2 or 4) combined by shuffles.  Called by whole warps
// (tid, nt as in tcopy).
template <int M, int K, int LDA, bool TA, int LPR, class Epi>
__device__ __forceinline__ void gemv(const double *A, const double *x, Epi epi, int tid = t_id(), int nt = TEAM) {
    static_assert(LPR == 1 || LPR == 2 || LPR == 4, "lanes per row");
    const int rpp = nt / LPR, part = tid % LPR;
    for (int r0 = 0; r0 < M; r0 += rpp) {
        const int i = r0 + tid / LPR;
        const bool act = i < M;
        const int ii = act ? i : M - 1;
        const double *pa = TA ? A + ii * LDA : A + ii;
        double a0 = 0.0, a1 = 0.0;
        // rolled (unroll 4): the matrix-vector products are latency-bound anyway and a fully unrolled copy per call site made up a fifth
        // of the one-kernel cycle's code (8 k of 40 k SASS instructions)
        constexpr int FULL = K / (2 * LPR);  // pairs of steps every lane takes
        DWBC_UNROLL(4)
        for (int s2 = 0; s2 < FULL; s2++) {
            const int k0 = (2 * s2) * LPR + part, k1 = k0 + LPR;
            a0 = fma(TA ? pa[k0] : pa[k0 * LDA], x[k0], a0);
            a1 = fma(TA ? pa[k1] : pa[k1 * LDA], x[k1], a1);
        }
        DWBC_UNROLL(2 * LPR)
        for (int k = FULL * 2 * LPR + part; k < K; k += LPR) a0 = fma(TA ? pa[k] : pa[k * LDA], x[k], a0);
        a0 += a1;
        if constexpr (LPR >= 2) a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
        if constexpr (LPR >= 4) a0 += __shfl_xor_sync(DWBC_FULL, a0, 2);
        if (act && part == 0) epi(i, a0);
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Blocked Cholesky.  Panel KB (columns 8 KB .. 8 KB + W - 1, rows from 8 KB down): executed by ONE warp, lane = row (two rows
// per lane when more than 32 rows remain), the W columns of a row in registers.  Step k: broadcast the pivot, scale column k by
// 1/sqrt, rank-one update of columns k+1.. with the multipliers L(j, k) broadcast from lane j.
template <int N, int LD, int KB>
__device__ __forceinline__ void chol_panel(double *A, double *rdiag, double *pivslot) {
    constexpr int C0 = 8 * KB, W = (N - C0 < 8) ? N - C0 : 8, ROWS = N - C0, SETS = (ROWS + 31) / 32;
    const int lane = lane_id();
    double a[SETS][W];
    DWBC_UNROLL(SETS)
    for (int s = 0; s < SETS; s++) {
        const int r = C0 + 32 * s + lane;
        DWBC_UNROLL(W)
        for (int k = 0; k < W; k++) a[s][k] = r < N ? A[r + (C0 + k) * LD] : 0.0;
    }
    double minp = 1e300;
    static_for<0, W>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        double d = __shfl_sync(DWBC_FULL, a[0][k], k);
        minp = d < minp ? d : minp;
        if (!(d > 1e-300)) d = 1e-300;
        const double rs = rsqrt(d);
        DWBC_UNROLL(SETS)
        for (int s = 0; s < SETS; s++) a[s][k] *= rs;
        if (lane == k) a[0][k] = d * rs;
        if (lane == 0) rdiag[C0 + k] = rs;
        static_for<k + 1, W>([&](auto j_) {
            constexpr int j = decltype(j_)::value;
            const double ljk = __shfl_sync(DWBC_FULL, a[0][k], j);
            DWBC_UNROLL(SETS)
            for (int s = 0; s < SETS; s++) a[s][j] = fma(-a[s][k], ljk, a[s][j]);
        });
    });
    DWBC_UNROLL(SETS)
    for (int s = 0; s < SETS; s++) {
        const int r = C0 + 32 * s + lane;
        DWBC_UNROLL(W)
        for (int k = 0; k < W; k++)
            if (r < N) A[r + (C0 + k) * LD] = (s == 0 && lane < k) ? 0.0 : a[s][k];  // exact zeros above the diagonal of the block
    }
    if (lane == 0) *pivslot = minp < *pivslot ? minp : *pivslot;
}

// Trailing update after panel KB: A22 -= L21 L21^T on the lower tiles (diagonal tiles in full), two DMMA per tile.
template <int N, int LD, int KB>
__device__ __forceinline__ void chol_trailing(double *A) {
    constexpr int NB = (N + 7) / 8, NR = NB - 1 - KB, NTILES = NR * (NR + 1) / 2, C0 = 8 * KB;
    const int lane = lane_id(), g = lane >> 2, t = lane & 3, warp = warp_id();
    for (int p = warp; p < NTILES; p += NW) {
        int bj = 0, rem = p;
        while (rem >= NR - bj) {
            rem -= NR - bj;
            bj++;
        }
        const int bi = bj + rem + KB + 1;
        bj += KB + 1;
        const int i = 8 * bi + g, jg = 8 * bj + g;
        const int ic = i < N ? i : N - 1, jc = jg < N ? jg : N - 1;
        const double *pa = A + ic + (C0 + t) * LD, *pb = A + jc + (C0 + t) * LD;
        const int cj = 8 * bj + 2 * t;
        const bool v0 = i < N && cj < N, v1 = i < N && cj + 1 < N;
        double c0 = v0 ? A[i + cj * LD] : 0.0, c1 = v1 ? A[i + (cj + 1) * LD] : 0.0;
        const double a0 = -pa[0], a1 = -pa[4 * LD], b0 = pb[0], b1 = pb[4 * LD];
        dmma884(c0, c1, a0, b0);
        dmma884(c0, c1, a1, b1);
        if (v0) A[i + cj * LD] = c0;
        if (v1) A[i + (cj + 1) * LD] = c1;
    }
}

// In-place lower Cholesky of the SPD matrix in A (N x N, LD; the full symmetric matrix must be finite).  On return the lower
// triangle holds L with exact zeros above the diagonal inside the 8x8 diagonal blocks, rdiag[i] = 1/L(i,i), *pivslot = smallest
// pivot (<= 0: not SPD; the factorisation continues with a clamped pivot).  Synchronises on entry and exit.
// `side`: work of the OTHER warps while the first panel is factored by its one warp (it must not touch A, rdiag or *pivslot, is called
// by every warp with the index of the panel warp, and is complete at the barrier that follows the panel).
struct NoSideWork {
    __device__ __forceinline__ void operator()(int) const {}
};
template <int N, int LD, class Side = NoSideWork>
__device__ __forceinline__ void chol_blocked(double *A, double *rdiag, double *pivslot, int wrot = 0, Side side = Side()) {
    constexpr int NB = (N + 7) / 8;
    tsync();
    if (t_id() == 0) *pivslot = 1e300;
    static_for<0, NB>([&](auto kb_) {
        constexpr int KB = decltype(kb_)::value;
        if constexpr (KB > 0) tsync();
        if (warp_id() == (KB + wrot) % NW) chol_panel<N, LD, KB>(A, rdiag, pivslot);
        if constexpr (KB == 0) side(wrot % NW);
        if constexpr (KB < NB - 1) {
            tsync();
            chol_trailing<N, LD, KB>(A);
        }
    });
    tsync();
}

// X = L^{-1} (lower triangular; the 8x8 diagonal blocks are written in full, the blocks above the diagonal are NOT touched).
// L, rdiag as left by chol_blocked.  Synchronises once inside, not on entry / exit.
template <int N, int LD, int LDX>
__device__ __forceinline__ void trinv_blocked(const double *L, const double *rdiag, double *X, int wrot = 0) {
    constexpr int NB = (N + 7) / 8;
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    const int warp = (warp_id() + NW - (wrot % NW)) % NW;
    // diagonal blocks: lane j owns column j of the block; x_i = (i == j) ? 1/l_ii : -(sum_{k<i} l_ik x_k) / l_ii  (x_k = 0 for k < j)
    for (int kb = warp; kb < NB; kb += NW) {
        const int c0 = 8 * kb, wd = N - c0 < 8 ? N - c0 : 8;
        double x[8];
        DWBC_UNROLL(8)
        for (int i = 0; i < 8; i++) {
            double s = 0.0;
            x[i] = 0.0;
            if (i < wd) {  // warp-uniform
                DWBC_UNROLL(8)
                for (int k = 0; k < i; k++) s = fma(L[c0 + i + (c0 + k) * LD], x[k], s);
                const double rd = rdiag[c0 + i];
                x[i] = (i == lane) ? rd : -s * rd;
            }
        }
        if (lane < wd) {
            DWBC_UNROLL(8)
            for (int i = 0; i < 8; i++)
                if (i < wd) X[c0 + i + (c0 + lane) * LDX] = x[i];
        }
    }
    tsync();
    // block column jb (one per warp): X_ij = -X_ii (sum_{k=j}^{i-1} L_ik X_kj), i = j+1 .. NB-1
    for (int jb = warp; jb < NB - 1; jb += NW) {
        for (int ib = jb + 1; ib < NB; ib++) {
            const int i = 8 * ib + g, ic = i < N ? i : N - 1;
            double c0 = 0.0, c1 = 0.0;
            for (int kb = jb; kb < ib; kb++) {
                const double *pa = L + ic + (8 * kb + t) * LD;
                const double *pb = X + (8 * kb + t) + (8 * jb + g) * LDX;
                dmma884(c0, c1, pa[0], pb[0]);
                dmma884(c0, c1, pa[4 * LD], pb[4]);
            }
            // park S = sum in the destination block, then X_ij = -X_ii S
            double *dst = X + i + (8 * jb + 2 * t) * LDX;
            if (i < N) {
                dst[0] = c0;
                dst[LDX] = c1;
            }
            __syncwarp();
            const bool k0v = 8 * ib + t < N, k1v = 8 * ib + t + 4 < N;
            const double *xa = X + ic + (8 * ib + t) * LDX;
            const double *sb = X + (8 * ib + t) + (8 * jb + g) * LDX;
            const double a0 = k0v ? xa[0] : 0.0, a1 = k1v ? xa[4 * LDX] : 0.0;
            const double b0 = k0v ? sb[0] : 0.0, b1 = k1v ? sb[4] : 0.0;
            double d0 = 0.0, d1 = 0.0;
            dmma884(d0, d1, a0, b0);
            dmma884(d0, d1, a1, b1);
            __syncwarp();
            if (i < N) {
                dst[0] = -d0;
                dst[LDX] = -d1;
            }
            __syncwarp();
        }
    }
}

// S(i,j) <- epi(i, j, (X^T X)(i,j)) for the lower-triangular X = L^{-1} of trinv_blocked: only rows k >= 8 max(bi, bj) of X
// contribute to tile (bi, bj).  Lower tiles only (the diagonal tiles' lower half); the epilogue stores (i, j) and (j, i).
template <int N, int LDX, class Epi>
__device__ __forceinline__ void syrk_inverse(const double *X, Epi epi, int wrot = 0) {
    constexpr int NB = (N + 7) / 8;
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    const int warp = (warp_id() + NW - (wrot % NW)) % NW;
    static_for<0, NB>([&](auto bi_) {
        constexpr int bi = decltype(bi_)::value;
        constexpr int K0 = 8 * bi, KS = (N - K0 + 3) / 4, KR = (N - K0) & 3;
        constexpr int P0 = bi * (bi + 1) / 2;
        for (int bj = (warp + NW - (P0 % NW)) % NW; bj <= bi; bj += NW) {  // tile index P0 + bj dealt round-robin
            const int i = 8 * bi + g, jg = 8 * bj + g;
            const int ic = i < N ? i : N - 1;
            const double *pa = X + (K0 + t) + ic * LDX, *pb = X + (K0 + t) + jg * LDX;  // bj <= bi < NB: column jg < N whenever bj < bi
            const double *pbc = (jg < N) ? pb : X + (K0 + t) + (N - 1) * LDX;
            double c0 = 0.0, c1 = 0.0;
            static_for<0, KS>([&](auto ks_) {
                constexpr int ks = decltype(ks_)::value;
                constexpr bool tail = (ks == KS - 1) && KR != 0;
                double a, b;
                if constexpr (tail) {
                    a = t < KR ? pa[ks * 4] : 0.0;
                    b = t < KR ? pbc[ks * 4] : 0.0;
                } else {
                    a = pa[ks * 4];
                    b = pbc[ks * 4];
                }
                dmma884(c0, c1, a, b);
            });
            const int j = 8 * bj + 2 * t;
            if (i < N && j < N && (bi > bj || i >= j)) epi(i, j, c0);
            if (i < N && j + 1 < N && (bi > bj || i >= j + 1)) epi(i, j + 1, c1);
        }
    });
}

// SPD inverse in place: on return A holds the full symmetric inverse; X (N x N, LDX) is scratch, rdiag N doubles, *pivslot the
// smallest Cholesky pivot.  Synchronises on entry and exit.
template <int N, int LD, int LDX, class Side = NoSideWork>
__device__ __forceinline__ void spd_inverse(double *A, double *X, double *rdiag, double *pivslot, int wrot = 0, Side side = Side()) {
    chol_blocked<N, LD, Side>(A, rdiag, pivslot, wrot, side);
    trinv_blocked<N, LD, LDX>(A, rdiag, X, wrot);
    tsync();
    syrk_inverse<N, LDX>(X, [&](int i, int j, double v) {
        A[i + j * LD] = v;
        A[j + i * LD] = v;
    }, wrot);
    tsync();
}

// SPD inverse of a small matrix (N <= 8) by ONE warp, in place (full symmetric result); Xs: N*N doubles of scratch private to
// the warp.  Only __syncwarp inside.  Returns min pivot / max diagonal (<= 0: not SPD), identical in every lane.
template <int N, int LD>
__device__ __forceinline__ double warp_spd_inverse_small(double *S, double *Xs) {
    static_assert(N <= 8, "one panel");
    const int lane = lane_id();
    double a[N], rs[N];
    DWBC_UNROLL(N)
    for (int k = 0; k < N; k++) a[k] = lane < N ? S[lane + k * LD] : 0.0;
    double minp = 1e300, dmax = 0.0;
    // max of the ORIGINAL diagonal (for the relative pivot): lane k holds S(k,k) in a[k] before the factorisation
    {
        double dk = 0.0;
        DWBC_UNROLL(N)
        for (int k = 0; k < N; k++)
            if (lane == k) dk = a[k];
        dmax = warp_max(dk);
    }
    static_for<0, N>([&](auto k_) {
        constexpr int k = decltype(k_)::value;
        double d = __shfl_sync(DWBC_FULL, a[k], k);
        minp = d < minp ? d : minp;
        if (!(d > 1e-300)) d = 1e-300;
        rs[k] = rsqrt(d);
        a[k] *= rs[k];
        if (lane == k) a[k] = d * rs[k];
        static_for<k + 1, N>([&](auto j_) {
            constexpr int j = decltype(j_)::value;
            const double ljk = __shfl_sync(DWBC_FULL, a[k], j);
            a[j] = fma(-a[k], ljk, a[j]);
        });
    });
    // column `lane` of L^{-1}: x_i = (i == lane) ? rs_i : -(sum_{k<i} L(i,k) x_k) rs_i, L(i,k) broadcast from lane i
    double x[N];
    static_for<0, N>([&](auto i_) {
        constexpr int i = decltype(i_)::value;
        double s = 0.0;
        static_for<0, i>([&](auto k_) {
            constexpr int k = decltype(k_)::value;
            const double lik = __shfl_sync(DWBC_FULL, a[k], i);
            s = fma(lik, x[k], s);
        });
        x[i] = (i == lane) ? rs[i] : -s * rs[i];
        if (i < lane) x[i] = 0.0;
    });
    if (lane < N) {
        DWBC_UNROLL(N)
        for (int i = 0; i < N; i++) Xs[i + lane * N] = x[i];
    }
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) {
        const int j = e / N, i = e - j * N;
        if (i < j) continue;
        double s = 0.0;
        for (int k = i; k < N; k++) s = fma(Xs[k + i * N], Xs[k + j * N], s);
        S[i + j * LD] = s;
        S[j + i * LD] = s;
    }
    __syncwarp();
    return dmax > 0.0 ? minp / dmax : 0.0;
}


// ===========================================================================================================================
// Run-time-dimension twins of the routines above, each compiled ONCE (__noinline__) and called from every site.
//
// Why (profiles/r02_ncu_lean_first.md): with one fully unrolled copy per call site the one-kernel cycle was 29 k SASS instructions
// (468 KB) of straight-line code that every CTA streams once per state; the SM's instruction cache (L1.5: 32 KB) hit only 69 % of the
// fetches (sm__icc_request_hit_rate) and the one warp working on a serial section spent its time waiting for instructions.  The
// blocked algorithms and the DMMA inner loops are the same; only the dimensions, leading dimensions and transposes are arguments, so
// ~2 k instructions of library code serve all ~40 call sites of a state and stay resident across the CTAs of an SM.
//
// MEASURED AND NOT USED by lean_cycle_kernel (profiles/r02_*_rt_primitives_dropped.*): the instruction-cache hit rate rose from 69 % to
// 84 % and the kernel shrank from 29 k to 13 k SASS instructions, but the one-kernel cycle took 21.9 ms instead of 10.1 ms per 65 536
// states -- every phase about twice as long (run-time strides, call overhead and the panel's worst-case register tiles cost more issue
// slots and dependent latency than the instruction fetch they save).  Kept, with their unit tests, as the documented alternative.
namespace rt {

constexpr int NTMAX = 5;  // column tiles per work item (N <= 40)

// C(i,j) <- alpha * sum_k opA(i,k) opB(k,j) + (beta ? C(i,j) : 0).  Element (i,k) of opA is A[i*ra + k*sa], (k,j) of opB is B[k*sb + j*rb]
// (any transpose / leading dimension); C is column-major with ldc.  A work item is one 8-row band times nt (<= NTMAX) adjacent column
// tiles; nt = all column tiles makes the product safe IN PLACE on A (an item only reads the band it later writes).  No synchronisation.
__device__ __noinline__ void gemm(int M, int N, int K, const double *A, int ra, int sa, const double *B, int sb, int rb, double *C, int ldc, double alpha,
                                  int beta, int nt, int wrot) {
    const int lane = lane_id(), g = lane >> 2, t = lane & 3;
    const int w = (warp_id() + NW - (wrot % NW)) % NW;
    const int TR = (M + 7) >> 3, TC = (N + 7) >> 3, GC = (TC + nt - 1) / nt, KF = K >> 2, KR = K & 3;
    DWBC_UNROLL(1)
    for (int item = w; item < TR * GC; item += NW) {
        const int gj = item / TR, bi = item - gj * TR;
        const int i = bi * 8 + g, ic = i < M ? i : M - 1;
        const double *pa = A + ic * ra + t * sa;
        const double *pb[NTMAX];
        double c[NTMAX][2];
        DWBC_UNROLL(NTMAX)
        for (int jj = 0; jj < NTMAX; jj++) {
            const int j = (gj * nt + jj) * 8 + g, jc = j < N ? j : N - 1;
            pb[jj] = B + jc * rb + t * sb;
            c[jj][0] = 0.0;
            c[jj][1] = 0.0;
        }
        const int ntile = TC - gj * nt < nt ? TC - gj * nt : nt;  // tiles of this item (warp-uniform)
        DWBC_UNROLL(2)
        for (int ks = 0; ks < KF; ks++) {
            const double a = *pa;
            pa += 4 * sa;
            DWBC_UNROLL(NTMAX)
            for (int jj = 0; jj < NTMAX; jj++) {
                if (jj < ntile) {
                    const double b = *pb[jj];
                    pb[jj] += 4 * sb;
                    dmma884(c[jj][0], c[jj][1], a, b);
                }
            }
        }
        if (KR) {
            const bool kv = t < KR;
            const double a = kv ? *pa : 0.0;
            DWBC_UNROLL(NTMAX)
            for (int jj = 0; jj < NTMAX; jj++) {
                if (jj < ntile) {
                    const double b = kv ? *pb[jj] : 0.0;
                    dmma884(c[jj][0], c[jj][1], a, b);
                }
            }
        }
        DWBC_UNROLL(NTMAX)
        for (int jj = 0; jj < NTMAX; jj++) {
            if (jj < ntile && i < M) {
                const int j = (gj * nt + jj) * 8 + 2 * t;
                double *pc = C + i + j * ldc;
                if (j < N) pc[0] = beta ? fma(alpha, c[jj][0], pc[0]) : alpha * c[jj][0];
                if (j + 1 < N) pc[ldc] = beta ? fma(alpha, c[jj][1], pc[ldc]) : alpha * c[jj][1];
            }
        }
    }
}

// y(i) <- alpha * sum_k opA(i,k) x(k) + (z ? z(i) : 0); two lanes per row.  (tid, nt): the participating threads (whole warps).
__device__ __noinline__ void gemv(int M, int K, const double *A, int ra, int sa, const double *x, double *y, double alpha, const double *z, int tid, int nt) {
    const int rpp = nt >> 1, part = tid & 1;
    DWBC_UNROLL(1)
    for (int r0 = 0; r0 < M; r0 += rpp) {
        const int i = r0 + (tid >> 1);
        const bool act = i < M;
        const double *pa = A + (act ? i : M - 1) * ra;
        double a0 = 0.0, a1 = 0.0;
        int k = part;
        DWBC_UNROLL(2)
        for (; k + 2 < K; k += 4) {
            a0 = fma(pa[k * sa], x[k], a0);
            a1 = fma(pa[(k + 2) * sa], x[k + 2], a1);
        }
        if (k < K) a0 = fma(pa[k * sa], x[k], a0);
        a0 += a1;
        a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
        if (act && part == 0) y[i] = fma(alpha, a0, z ? z[i] : 0.0);
    }
}

// Panel of the blocked Cholesky: columns c0 .. c0 + w - 1 (w <= 8), rows c0 .. n - 1 (at most 64), executed by one warp (see chol_panel).
__device__ __forceinline__ void chol_panel(double *A, int ld, int n, int c0, int w, double *rdiag, double *pivslot) {
    const int lane = lane_id();
    double a[2][8];
    DWBC_UNROLL(2)
    for (int s = 0; s < 2; s++) {
        const int r = c0 + 32 * s + lane;
        DWBC_UNROLL(8)
        for (int k = 0; k < 8; k++) a[s][k] = (r < n && k < w) ? A[r + (c0 + k) * ld] : 0.0;
    }
    double minp = 1e300;
    DWBC_UNROLL(8)
    for (int k = 0; k < 8; k++) {
        if (k < w) {  // warp-uniform
            double d = __shfl_sync(DWBC_FULL, a[0][k], k);
            minp = d < minp ? d : minp;
            if (!(d > 1e-300)) d = 1e-300;
            const double rs = rsqrt(d);
            a[0][k] *= rs;
            a[1][k] *= rs;
            if (lane == k) a[0][k] = d * rs;
            if (lane == 0) rdiag[c0 + k] = rs;
            DWBC_UNROLL(8)
            for (int j = k + 1; j < 8; j++) {
                const double ljk = __shfl_sync(DWBC_FULL, a[0][k], j);
                a[0][j] = fma(-a[0][k], ljk, a[0][j]);
                a[1][j] = fma(-a[1][k], ljk, a[1][j]);
            }
        }
    }
    DWBC_UNROLL(2)
    for (int s = 0; s < 2; s++) {
        const int r = c0 + 32 * s + lane;
        DWBC_UNROLL(8)
        for (int k = 0; k < 8; k++)
            if (r < n && k < w) A[r + (c0 + k) * ld] = (s == 0 && lane < k) ? 0.0 : a[s][k];
    }
    if (lane == 0) *pivslot = minp < *pivslot ? minp : *pivslot;
}

// SPD inverse in place (see spd_inverse): A n x n (ld) -> full symmetric inverse; X (n x n, ldx) scratch; rdiag n doubles; *pivslot the
// smallest pivot.  Synchronises on entry and exit.
__device__ __noinline__ void spd_inverse(double *A, int ld, double *X, int ldx, int n, double *rdiag, double *pivslot, int wrot) {
    const int NB = (n + 7) >> 3;
    const int lane = lane_id(), g = lane >> 2, t = lane & 3, wid = warp_id();
    const int warp = (wid + NW - (wrot % NW)) % NW;
    tsync();
    if (t_id() == 0) *pivslot = 1e300;
    // ---- blocked Cholesky ----
    DWBC_UNROLL(1)
    for (int kb = 0; kb < NB; kb++) {
        const int c0 = 8 * kb, w = n - c0 < 8 ? n - c0 : 8;
        if (kb > 0) tsync();
        if (wid == (kb + wrot) % NW) chol_panel(A, ld, n, c0, w, rdiag, pivslot);
        if (kb < NB - 1) {
            tsync();
            const int NR = NB - 1 - kb, ntiles = NR * (NR + 1) / 2;
            DWBC_UNROLL(1)
            for (int p = wid; p < ntiles; p += NW) {
                int bj = 0, rem = p;
                while (rem >= NR - bj) {
                    rem -= NR - bj;
                    bj++;
                }
                const int bi = bj + rem + kb + 1;
                bj += kb + 1;
                const int i = 8 * bi + g, jg = 8 * bj + g;
                const int ic = i < n ? i : n - 1, jc = jg < n ? jg : n - 1;
                const double *pa = A + ic + (c0 + t) * ld, *pb = A + jc + (c0 + t) * ld;
                const int cj = 8 * bj + 2 * t;
                const bool v0 = i < n && cj < n, v1 = i < n && cj + 1 < n;
                double c0v = v0 ? A[i + cj * ld] : 0.0, c1v = v1 ? A[i + (cj + 1) * ld] : 0.0;
                const double a0 = -pa[0], a1 = -pa[4 * ld], b0 = pb[0], b1 = pb[4 * ld];
                dmma884(c0v, c1v, a0, b0);
                dmma884(c0v, c1v, a1, b1);
                if (v0) A[i + cj * ld] = c0v;
                if (v1) A[i + (cj + 1) * ld] = c1v;
            }
        }
    }
    tsync();
    // ---- X = L^{-1}: diagonal blocks (one warp each), then one warp per block column ----
    DWBC_UNROLL(1)
    for (int kb = warp; kb < NB; kb += NW) {
        const int c0 = 8 * kb, wd = n - c0 < 8 ? n - c0 : 8;
        double x[8];
        DWBC_UNROLL(8)
        for (int i = 0; i < 8; i++) {
            double s = 0.0;
            x[i] = 0.0;
            if (i < wd) {
                DWBC_UNROLL(8)
                for (int k = 0; k < i; k++) s = fma(A[c0 + i + (c0 + k) * ld], x[k], s);
                const double rd = rdiag[c0 + i];
                x[i] = (i == lane) ? rd : -s * rd;
            }
        }
        if (lane < wd) {
            DWBC_UNROLL(8)
            for (int i = 0; i < 8; i++)
                if (i < wd) X[c0 + i + (c0 + lane) * ldx] = x[i];
        }
    }
    tsync();
    DWBC_UNROLL(1)
    for (int jb = warp; jb < NB - 1; jb += NW) {
        DWBC_UNROLL(1)
        for (int ib = jb + 1; ib < NB; ib++) {
            const int i = 8 * ib + g, ic = i < n ? i : n - 1;
            double c0 = 0.0, c1 = 0.0;
            DWBC_UNROLL(1)
            for (int kb = jb; kb < ib; kb++) {
                const double *pa = A + ic + (8 * kb + t) * ld;
                const double *pb = X + (8 * kb + t) + (8 * jb + g) * ldx;
                dmma884(c0, c1, pa[0], pb[0]);
                dmma884(c0, c1, pa[4 * ld], pb[4]);
            }
            double *dst = X + i + (8 * jb + 2 * t) * ldx;
            if (i < n) {
                dst[0] = c0;
                dst[ldx] = c1;
            }
            __syncwarp();
            const bool k0v = 8 * ib + t < n, k1v = 8 * ib + t + 4 < n;
            const double *xa = X + ic + (8 * ib + t) * ldx;
            const double *sb = X + (8 * ib + t) + (8 * jb + g) * ldx;
            const double a0 = k0v ? xa[0] : 0.0, a1 = k1v ? xa[4 * ldx] : 0.0;
            const double b0 = k0v ? sb[0] : 0.0, b1 = k1v ? sb[4] : 0.0;
            double d0 = 0.0, d1 = 0.0;
            dmma884(d0, d1, a0, b0);
            dmma884(d0, d1, a1, b1);
            __syncwarp();
            if (i < n) {
                dst[0] = -d0;
                dst[ldx] = -d1;
            }
            __syncwarp();
        }
    }
    tsync();
    // ---- A^{-1} = X^T X on the lower tiles, mirrored ----
    {
        const int ntiles = NB * (NB + 1) / 2;
        DWBC_UNROLL(1)
        for (int p = warp; p < ntiles; p += NW) {
            int bi = 0, rem = p;  // row-major over the lower triangle: tile (bi, bj), bj <= bi
            while (rem > bi) {
                rem -= bi + 1;
                bi++;
            }
            const int bj = rem;
            const int i = 8 * bi + g, jg = 8 * bj + g, k0 = 8 * bi;
            const int ic = i < n ? i : n - 1, jc = jg < n ? jg : n - 1;
            const double *pa = X + (k0 + t) + ic * ldx, *pb = X + (k0 + t) + jc * ldx;
            double c0 = 0.0, c1 = 0.0;
            const int KF = (n - k0) >> 2, KR = (n - k0) & 3;
            DWBC_UNROLL(2)
            for (int ks = 0; ks < KF; ks++) {
                dmma884(c0, c1, pa[0], pb[0]);
                pa += 4;
                pb += 4;
            }
            if (KR) {
                const bool kv = t < KR;
                dmma884(c0, c1, kv ? pa[0] : 0.0, kv ? pb[0] : 0.0);
            }
            const int j = 8 * bj + 2 * t;
            if (i < n && j < n && (bi > bj || i >= j)) A[i + j * ld] = c0, A[j + i * ld] = c0;
            if (i < n && j + 1 < n && (bi > bj || i >= j + 1)) A[i + (j + 1) * ld] = c1, A[j + 1 + i * ld] = c1;
        }
    }
    tsync();
}

// SPD inverse of an n x n matrix, n <= 8, by ONE warp (see warp_spd_inverse_small); Xs: n*n doubles private to the warp.
__device__ __noinline__ double warp_spd_inverse_small(double *S, int ld, int n, double *Xs) {
    const int lane = lane_id();
    double a[8], rs[8], x[8];
    DWBC_UNROLL(8)
    for (int k = 0; k < 8; k++) a[k] = (lane < n && k < n) ? S[lane + k * ld] : 0.0, rs[k] = 0.0;
    double dk = 0.0;
    DWBC_UNROLL(8)
    for (int k = 0; k < 8; k++)
        if (lane == k) dk = a[k];
    const double dmax = warp_max(dk);
    double minp = 1e300;
    DWBC_UNROLL(8)
    for (int k = 0; k < 8; k++) {
        if (k < n) {
            double d = __shfl_sync(DWBC_FULL, a[k], k);
            minp = d < minp ? d : minp;
            if (!(d > 1e-300)) d = 1e-300;
            rs[k] = rsqrt(d);
            a[k] *= rs[k];
            if (lane == k) a[k] = d * rs[k];
            DWBC_UNROLL(8)
            for (int j = k + 1; j < 8; j++) {
                const double ljk = __shfl_sync(DWBC_FULL, a[k], j);
                a[j] = fma(-a[k], ljk, a[j]);
            }
        }
    }
    DWBC_UNROLL(8)
    for (int i = 0; i < 8; i++) {
        double s = 0.0;
        x[i] = 0.0;
        if (i < n) {
            DWBC_UNROLL(8)
            for (int k = 0; k < i; k++) {
                const double lik = __shfl_sync(DWBC_FULL, a[k], i);
                s = fma(lik, x[k], s);
            }
            x[i] = (i == lane) ? rs[i] : -s * rs[i];
            if (i < lane) x[i] = 0.0;
        }
    }
    if (lane < n) {
        DWBC_UNROLL(8)
        for (int i = 0; i < 8; i++)
            if (i < n) Xs[i + lane * n] = x[i];
    }
    __syncwarp();
    for (int e = lane; e < n * n; e += 32) {
        const int j = e / n, i = e - j * n;
        if (i < j) continue;
        double s = 0.0;
        for (int k = i; k < n; k++) s = fma(Xs[k + i * n], Xs[k + j * n], s);
        S[i + j * ld] = s;
        S[j + i * ld] = s;
    }
    __syncwarp();
    return dmax > 0.0 ? minp / dmax : 0.0;
}

}  // namespace rt

}  // namespace lean
}  // namespace dwbc
