// TEST INFRASTRUCTURE: unit harness for libdwbc_b200/csrc/lean_linalg.cuh compiled through the CUDA emulation
// (tests/emul/cuda_emul.h).  Exposes a few C entry points that tests/test_lean_linalg.py drives with numpy inputs.
#include "cuda_compat.h"
#include "batch.h"
#include "lean_linalg.cuh"

using namespace dwbc;

template <int N, int LD, int LDX>
__global__ void k_spd_inverse(const double *in, double *out, double *piv) {
    DWBC_DYN_SMEM(smem);
    double *A = (double *)smem, *X = A + LD * N + 8, *rd = X + LDX * N + 8, *ps = rd + N + 2;
    for (int e = t_id(); e < N * N; e += t_n()) A[(e % N) + (e / N) * LD] = in[e];
    lean::spd_inverse<N, LD, LDX>(A, X, rd, ps);
    for (int e = t_id(); e < N * N; e += t_n()) out[e] = A[(e % N) + (e / N) * LD];
    if (t_id() == 0) *piv = *ps;
}
template <int N, int LD, int LDX>
static void run_spd(const double *in, double *out, double *piv) {
    DWBC_LAUNCH((k_spd_inverse<N, LD, LDX>), 1, TEAM, sizeof(double) * (LD * N + LDX * N + N + 32), nullptr, in, out, piv);
}

template <int N, int LD>
__global__ void k_small_inverse(const double *in, double *out, double *cond) {
    DWBC_DYN_SMEM(smem);
    double *S = (double *)smem, *Xs = S + LD * N + 8;
    for (int e = t_id(); e < N * N; e += t_n()) S[(e % N) + (e / N) * LD] = in[e];
    tsync();
    if (lean::warp_id() == 1) {
        const double c = lean::warp_spd_inverse_small<N, LD>(S, Xs);
        if (lane_id() == 0) *cond = c;
    }
    tsync();
    for (int e = t_id(); e < N * N; e += t_n()) out[e] = S[(e % N) + (e / N) * LD];
}

template <int M, int N, int K, bool TA, bool TB, int NT>
__global__ void k_gemm(const double *a, const double *b, double *c) {
    DWBC_DYN_SMEM(smem);
    constexpr int LDA = (TA ? K : M) + 3, LDB = (TB ? N : K) + 5;
    double *A = (double *)smem, *B = A + LDA * (TA ? M : K) + 4;
    for (int e = t_id(); e < M * K; e += t_n()) {
        const int i = e % M, k = e / M;
        A[TA ? k + i * LDA : i + k * LDA] = a[e];
    }
    for (int e = t_id(); e < K * N; e += t_n()) {
        const int k = e % K, j = e / K;
        B[TB ? j + k * LDB : k + j * LDB] = b[e];
    }
    tsync();
    lean::mma_gemm<M, N, K, LDA, LDB, TA, TB, NT>(A, B, [&](int i, int j, double v) { c[i + j * M] = v; }, 1);
}
template <int M, int N, int K, bool TA, bool TB, int NT>
static void run_gemm(const double *a, const double *b, double *c) {
    DWBC_LAUNCH((k_gemm<M, N, K, TA, TB, NT>), 1, TEAM, sizeof(double) * ((M + K + 8) * (K + N + 8) * 2), nullptr, a, b, c);
}

template <int N, int K>
__global__ void k_gemm_sym(const double *a, double *c) {  // C = A A^T, A is N x K
    DWBC_DYN_SMEM(smem);
    constexpr int LDA = N + 3;
    double *A = (double *)smem;
    for (int e = t_id(); e < N * K; e += t_n()) A[(e % N) + (e / N) * LDA] = a[e];
    tsync();
    lean::mma_gemm_sym<N, K, LDA, LDA, false, true>(A, A, [&](int i, int j, double v) { c[i + j * N] = v; c[j + i * N] = v; });
}

template <int M, int K, bool TA, int LPR>
__global__ void k_gemv(const double *a, const double *x, double *y) {
    DWBC_DYN_SMEM(smem);
    constexpr int LDA = (TA ? K : M) + 1;
    double *A = (double *)smem, *xs = A + LDA * (TA ? M : K) + 2;
    for (int e = t_id(); e < M * K; e += t_n()) {
        const int i = e % M, k = e / M;
        A[TA ? k + i * LDA : i + k * LDA] = a[e];
    }
    for (int e = t_id(); e < K; e += t_n()) xs[e] = x[e];
    tsync();
    lean::gemv<M, K, LDA, TA, LPR>(A, xs, [&](int i, double v) { y[i] = v; });
}

// ---- run-time-dimension twins (lean::rt) ----
__global__ void k_rt_spd_inverse(int n, int ld, int ldx, const double *in, double *out, double *piv) {
    DWBC_DYN_SMEM(smem);
    double *A = (double *)smem, *X = A + ld * n + 8, *rd = X + ldx * n + 8, *ps = rd + n + 2;
    for (int e = t_id(); e < n * n; e += t_n()) A[(e % n) + (e / n) * ld] = in[e];
    lean::rt::spd_inverse(A, ld, X, ldx, n, rd, ps, 1);
    for (int e = t_id(); e < n * n; e += t_n()) out[e] = A[(e % n) + (e / n) * ld];
    if (t_id() == 0) *piv = *ps;
}
__global__ void k_rt_small_inverse(int n, int ld, const double *in, double *out, double *cond) {
    DWBC_DYN_SMEM(smem);
    double *S = (double *)smem, *Xs = S + ld * n + 8;
    for (int e = t_id(); e < n * n; e += t_n()) S[(e % n) + (e / n) * ld] = in[e];
    tsync();
    if (lean::warp_id() == 2) {
        const double c = lean::rt::warp_spd_inverse_small(S, ld, n, Xs);
        if (lane_id() == 0) *cond = c;
    }
    tsync();
    for (int e = t_id(); e < n * n; e += t_n()) out[e] = S[(e % n) + (e / n) * ld];
}
// C = alpha op(A) op(B) + beta C0, operands stored with transposes ta / tb
__global__ void k_rt_gemm(int M, int N, int K, int ta, int tb, int nt, double alpha, int beta, const double *a, const double *b, double *c) {
    DWBC_DYN_SMEM(smem);
    const int lda = (ta ? K : M) + 3, ldb = (tb ? N : K) + 5, ldc = M + 1;
    double *A = (double *)smem, *B = A + lda * (ta ? M : K) + 4, *Cm = B + ldb * (tb ? K : N) + 4;
    for (int e = t_id(); e < M * K; e += t_n()) {
        const int i = e % M, k = e / M;
        A[ta ? k + i * lda : i + k * lda] = a[e];
    }
    for (int e = t_id(); e < K * N; e += t_n()) {
        const int k = e % K, j = e / K;
        B[tb ? j + k * ldb : k + j * ldb] = b[e];
    }
    for (int e = t_id(); e < M * N; e += t_n()) Cm[(e % M) + (e / M) * ldc] = c[e];
    tsync();
    lean::rt::gemm(M, N, K, A, ta ? lda : 1, ta ? 1 : lda, B, tb ? ldb : 1, tb ? 1 : ldb, Cm, ldc, alpha, beta, nt, 3);
    tsync();
    for (int e = t_id(); e < M * N; e += t_n()) c[e] = Cm[(e % M) + (e / M) * ldc];
}
__global__ void k_rt_gemv(int M, int K, int ta, double alpha, int usez, const double *a, const double *x, double *y) {
    DWBC_DYN_SMEM(smem);
    const int lda = (ta ? K : M) + 1;
    double *A = (double *)smem, *xs = A + lda * (ta ? M : K) + 2, *zs = xs + K + 2, *ys = zs + M + 2;
    for (int e = t_id(); e < M * K; e += t_n()) {
        const int i = e % M, k = e / M;
        A[ta ? k + i * lda : i + k * lda] = a[e];
    }
    for (int e = t_id(); e < K; e += t_n()) xs[e] = x[e];
    for (int e = t_id(); e < M; e += t_n()) zs[e] = y[e];
    tsync();
    lean::rt::gemv(M, K, A, ta ? lda : 1, ta ? 1 : lda, xs, ys, alpha, usez ? zs : nullptr, t_id(), t_n());
    tsync();
    for (int e = t_id(); e < M; e += t_n()) y[e] = ys[e];
}

extern "C" {
int lean_rt_spd_inverse(int n, int ld, int ldx, const double *in, double *out, double *piv) {
    DWBC_LAUNCH(k_rt_spd_inverse, 1, TEAM, sizeof(double) * (ld * n + ldx * n + n + 64), nullptr, n, ld, ldx, in, out, piv);
    return 0;
}
int lean_rt_small_inverse(int n, int ld, const double *in, double *out, double *cond) {
    DWBC_LAUNCH(k_rt_small_inverse, 1, TEAM, 8192, nullptr, n, ld, in, out, cond);
    return 0;
}
int lean_rt_gemm(int M, int N, int K, int ta, int tb, int nt, double alpha, int beta, const double *a, const double *b, double *c) {
    DWBC_LAUNCH(k_rt_gemm, 1, TEAM, sizeof(double) * 3 * (M + K + N + 16) * (M + K + N + 16), nullptr, M, N, K, ta, tb, nt, alpha, beta, a, b, c);
    return 0;
}
int lean_rt_gemv(int M, int K, int ta, double alpha, int usez, const double *a, const double *x, double *y) {
    DWBC_LAUNCH(k_rt_gemv, 1, TEAM, sizeof(double) * ((M + 4) * (K + 4) + 2 * M + K + 32), nullptr, M, K, ta, alpha, usez, a, x, y);
    return 0;
}
int lean_spd_inverse(int n, const double *in, double *out, double *piv) {
    switch (n) {
    case 39: run_spd<39, 44, 44>(in, out, piv); return 0;
    case 40: run_spd<40, 40, 44>(in, out, piv); return 0;
    case 33: run_spd<33, 44, 36>(in, out, piv); return 0;
    case 21: run_spd<21, 28, 28>(in, out, piv); return 0;
    case 12: run_spd<12, 12, 12>(in, out, piv); return 0;
    case 24: run_spd<24, 28, 28>(in, out, piv); return 0;
    case 18: run_spd<18, 20, 20>(in, out, piv); return 0;
    case 6: run_spd<6, 12, 12>(in, out, piv); return 0;
    case 9: run_spd<9, 12, 12>(in, out, piv); return 0;
    }
    return -1;
}
int lean_small_inverse(int n, const double *in, double *out, double *cond) {
    switch (n) {
    case 3: DWBC_LAUNCH((k_small_inverse<3, 4>), 1, TEAM, 4096, nullptr, in, out, cond); return 0;
    case 6: DWBC_LAUNCH((k_small_inverse<6, 12>), 1, TEAM, 4096, nullptr, in, out, cond); return 0;
    case 8: DWBC_LAUNCH((k_small_inverse<8, 12>), 1, TEAM, 4096, nullptr, in, out, cond); return 0;
    case 1: DWBC_LAUNCH((k_small_inverse<1, 4>), 1, TEAM, 4096, nullptr, in, out, cond); return 0;
    }
    return -1;
}
int lean_gemm(int which, const double *a, const double *b, double *c) {
    switch (which) {
    case 0: run_gemm<30, 39, 39, false, false, 5>(a, b, c); return 0;
    case 1: run_gemm<39, 12, 39, false, true, 2>(a, b, c); return 0;
    case 2: run_gemm<12, 39, 12, true, false, 1>(a, b, c); return 0;
    case 3: run_gemm<33, 30, 33, true, true, 2>(a, b, c); return 0;
    case 4: run_gemm<3, 6, 33, false, false, 1>(a, b, c); return 0;
    case 5: run_gemm<33, 21, 9, false, false, 3>(a, b, c); return 0;
    case 6: run_gemm<20, 27, 33, false, false, 2>(a, b, c); return 0;
    }
    return -1;
}
int lean_gemm_sym(int which, const double *a, double *c) {
    switch (which) {
    case 0: DWBC_LAUNCH((k_gemm_sym<33, 6>), 1, TEAM, 65536, nullptr, a, c); return 0;
    case 1: DWBC_LAUNCH((k_gemm_sym<39, 12>), 1, TEAM, 65536, nullptr, a, c); return 0;
    case 2: DWBC_LAUNCH((k_gemm_sym<21, 33>), 1, TEAM, 65536, nullptr, a, c); return 0;
    }
    return -1;
}
int lean_gemv(int which, const double *a, const double *x, double *y) {
    switch (which) {
    case 0: DWBC_LAUNCH((k_gemv<33, 39, false, 2>), 1, TEAM, 65536, nullptr, a, x, y); return 0;
    case 1: DWBC_LAUNCH((k_gemv<12, 39, false, 4>), 1, TEAM, 65536, nullptr, a, x, y); return 0;
    case 2: DWBC_LAUNCH((k_gemv<33, 21, true, 1>), 1, TEAM, 65536, nullptr, a, x, y); return 0;
    case 3: DWBC_LAUNCH((k_gemv<70, 7, false, 2>), 1, TEAM, 65536, nullptr, a, x, y); return 0;
    }
    return -1;
}
}
