// Stages a1 (A^{-1}), a2 (contact Jacobian stack), a3 (contact-consistent dynamics) and a4 (gravity
// compensation).  One warp per robot state, operands staged in shared memory.
//
// Reference: RobotData::UpdateKinematics `A_inv_ = A_.llt().solve(I)` (src/dwbc.cpp:307),
// SetContact/UpdateContactConstraint (include/dwbc.h:432-474, src/dwbc.cpp:433-454),
// CalculateContactConstraint + PinvCODWB (src/wbd.cpp:108-143, 32-53),
// CalculateGravityCompensation (src/wbd.cpp:186-192).
//
// The arithmetic is restructured (results identical up to rounding):
//   X = A^{-1} Jc^T,  Lambda_c = (Jc X)^{-1} by Cholesky (the matrix is SPD),  Jbar^T = Lambda_c X^T,
//   A^{-1} N_c = A^{-1} - X Jbar^T                      (rank-c update instead of a 39^3 product),
//   null(W) is known in closed form: {S Jc^T f : (Jc^T f)[0:6] = 0}; with an orthonormal basis V of it
//   W^+ = (W + V V^T)^{-1} - V V^T                      (one SPD inverse instead of a COD),
//   NwJw = V (Jbar^T[0:r, 6:] V)^{-1},  tau_g = W^+ ((A^{-1}N_c)[6:, :] G),  P_C = Jbar^T G.
#pragma once
#include "batch.h"
#include "kernels_rbd.cuh"
#include "team_linalg.cuh"
#include "warp_linalg.cuh"

namespace dwbc {

// ---------------------------------------------------------------- a1: rigid-body dynamics
__global__ void rbd_kernel(const DevModel *__restrict__ md, const double *__restrict__ q, const double *__restrict__ qd, Work w) {
    DWBC_DYN_SMEM(smem);
    const int wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
    const long long s = (long long)blockIdx.x * wpb + wid;
    if (s >= w.B) return;
    double *rb = (double *)smem + (size_t)wid * rbd_smem_doubles(md->nb);
    RbdOut out{w.A, w.G, w.Bnl, w.pose, w.com, w.linkvel};
    rbd_state(md, q + s * (md->n + 1), qd + s * md->n, rb, out, s);
}

// ---------------------------------------------------------------- a1: A^{-1} by Cholesky
// One CTA (team) per state.  A = L L^T with register-resident ownership (team_linalg.cuh), Xi = L^{-1} by forward
// substitution (one column per thread), A^{-1} = Xi^T Xi on the lower triangle, mirrored.
__host__ __device__ inline int ainv_smem_doubles(int n) { return 2 * (n | 1) * n + n + 1; }

template <int MAXOWN>
__device__ __forceinline__ double team_spd_inverse_t(int n, double *L, int ld, double *Xi, int ldx, double *rd) {
    const double minpiv = tchol_lower<MAXOWN>(n, L, ld, rd);
    ttri_inverse(n, L, ld, rd, Xi, ldx);
    tsync();
    tsyrk_inverse(n, Xi, ldx, [&](int i, int j, double v) { L[i + j * ld] = v; });
    tsync();
    return minpiv;
}
// SPD inverse in place: on return the full symmetric inverse is in L (n x n, ld); Xi is scratch of the same size.
// Synchronises on entry and exit.  Returns the smallest Cholesky pivot.
__device__ __forceinline__ double team_spd_inverse(int n, double *L, int ld, double *Xi, int ldx, double *rd) {
    const int nt = t_n();
    const int slots = ((n + 1) >> 1) * (n + 1);  // ownership slots of tchol_lower
    if (slots <= 2 * nt) return team_spd_inverse_t<2>(n, L, ld, Xi, ldx, rd);
    if (slots <= 5 * nt) return team_spd_inverse_t<5>(n, L, ld, Xi, ldx, rd);  // n <= 33 with 128 threads
    if (slots <= 7 * nt) return team_spd_inverse_t<7>(n, L, ld, Xi, ldx, rd);
    if constexpr (TEAM >= 128) {
        return team_spd_inverse_t<12>(n, L, ld, Xi, ldx, rd);  // n <= 54
    } else {
        if (slots <= 12 * nt) return team_spd_inverse_t<12>(n, L, ld, Xi, ldx, rd);
        return team_spd_inverse_t<45>(n, L, ld, Xi, ldx, rd);  // small teams (emulation builds of the tests)
    }
}

__global__ void __launch_bounds__(TEAM, 6) ainv_kernel(Work w) {
    DWBC_DYN_SMEM(smem);
    const long long s = blockIdx.x;
    const int n = w.n, ld = n | 1;
    double *L = (double *)smem;
    double *X = L + ld * n, *rd = X + ld * n;
    tcopy(n, n, w.A + s * n * n, n, L, ld);
    const double minpiv = team_spd_inverse(n, L, ld, X, ld, rd);
    tcopy(n, n, L, ld, w.Ainv + s * n * n, n);
    if (t_id() == 0 && !(minpiv > 0.0)) w.status[s] = 0;  // mass matrix not SPD: everything downstream is invalid
}

// ---------------------------------------------------------------- a2: contact Jacobian stack
__global__ void contact_jacobian_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, const uint8_t *__restrict__ flags,
                                        Work w) {
    const int wpb = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = lane_id();
    const long long s = (long long)blockIdx.x * wpb + wid;
    if (s >= w.B) return;
    const int n = w.n;
    double *Jc = w.Jc + s * CMAX * n;
    for (int e = lane; e < CMAX * n; e += 32) Jc[e] = 0.0;
    __syncwarp();
    int k = 0, over = 0;
    for (int i = 0; i < cfg->ncontact; i++) {
        if (!flags[s * cfg->ncontact + i]) continue;
        if (k >= MAX_ACTIVE_CONTACTS) {
            over = 1;
            break;
        }
        const ContactCfg &c = cfg->contacts[i];
        point_jacobian(md, w.pose + s * md->nb * POSE_STRIDE, c.link, v3(c.point[0], c.point[1], c.point[2]), Jc, CMAX, 6 * k, 0, 6);
        if (lane == 0) w.cidx[s * MAX_ACTIVE_CONTACTS + k] = i;
        k++;
    }
    if (lane == 0) {
        for (int j = k; j < MAX_ACTIVE_CONTACTS; j++) w.cidx[s * MAX_ACTIVE_CONTACTS + j] = -1;
        w.cdof[s] = over ? -1 : 6 * k;
    }
}

// ---------------------------------------------------------------- a3 + a4
// shared-memory arena (doubles) of one state
struct ContactSmem {
    int ldn, ldc, ldm;
    int P, Q, Jc, X, Jbar, Lam, Lw, V, Aug, vec, total;
    __host__ __device__ ContactSmem(int n, int m) {
        ldn = n | 1; ldc = CMAX | 1; ldm = m | 1;
        int o = 0;
        P = o; o += ldn * n;        // A^{-1} -> A^{-1}N_c ; its [6:,6:] block later holds W + V V^T, then W^+
        // Q (scratch of the W + V V^T inverse) is only live after Jc, X, Lambda_c and Lw are dead: it shares their storage,
        // which brings the arena under 36.8 KB = six states per SM instead of five
        Q = o;
        Jc = o; o += ldc * n;
        X = o; o += ldn * CMAX;
        Lam = o; o += ldc * CMAX;   // Jc X -> Lambda_c
        Lw = o; o += ldc * CMAX;    // scratch of the c x c inverse
        if (o - Q < ldm * m) o = Q + ldm * m;
        Jbar = o; o += ldc * n;
        V = o; o += ldm * (RMAX > 0 ? RMAX : 1);
        Aug = o; o += 13 * 32;      // Gauss-Jordan scratch, up to 12 rows x 32 columns, ld 13
        vec = o; o += 4 * (n + 1);  // rdiag / G / u / flag word
        total = o;
    }
};

// Upper Cholesky factor G = R^T R of an r x r SPD matrix (r <= 7) by ONE warp: lane e owns entry (i, j), i <= j, e = j(j+1)/2 + i,
// in a register; step k scales row k by 1/sqrt(g_kk) and applies g_ij -= R_ki R_kj (i > k) with two shuffles.  Per entry this is the
// sequence  sacc = fma(-R_ki, R_kj, sacc), k ascending, then sqrt or the division  of the scalar loop it replaces.
// Returns false if a pivot is not positive (that pivot is replaced by 1, like the scalar loop).  All 32 lanes must call it.
__device__ __forceinline__ bool wchol_upper_packed(int r, double *G, int ldg, int lane) {
    int j = 0;
    while ((j + 1) * (j + 2) / 2 <= lane) j++;
    const int i = lane - j * (j + 1) / 2;
    const bool own = lane < r * (r + 1) / 2;
    double g = own ? G[i + j * ldg] : 0.0;
    bool ok = true;
    for (int k = 0; k < r; k++) {
        double d2 = __shfl_sync(DWBC_FULL, g, k * (k + 1) / 2 + k);
        if (!(d2 > 1e-280)) {
            d2 = 1.0;
            ok = false;
        }
        const double d = sqrt(d2);
        if (own && i == k) g = (j == k) ? d : g / d;
        const bool below = own && i > k;
        const double rki = __shfl_sync(DWBC_FULL, g, below ? i * (i + 1) / 2 + k : 0);
        const double rkj = __shfl_sync(DWBC_FULL, g, below ? j * (j + 1) / 2 + k : 0);
        if (below) g = fma(-rki, rkj, g);
    }
    if (own) G[i + j * ldg] = g;
    return ok;
}

__global__ void __launch_bounds__(TEAM, 6) contact_dynamics_kernel(const DevModel *__restrict__ md, Work w) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n(), warp = tid >> 5, lane = lane_id();
    const long long s = blockIdx.x;
    const int n = w.n, m = w.m;
    const ContactSmem L(n, m);
    double *sm = (double *)smem;
    double *P = sm + L.P, *Q = sm + L.Q, *Jc = sm + L.Jc, *X = sm + L.X, *Jbar = sm + L.Jbar, *Lam = sm + L.Lam, *Lw = sm + L.Lw;
    double *V = sm + L.V, *Aug = sm + L.Aug, *rd = sm + L.vec, *Gs = rd + n + 1, *u = Gs + n + 1;
    int &s_ok = *(int *)(u + n + 1);  // CTA-wide validity flag (lives in the arena: no static shared memory)
    const int ldn = L.ldn, ldc = L.ldc, ldm = L.ldm;
    const int c = w.cdof[s];
    const int r = c > 6 ? c - 6 : 0;
    if (!(c > 0 && c <= CMAX) || w.status[s] == 0) {  // no contact / too many contacts / A not SPD (uniform over the CTA)
        if (tid == 0) w.status[s] = 0;
        return;
    }
    if (tid == 0) s_ok = 1;
    tcopy(n, n, w.Ainv + s * n * n, n, P, ldn);
    tcopy(c, n, w.Jc + s * CMAX * n, CMAX, Jc, ldc);
    for (int e = tid; e < n; e += nt) Gs[e] = w.G[s * n + e];
    tsync();
    // X = A^{-1} Jc^T  (n x c)
    tgemm_mma<2, false, true>(n, c, n, P, ldn, Jc, ldc, [&](int i, int j, double v) { X[i + j * ldn] = v; });
    tsync();
    // Jc X (c x c, SPD) -> Lambda_c (in Lam)
    tgemm_mma<1, false, false>(c, c, n, Jc, ldc, X, ldn, [&](int i, int j, double v) { Lam[i + j * ldc] = v; });
    const double piv_c = team_spd_inverse(c, Lam, ldc, Lw, ldc, rd);
    if (!(piv_c > 0.0) && tid == 0) s_ok = 0;
    // Jbar^T = Lambda_c X^T  (c x n)
    tgemm_mma<2, false, true>(c, n, c, Lam, ldc, X, ldn, [&](int i, int j, double v) { Jbar[i + j * ldc] = v; });
    tsync();
    // A^{-1} N_c = A^{-1} - X Jbar^T
    tgemm_mma<2, false, false>(n, n, c, X, ldn, Jbar, ldc, [&](int i, int j, double v) { P[i + j * ldn] -= v; });
    tsync();
    // outputs of a3 that do not depend on the pseudo-inverse
    {
        double *g = w.Lam + s * CMAX * CMAX;
        for (int e = tid; e < CMAX * CMAX; e += nt) {
            int j = e / CMAX, i = e - j * CMAX;
            g[e] = (i < c && j < c) ? Lam[i + j * ldc] : 0.0;
        }
        g = w.Jbar + s * CMAX * n;
        for (int e = tid; e < CMAX * n; e += nt) {
            int j = e / CMAX, i = e - j * CMAX;
            g[e] = (i < c) ? Jbar[i + j * ldc] : 0.0;
        }
        tcopy(n, n, P, ldn, w.AinvNc + s * n * n, n);
        tcopy(m, m, P + 6 + 6 * ldn, ldn, w.W + s * m * m, m);
    }
    // u = (A^{-1}N_c)[6:, :] G   and   P_C = Jbar^T G
    // With a null space to build (r > 0) warp 0 runs the one-warp Gauss-Jordan steps below while the other warps do the
    // independent products / copies of the same phase; without one the whole team does them.
    const bool split = r > 0 && nt >= 64;
    const int stid = split ? tid - 32 : tid, snt = split ? nt - 32 : nt;
    if (!(split && warp == 0)) {
        tgemv<false>(m, n, P + 6, ldn, Gs, [&](int i, double v) { u[i] = v; }, stid, snt);
        tgemv<false>(c, n, Jbar, ldc, Gs, [&](int i, double v) { w.PC[s * CMAX + i] = v; }, stid, snt);
        for (int i = c + stid; i < CMAX; i += snt) w.PC[s * CMAX + i] = 0.0;
    }
    double *Wblk = P + 6 + 6 * ldn;  // view of W inside P, leading dimension ldn
    const int lda = 13;
    if (r > 0) {
        // ---- orthonormal basis V (m x r) of null(W) ----
        // T = B_1^{-T} [B_2^T ...]  with  B_i = Jc[6i:6i+6, 0:6] : Gauss-Jordan on [B_1^T | B_2^T ...]  (warp 0)
        if (warp == 0) {
            for (int e = lane; e < 6 * (6 + r); e += 32) {
                int col = e / 6, row = e - col * 6;
                Aug[row + col * lda] = Jc[col + row * ldc];
            }
            __syncwarp();
            const double cond1 = wgauss_jordan(6, r, Aug, lda);
            if (!(cond1 > 1e-14) && lane == 0) s_ok = 0;
        }
        tsync();
        // V_raw(a, b) = Jc[6+b][6+a] - sum_e Jc[e][6+a] * T(e, b)
        for (int e = tid; e < m * r; e += nt) {
            int b = e / m, a = e - b * m;
            double v = Jc[(6 + b) + (6 + a) * ldc];
            for (int k = 0; k < 6; k++) v = fma(-Jc[k + (6 + a) * ldc], Aug[k + (6 + b) * lda], v);
            V[a + b * ldm] = v;
        }
        tsync();
        // orthonormalise V (m x r) by Cholesky-QR, twice: G = V^T V = R^T R, V <- V R^{-1}.  (Two-pass modified Gram-Schmidt
        // on one warp was a fifth of this kernel's stall samples: three warps waited for r(r-1) dependent warp reductions.)
        for (int pass = 0; pass < 2; pass++) {
            for (int e = tid; e < r * r; e += nt) {
                const int j = e / r, i = e - j * r;
                if (i <= j) {
                    double d = 0.0;
                    for (int a = 0; a < m; a++) d = fma(V[a + i * ldm], V[a + j * ldm], d);
                    Aug[i + j * lda] = d;
                }
            }
            tsync();
            if (r <= 7) {
                // r(r+1)/2 <= 28 entries: one lane per entry, factor in registers (the one-thread loop below held three warps at
                // the barrier for ~12 % of the kernel); same fma order per entry, so the factor is bit-identical
                if (warp == 0 && !wchol_upper_packed(r, Aug, lda, lane) && lane == 0) s_ok = 0;
            } else if (tid == 0) {
                for (int j = 0; j < r; j++)
                    for (int i = 0; i <= j; i++) {
                        double sacc = Aug[i + j * lda];
                        for (int k = 0; k < i; k++) sacc = fma(-Aug[k + i * lda], Aug[k + j * lda], sacc);
                        if (i == j) {
                            if (!(sacc > 1e-280)) {
                                sacc = 1.0;
                                s_ok = 0;
                            }
                            Aug[j + j * lda] = sqrt(sacc);
                        } else {
                            Aug[i + j * lda] = sacc / Aug[i + i * lda];
                        }
                    }
            }
            tsync();
            for (int a = tid; a < m; a += nt)
                for (int j = 0; j < r; j++) {
                    double x = V[a + j * ldm];
                    for (int k = 0; k < j; k++) x = fma(-V[a + k * ldm], Aug[k + j * lda], x);
                    V[a + j * ldm] = x / Aug[j + j * lda];
                }
            tsync();
        }
        // V2 (rows = basis vectors), reference layout r x m
        {
            double *g = w.V2 + s * RMAX * m;
            for (int e = tid; e < RMAX * m; e += nt) {
                int a = e / RMAX, b = e - a * RMAX;
                g[e] = b < r ? V[a + b * ldm] : 0.0;
            }
        }
        // W + V V^T in place
        tgemm_mma<2, false, true>(m, m, r, V, ldm, V, ldm, [&](int i, int j, double v) { Wblk[i + j * ldn] += v; });
    } else {
        double *g = w.V2 + s * RMAX * m;
        for (int e = tid; e < RMAX * m; e += nt) g[e] = 0.0;
    }
    // (W + V V^T)^{-1}, in place
    const double piv_w = team_spd_inverse(m, Wblk, ldn, Q, ldm, rd);
    if (!(piv_w > 0.0) && tid == 0) s_ok = 0;
    if (r > 0) {
        tgemm_mma<2, false, true>(m, m, r, V, ldm, V, ldm, [&](int i, int j, double v) { Wblk[i + j * ldn] -= v; });
        // Mx = Jbar^T[0:r, 6:6+m] V  (r x r);  NwJw = V Mx^{-1}
        tgemm<1, 1, false, false>(r, r, m, Jbar + 6 * ldc, ldc, V, ldm, [&](int i, int j, double v) { Aug[i + j * lda] = v; });
        for (int e = tid; e < r * r; e += nt) {
            int j = e / r, i = e - j * r;
            Aug[i + (r + j) * lda] = (i == j) ? 1.0 : 0.0;
        }
        tsync();
        if (warp == 0) {
            const double cond2 = wgauss_jordan(r, r, Aug, lda);
            if (!(cond2 > 1e-14) && lane == 0) s_ok = 0;
        }
        if (!(split && warp == 0)) {  // W^+ is final: its copy and tau_g = W^+ u do not wait for Mx^{-1}
            tcopy(m, m, Wblk, ldn, w.Winv + s * m * m, m, stid, snt);
            tgemv<false>(m, m, Wblk, ldn, u, [&](int i, double v) { w.tau_grav[s * m + i] = v; }, stid, snt);
        }
        tsync();
        double *g = w.NwJw + s * m * RMAX;
        tgemm<1, 2, false, false>(m, r, r, V, ldm, Aug + r * lda, lda, [&](int i, int j, double v) { g[i + j * m] = v; });
        for (int e = m * r + tid; e < m * RMAX; e += nt) g[e] = 0.0;
    } else {
        double *g = w.NwJw + s * m * RMAX;
        for (int e = tid; e < m * RMAX; e += nt) g[e] = 0.0;
        tsync();
        tcopy(m, m, Wblk, ldn, w.Winv + s * m * m, m);
        // tau_g = W^+ u
        tgemv<false>(m, m, Wblk, ldn, u, [&](int i, double v) { w.tau_grav[s * m + i] = v; });
    }
    tsync();
    if (tid == 0) w.status[s] = s_ok ? 1 : 0;
}

// N_C = I - Jc^T Jbar^T, materialised only when a caller asks for the public field (dwbc_batch_get_matrix)
__global__ void nc_kernel(Work w, double *out) {
    const int wpb = blockDim.x >> 5, wid = threadIdx.x >> 5;
    const long long s = (long long)blockIdx.x * wpb + wid;
    if (s >= w.B) return;
    const int n = w.n;
    const int c = w.cdof[s] > 0 ? w.cdof[s] : 0;
    const double *Jc = w.Jc + s * CMAX * n, *Jb = w.Jbar + s * CMAX * n;
    double *o = out + s * n * n;
    wgemm<4, 4, true, false>(n, n, c, Jc, CMAX, Jb, CMAX, [&](int i, int j, double v) { o[i + j * n] = (i == j ? 1.0 : 0.0) - v; });
}

}  // namespace dwbc

namespace dwbc {
// Export of kinematic public fields: mode 0 com_pos (3), 1 link_[i].xpos (3 x nb), 2 link_[i].rotm
// (9 x nb, column-major 3x3 per link; reference src/link.cpp:78-79).
__global__ void pose_export_kernel(Work w, int mode, double *out) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int nb = w.nb;
    if (mode == 0) {
        if (tid < (long long)w.B * 3) out[tid] = w.com[(tid / 3) * 16 + tid % 3];
        return;
    }
    for (long long e = tid; e < (long long)w.B * nb; e += (long long)gridDim.x * blockDim.x) {
        const double *po = w.pose + e * POSE_STRIDE;
        if (mode == 1) {
            for (int k = 0; k < 3; k++) out[e * 3 + k] = po[9 + k];
        } else {
            for (int a = 0; a < 3; a++)
                for (int c = 0; c < 3; c++) out[e * 9 + c * 3 + a] = po[a * 3 + c];
        }
    }
}
}  // namespace dwbc
