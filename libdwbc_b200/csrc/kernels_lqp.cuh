// Stage a10: the LQP / centroidal simplified-WBD variant (Ahn 2025) -- a null-space hierarchy of QPs over
// y = [qdd (n); Fc (c)], one CTA (team) per robot state.
//
// Reference: RobotData::ConfigureLQP (src/dwbc.cpp:4304-4430), CalcControlTorqueLQP (:4432-4454),
// HQP::prepare / solveSequentialSingle (src/dwbc_hqp.cpp:23-85, 291-395), HQP_Hierarch::normalizeConstraintMatrix /
// solveOSQP (:555-581, 583-631), getNullSpace (src/math.cpp:349-360).
//
// Hierarchy (rows of every block are normalised, src/dwbc_hqp.cpp:555-581):
//   level 0  eq   A[0:6,:] qdd + Jc^T[0:6,:] Fc + B[0:6] = 0            preset answer y0 = [-A^{-1} B; 0]   (:4343-4381)
//            ineq |A[6:,:] qdd + Jc^T[6:,:] Fc + B[6:]| <= 200          carried as hard rows, never solved
//   level 1  eq   Jc qdd = 0;  ineq friction/ZMP rows on Fc, |qdd_joint| <= 5;  cost 5 A/|A|_F on qdd   (:4383-4407, :4341)
//   level 2+ eq   J_task qdd = f*  (one level per task space), same cost                              (:4409-4425)
// Level i >= 1 (src/dwbc_hqp.cpp:291-395):  Z_i = Z_{i-1} null(B_i Z_{i-1}),  QP over [u; v]:
//     min 1/2 |B_i (y + Z u) + b_i|^2 + 1/2 |v|^2 + 1/2 (y + Z u)^T H (y + Z u)
//     s.t. A_i (y + Z u) + a_i <= v,   A_j (y + Z u) + a_j <= v_j (frozen) for j < i;      y <- y + Z u
// The reference hands each level to OSQP (ADMM, eps 1e-3, sigma 1e-6): PARITY UNPINNED.  Here -- as in
// the oracle -- every level is solved EXACTLY with OSQP's sigma on the null-space coordinates u, which makes each level
// strictly convex, so y, v do not depend on the choice of the orthonormal basis Z.
//
// Restructuring: Z is kept as one nvar x nvar orthogonal matrix updated in place by the Householder reflections of
// (B_i Z)^T; a level's window is its trailing columns.  With Lc Lc^T = Z^T(B^T B + H)Z + sigma I the level becomes a
// least-distance problem in xh = Lc^T u + Lc^{-1} g with rows G = (A Z) Lc^{-T}; the +/- row pairs share one stored row.
// The slack variables never appear: a slack row's extended normal is [g_i; -e_i], so the Gram matrix of the active
// normals is G_W G_W^T + D (D = 1 on slack rows), v_i equals the row's multiplier, and the dual active-set method
// (Goldfarb-Idnani) runs on nu <= 45 coordinates with an upper-triangular factor R^T R = Gram updated per add / drop and
// one step of iterative refinement per direction (corrected semi-normal equations).
#pragma once
#include "batch.h"
#include "kernels_hqp.cuh"
#include "kernels_task.cuh"
#include "team_linalg.cuh"

namespace dwbc {

constexpr int LQP_WCAP = 48;  // largest active set (>= nu_max = n + CMAX - 6 for TOCABI)
constexpr double LQP_SIGMA = 1.0e-6;
constexpr double LQP_TORQUE_LIMIT = 200.0, LQP_ACC_LIMIT = 5.0;  // reference src/dwbc.cpp:4357-4371, 4395-4403
constexpr int LQP_MAX_LEVELS = MAX_LEVELS + 2;
constexpr int LQP_CHOL_OWN = 9;  // 45*46/2 entries of the lower triangle over 128 threads
constexpr int LQP_NU_TILES = (LQP_WCAP + 7) / 8;  // column tiles of a level's row matrix (nu <= LQP_WCAP)

struct LqpOut {
    double *y;      // [B][LQP_MAX_LEVELS][n + CMAX]   answer after every level
    double *v1;     // [B][2m + 10*MAX_ACTIVE_CONTACTS] slack of the level-1 inequality rows
    double *tau;    // [B][m]   A[6:,:] qdd + Jc^T[6:,:] Fc + B[6:]
    int *status;    // [B]  number of levels solved (== levels: ok), 0 = contact constraint failed
    int *iters;     // [B]  active-set iterations
};

// What differs between ConfigureLQP and ConfigureLQP_R besides the matrices in `Work` (reference src/dwbc.cpp:4341, 4357-4371 vs
// :4519, 4529-4531): the torque bound of every joint row and the matrix whose Frobenius norm scales the cost.
struct LqpParams {
    const double *anorm_src;  // [B][anorm_n * anorm_n]  |A_|_F is always the FULL mass matrix
    int anorm_n;
    double tlim[MAXB];        // torque bound of row i (200; ConfigureLQP_R: 600 on row m-4)
};

struct LqpSmem {
    int nvmax, numax, urmax, rtmax, ldz, ldg, ldh, lde, ldr, ldt;
    int Z, G, H, E, R, vec, total;
    int y, u, xh, zx, gh, gg, ee, hv_tau, ta, ub, rin, tol, vfro, lam, gv, hv, rv, dv, bw, red, cm, ints;
    __host__ __device__ LqpSmem(int n, int m, int tmax) {
        nvmax = n + CMAX;
        numax = nvmax - 6;
        urmax = 2 * m + 10 * MAX_ACTIVE_CONTACTS;
        rtmax = 4 * m + 10 * MAX_ACTIVE_CONTACTS;
        ldz = nvmax | 1;
        ldg = urmax | 1;
        ldh = numax | 1;
        const int emax = tmax > CMAX ? tmax : CMAX;
        lde = emax | 1;
        ldr = LQP_WCAP | 1;
        ldt = n | 1;
        int o = 0;
        Z = o; o += ldz * nvmax;
        G = o; o += ldg * numax;
        H = o; o += ldh * numax;
        E = o; o += lde * nvmax;  // level 0 uses all nvar columns
        int r = ldr * LQP_WCAP, t = ldt * numax, j = lde * n;
        r = r > t ? r : t;
        r = r > j ? r : j;
        R = o; o += r;  // also T = A Z_a and J_task while a level is assembled
        vec = o;
        y = o; o += nvmax + 1;
        u = o; o += nvmax + 1;
        xh = o; o += nvmax + 1;
        zx = o; o += nvmax + 1;
        gh = o; o += nvmax + 1;
        gg = o; o += nvmax + 1;
        ee = o; o += lde + 1;
        hv_tau = o; o += lde + 1;
        ta = o; o += n + 1;
        ub = o; o += rtmax;
        rin = o; o += rtmax;
        tol = o; o += rtmax;
        vfro = o; o += rtmax;
        lam = o; o += LQP_WCAP;
        gv = o; o += LQP_WCAP;
        hv = o; o += LQP_WCAP;
        rv = o; o += LQP_WCAP;
        dv = o; o += LQP_WCAP;
        bw = o; o += LQP_WCAP;
        red = o; o += 24;
        cm = o; o += 10 * MAX_ACTIVE_CONTACTS * 7;  // friction/ZMP rows (10 kc x 6) and 1 / row norm
        ints = o; o += (LQP_WCAP + rtmax + 24 + 1) / 2;  // act[W], pos[rows], ired[24]
        total = o;
    }
};

// constraint row i -> stored row and sign.  Order: torque + (m), torque - (m), friction/ZMP (10 kc), accel + (m), accel - (m)
struct LqpRows {
    int m, kc;
    __device__ __forceinline__ int total() const { return 4 * m + 10 * kc; }
    __device__ __forceinline__ int level1_start() const { return 2 * m; }
    __device__ __forceinline__ void map(int i, int &ur, double &sgn) const {
        if (i < m) ur = i, sgn = 1.0;
        else if (i < 2 * m) ur = i - m, sgn = -1.0;
        else if (i < 3 * m + 10 * kc) ur = i - m, sgn = 1.0;
        else ur = i - 2 * m, sgn = -1.0;
    }
};

__device__ __forceinline__ double lq_team_sum(double v, double *red) {
    v = warp_sum(v);
    if (lane_id() == 0) red[t_id() >> 5] = v;
    tsync();
    double r = 0.0;
    for (int w = 0; w < (t_n() >> 5); w++) r += red[w];
    tsync();
    return r;
}

// sum_j G[ur, j] x[j]
__device__ __forceinline__ double lq_rowdot(const double *G, int ldg, int ur, int nu, const double *x) {
    const double *p = G + ur;
    double s0 = 0.0, s1 = 0.0;
    int j = 0;
    DWBC_UNROLL(1)
    for (; j + 1 < nu; j += 2) {
        s0 = fma(p[j * ldg], x[j], s0);
        s1 = fma(p[(j + 1) * ldg], x[j + 1], s1);
    }
    if (j < nu) s0 = fma(p[j * ldg], x[j], s0);
    return s0 + s1;
}

// R^T R x = b for the q x q upper-triangular R (warp 0 only; b is overwritten by x; `mid` (may be null) receives R^{-T} b).
__device__ __forceinline__ void lq_warp_solve_rtr(const double *R, int ldr, int q, double *b, double *mid) {
    const int lane = lane_id();
    for (int k = 0; k < q; k++) {
        __syncwarp();
        const double hk = b[k] / R[k + k * ldr];
        __syncwarp();
        if (lane == 0) b[k] = hk;
        for (int j = k + 1 + lane; j < q; j += 32) b[j] = fma(-R[k + j * ldr], hk, b[j]);
    }
    __syncwarp();
    if (mid)
        for (int j = lane; j < q; j += 32) mid[j] = b[j];
    for (int k = q - 1; k >= 0; k--) {
        __syncwarp();
        const double rk = b[k] / R[k + k * ldr];
        __syncwarp();
        if (lane == 0) b[k] = rk;
        for (int j = lane; j < k; j += 32) b[j] = fma(-R[j + k * ldr], rk, b[j]);
    }
    __syncwarp();
}

// Goldfarb-Idnani on  min 1/2 |xh|^2 + 1/2 |v|^2  s.t.  sgn_i G[ur_i,:] xh - [i in slack] v_i <= ub_i  for the rows [0, rows);
// slack rows are [slack_lo, rows) (empty when slack_lo >= rows).  On return xh holds the minimiser, act/lam (first q entries)
// the active rows and their multipliers (= v_i on slack rows).  Returns q, or -1 (infeasible / iteration or capacity limit).
// Whole team; synchronises on entry and exit.
__device__ __noinline__ int lqp_ldp_solve(double *sm, int n, int m, int tmax, int kc, int nu, int slack_lo, int *iters_out) {
    const LqpSmem L(n, m, tmax);
    const LqpRows RW{m, kc};
    const int rows = RW.total();
    const double *G = sm + L.G, *ub = sm + L.ub;
    const int ldg = L.ldg, ldr = L.ldr;
    double *R = sm + L.R, *xh = sm + L.xh, *zx = sm + L.zx, *rin = sm + L.rin, *tolv = sm + L.tol;
    double *lam = sm + L.lam, *gv = sm + L.gv, *hv = sm + L.hv, *rv = sm + L.rv, *dv = sm + L.dv, *bw = sm + L.bw, *red = sm + L.red;
    int *act = (int *)(sm + L.ints), *pos = act + LQP_WCAP, *ired = pos + L.rtmax;
    const int tid = t_id(), nt = t_n(), lane = lane_id(), warp = tid >> 5, nw = nt >> 5;
    tsync();
    // 1 / extended row norm; rows with a numerically zero normal never enter the active set (same rule as kernels_hqp.cuh)
    double smax = 0.0;
    for (int i = tid; i < rows; i += nt) {
        int ur;
        double sg;
        RW.map(i, ur, sg);
        double s = 0.0;
        for (int j = 0; j < nu; j++) {
            const double v = G[ur + j * ldg];
            s = fma(v, v, s);
        }
        rin[i] = s;
        smax = s > smax ? s : smax;
        pos[i] = -1;
    }
    smax = warp_max(smax);
    if (lane == 0) red[warp] = smax;
    for (int j = tid; j < nu; j += nt) xh[j] = 0.0;
    tsync();
    smax = red[0];
    for (int w = 1; w < nw; w++) smax = red[w] > smax ? red[w] : smax;
    tsync();
    double uabs = 0.0;  // magnitude of the right-hand sides (margin of the numerically zero rows, as in kernels_hqp.cuh)
    for (int i = tid; i < rows; i += nt) uabs = fmax(uabs, fabs(ub[i]));
    uabs = warp_max(uabs);
    if (lane == 0) red[warp] = uabs;
    tsync();
    uabs = red[0];
    for (int w = 1; w < nw; w++) uabs = red[w] > uabs ? red[w] : uabs;
    tsync();
    bool bad = false;
    for (int i = tid; i < rows; i += nt) {
        const double s = rin[i];
        const bool sl = i >= slack_lo;
        tolv[i] = 1e-10 * fmax(sqrt(s), fabs(ub[i]));
        if (sl)
            rin[i] = rsqrt(s + 1.0);
        else if (s > 1e-20 * smax && s > 0.0)
            rin[i] = rsqrt(s);
        else {
            rin[i] = 0.0;
            if (ub[i] < -1e-9 * (1.0 + uabs)) bad = true;
        }
    }
    if (lane == 0) ired[warp] = 0;
    tsync();
    if (bad) ired[warp] = 1;
    tsync();
    bool anybad = false;
    for (int w = 0; w < nw; w++) anybad = anybad || ired[w];
    tsync();
    if (anybad) return -1;
    int q = 0, iters = 0, status = 0;
    const int max_iter = 8 * (nu + 16);
    while (iters < max_iter) {
        // ---- most violated row (distance in the extended space) ----
        double best = 0.0;
        int bi = -1;
        for (int i = tid; i < rows; i += nt) {
            if (pos[i] >= 0 || rin[i] == 0.0) continue;
            int ur;
            double sg;
            RW.map(i, ur, sg);
            const double s1 = sg * lq_rowdot(G, ldg, ur, nu, xh) - ub[i];
            if (s1 > tolv[i] && s1 * rin[i] > best) best = s1 * rin[i], bi = i;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(DWBC_FULL, best, o);
            const int oi = __shfl_xor_sync(DWBC_FULL, bi, o);
            if (oi >= 0 && (bi < 0 || ob > best || (ob == best && oi < bi))) best = ob, bi = oi;
        }
        if (lane == 0) red[warp] = best, ired[warp] = bi;
        tsync();
        best = red[0], bi = ired[0];
        for (int w = 1; w < nw; w++) {
            const double ob = red[w];
            const int oi = ired[w];
            if (oi >= 0 && (bi < 0 || ob > best || (ob == best && oi < bi))) best = ob, bi = oi;
        }
        tsync();
        if (bi < 0) {
            status = 1;
            break;
        }
        const int p = bi;
        int pur;
        double psg;
        RW.map(p, pur, psg);
        const double pslack = p >= slack_lo ? 1.0 : 0.0;
        double sp = best / rin[p];
        double up = 0.0;
        bool added = false;
        while (!added && iters < max_iter) {
            iters++;
            // ---- gv = G_W g_p (four threads per active row) ----
            {
                const int part = tid & 3;
                for (int k0 = 0; k0 < q; k0 += nt >> 2) {
                    const int k = k0 + (tid >> 2);
                    double a0 = 0.0;
                    if (k < q) {
                        int ur;
                        double sg;
                        RW.map(act[k], ur, sg);
                        for (int j = part; j < nu; j += 4) a0 = fma(G[ur + j * ldg], G[pur + j * ldg], a0);
                        a0 *= sg * psg;
                    }
                    a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
                    a0 += __shfl_xor_sync(DWBC_FULL, a0, 2);
                    if (k < q && part == 0) gv[k] = a0, bw[k] = a0;
                }
            }
            tsync();
            if (warp == 0) {
                lq_warp_solve_rtr(R, ldr, q, bw, hv);
                for (int k = lane; k < q; k += 32) rv[k] = bw[k];
            }
            tsync();
            // ---- z = g_p - G_W^T r, then one refinement step: delta = N_W^T z, r += (R^T R)^{-1} delta ----
            for (int pass = 0; pass < 2; pass++) {
                for (int j = tid; j < nu; j += nt) {
                    double s = pass == 0 ? psg * G[pur + j * ldg] : zx[j];
                    const double *rr = pass == 0 ? rv : dv;
                    for (int k = 0; k < q; k++) {
                        int ur;
                        double sg;
                        RW.map(act[k], ur, sg);
                        s = fma(-sg * rr[k], G[ur + j * ldg], s);
                    }
                    zx[j] = s;
                }
                tsync();
                if (pass == 1) break;
                {
                    const int part = tid & 3;
                    for (int k0 = 0; k0 < q; k0 += nt >> 2) {
                        const int k = k0 + (tid >> 2);
                        double a0 = 0.0;
                        if (k < q) {
                            int ur;
                            double sg;
                            RW.map(act[k], ur, sg);
                            for (int j = part; j < nu; j += 4) a0 = fma(G[ur + j * ldg], zx[j], a0);
                            a0 *= sg;
                        }
                        a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
                        a0 += __shfl_xor_sync(DWBC_FULL, a0, 2);
                        if (k < q && part == 0) bw[k] = a0 - (act[k] >= slack_lo ? rv[k] : 0.0);
                    }
                }
                tsync();
                if (warp == 0) {
                    lq_warp_solve_rtr(R, ldr, q, bw, nullptr);
                    for (int k = lane; k < q; k += 32) dv[k] = bw[k], rv[k] += bw[k];
                }
                tsync();
            }
            // ---- step lengths ----
            double zz = 0.0, nn = 0.0;
            for (int j = lane; j < nu; j += 32) {  // identical in every warp
                zz = fma(zx[j], zx[j], zz);
                const double g = G[pur + j * ldg];
                nn = fma(g, g, nn);
            }
            for (int k = lane; k < q; k += 32)
                if (act[k] >= slack_lo) zz = fma(rv[k], rv[k], zz);
            zz = warp_sum(zz) + pslack;
            nn = warp_sum(nn) + pslack;
            double t1 = 1e300;
            int l1 = -1;
            for (int k = lane; k < q; k += 32)
                if (rv[k] > 0.0) {
                    const double tk = lam[k] / rv[k];
                    if (tk < t1) t1 = tk, l1 = k;
                }
            for (int o = 16; o > 0; o >>= 1) {
                const double ot = __shfl_xor_sync(DWBC_FULL, t1, o);
                const int ol = __shfl_xor_sync(DWBC_FULL, l1, o);
                if (ol >= 0 && (l1 < 0 || ot < t1 || (ot == t1 && ol < l1))) t1 = ot, l1 = ol;
            }
            const bool dependent = !(zz > 1e-16 * nn);
            if (dependent && l1 < 0) {
                if (sp <= 100.0 * tolv[p]) {  // violated by rounding only: leave it
                    added = true;
                    tsync();
                    if (tid == 0) rin[p] = 0.0;  // do not select it again
                    tsync();
                    continue;
                }
                status = -1;
                break;
            }
            const double t2 = dependent ? 1e300 : sp / zz;
            const double t = t1 < t2 ? t1 : t2;
            tsync();
            for (int k = tid; k < q; k += nt) lam[k] = fma(-t, rv[k], lam[k]);
            up += t;
            if (!dependent)
                for (int j = tid; j < nu; j += nt) xh[j] = fma(-t, zx[j], xh[j]);
            if (!dependent && t2 <= t1) {
                if (q >= LQP_WCAP) {
                    status = -1;
                    break;
                }
                for (int k = tid; k < q; k += nt) R[k + q * ldr] = hv[k];
                if (tid == 0) {
                    // diagonal of the new column: |z| computed explicitly (more accurate than sqrt(|n|^2 - |h|^2))
                    R[q + q * ldr] = sqrt(zz);
                    act[q] = p;
                    lam[q] = up;
                    pos[p] = q;
                }
                q++;
                added = true;
                tsync();
            } else {
                // ---- the blocking row at position l1 leaves the active set ----
                const int l = l1;
                tsync();
                if (warp == 0) {
                    for (int k = l; k + 1 < q; k++) {  // shift columns left
                        for (int i = lane; i <= k + 1; i += 32) R[i + k * ldr] = R[i + (k + 1) * ldr];
                        __syncwarp();
                    }
                    for (int k = l; k + 1 < q; k++) {  // Givens on rows k, k+1 (columns k .. q-2)
                        const double a = R[k + k * ldr], b2 = R[k + 1 + k * ldr];
                        const double h = sqrt(a * a + b2 * b2);
                        const double cs = h > 0.0 ? a / h : 1.0, sn = h > 0.0 ? b2 / h : 0.0;
                        __syncwarp();
                        for (int j = k + lane; j < q - 1; j += 32) {
                            const double r0 = R[k + j * ldr], r1 = R[k + 1 + j * ldr];
                            R[k + j * ldr] = cs * r0 + sn * r1;
                            R[k + 1 + j * ldr] = -sn * r0 + cs * r1;
                        }
                        __syncwarp();
                    }
                    if (lane == 0) {
                        pos[act[l]] = -1;
                        for (int k = l; k + 1 < q; k++) {
                            act[k] = act[k + 1];
                            lam[k] = lam[k + 1];
                            pos[act[k]] = k;
                        }
                    }
                }
                q--;
                tsync();
                // slack of row p at the new point
                double part = 0.0;
                for (int j = lane; j < nu; j += 32) part = fma(G[pur + j * ldg], xh[j], part);
                sp = psg * warp_sum(part) - ub[p];
                if (sp < 0.0) sp = 0.0;
            }
        }
        if (status < 0) break;
    }
    tsync();
    if (iters_out) *iters_out += iters;
    // clean multipliers
    for (int k = tid; k < q; k += nt)
        if (lam[k] < 0.0) lam[k] = 0.0;
    tsync();
    return status == 1 ? q : -1;
}

__global__ void __launch_bounds__(TEAM, 2) lqp_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, LqpOut out, int tmax,
                                                      LqpParams prm) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n(), lane = lane_id(), warp = tid >> 5, nw = nt >> 5;
    const long long s = blockIdx.x;
    const int n = w.n, m = w.m, nb = w.nb;
    const LqpSmem L(n, m, tmax);
    double *sm = (double *)smem;
    double *Z = sm + L.Z, *G = sm + L.G, *H = sm + L.H, *E = sm + L.E, *R = sm + L.R;
    double *y = sm + L.y, *u = sm + L.u, *xh = sm + L.xh, *gh = sm + L.gh, *gg = sm + L.gg, *ee = sm + L.ee, *htau = sm + L.hv_tau, *ta = sm + L.ta;
    double *ub = sm + L.ub, *vfro = sm + L.vfro, *lam = sm + L.lam, *red = sm + L.red, *cm = sm + L.cm;
    int *act = (int *)(sm + L.ints);
    const int ldz = L.ldz, ldg = L.ldg, ldh = L.ldh, lde = L.lde, ldt = L.ldt;
    const int nvs = n + CMAX;  // stride of the per-level answers
    double *yout = out.y + s * (long long)LQP_MAX_LEVELS * nvs;
    const int st0 = w.status[s];
    const int c = (st0 & 1) ? w.cdof[s] : 0;
    const int kc = c / 6;
    const int nlev = 2 + cfg->nlevels;
    for (int e = tid; e < LQP_MAX_LEVELS * nvs; e += nt) yout[e] = 0.0;
    for (int i = tid; i < L.urmax; i += nt) out.v1[s * L.urmax + i] = 0.0;
    for (int i = tid; i < m; i += nt) out.tau[s * m + i] = 0.0;
    if (c == 0) {  // uniform over the CTA
        if (tid == 0) out.status[s] = 0, out.iters[s] = 0;
        return;
    }
    const int nvar = n + c;
    const LqpRows RW{m, kc};
    const int rows = RW.total();
    const double *A = w.A + s * (long long)n * n, *Ainv = w.Ainv + s * (long long)n * n, *Bnl = w.Bnl + s * n;
    const double *Jc = w.Jc + s * (long long)CMAX * n;
    const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
    int iters = 0;
    // ---- y0 = [-A^{-1} B; 0], |A|_F, friction/ZMP rows C = -A_const A_rot (reference src/dwbc.cpp:480-513) ----
    tgemv<false>(n, n, Ainv, n, Bnl, [&](int i, double v) { y[i] = -v; });
    for (int i = tid; i < c; i += nt) y[n + i] = 0.0;
    for (int i = tid; i < L.rtmax; i += nt) vfro[i] = 0.0;
    double an = 0.0;
    {
        const double *An = prm.anorm_src + s * (long long)prm.anorm_n * prm.anorm_n;
        for (int e = tid; e < prm.anorm_n * prm.anorm_n; e += nt) an = fma(An[e], An[e], an);
    }
    an = sqrt(lq_team_sum(an, red));
    const double hs = 5.0 / an;  // reference src/dwbc.cpp:4341
    for (int e = tid; e < kc * 6; e += nt) {
        const int k = e / 6, d = e - 6 * k;
        const ContactCfg &cc = cfg->contacts[w.cidx[s * MAX_ACTIVE_CONTACTS + k]];
        double o[10];
        contact_rows(cc, pose + cc.link * POSE_STRIDE, v3(d == 0, d == 1, d == 2), v3(d == 3, d == 4, d == 5), o);
        for (int a = 0; a < 10; a++) cm[(10 * k + a) * 7 + d] = -o[a];
    }
    for (int e = tid; e < nvar * nvar; e += nt) {
        const int j = e / nvar, i = e - j * nvar;
        Z[i + j * ldz] = i == j ? 1.0 : 0.0;
    }
    tsync();
    for (int k = tid; k < 10 * kc; k += nt) {
        double s2 = 0.0;
        for (int d = 0; d < 6; d++) s2 = fma(cm[k * 7 + d], cm[k * 7 + d], s2);
        cm[k * 7 + 6] = s2 > 0.0 ? rsqrt(s2) : 0.0;
    }
    for (int e = tid; e < nvs; e += nt) yout[e] = e < nvar ? y[e] : 0.0;
    int win = 0;  // the level's null-space basis is Z[:, win:nvar)
    int solved = 1;
    for (int lev = 0; lev < nlev; lev++) {
        const int nu = nvar - win;
        double *Zw = Z + win * ldz;
        int neq;
        tsync();
        // ---- E = rows-normalised B_lev Z (neq x nu), ee = (B_lev y + b_lev) / |row| ----
        if (lev == 0) {
            neq = 6;
            for (int r = tid; r < 6; r += nt) {
                double s2 = 0.0;
                for (int j = 0; j < n; j++) s2 = fma(A[r + j * n], A[r + j * n], s2);
                for (int j = 0; j < c; j++) s2 = fma(Jc[j + r * CMAX], Jc[j + r * CMAX], s2);
                ee[r] = s2 > 0.0 ? rsqrt(s2) : 0.0;
            }
            tsync();
            for (int e = tid; e < 6 * nvar; e += nt) {
                const int j = e / 6, r = e - 6 * j;
                E[r + j * lde] = (j < n ? A[r + j * n] : Jc[(j - n) + r * CMAX]) * ee[r];
            }
        } else {
            const double *Bm;
            int ldb;
            if (lev == 1) {
                neq = c;
                Bm = Jc;
                ldb = CMAX;
            } else {
                neq = cfg->tasks[lev - 2].dof;
                build_task_jacobian(md, cfg, w, s, lev - 2, R, lde);
                Bm = R;
                ldb = lde;
                tsync();
            }
            for (int r = tid; r < neq; r += nt) {
                double s2 = 0.0, d = 0.0;
                for (int j = 0; j < n; j++) {
                    const double v = Bm[r + j * ldb];
                    s2 = fma(v, v, s2);
                    d = fma(v, y[j], d);
                }
                const double rn = s2 > 0.0 ? rsqrt(s2) : 0.0;
                htau[r] = rn;
                const double bb = lev == 1 ? 0.0 : -w.fstar[lev - 2][s * neq + r];
                ee[r] = (d + bb) * rn;
            }
            tsync();
            tgemm_mma<2, false, false>(neq, nu, n, Bm, ldb, Zw, ldz, [&](int i, int j, double v) { E[i + j * lde] = v * htau[i]; });
        }
        tsync();
        if (lev > 0) {
            // ---- T = A Z_a (n x nu, in the R arena), ta = A y_a ----
            double *T = R;
            tgemm_mma<2, false, false>(n, nu, n, A, n, Zw, ldz, [&](int i, int j, double v) { T[i + j * ldt] = v; });
            tgemv<false>(n, n, A, n, y, [&](int i, double v) { ta[i] = v; });
            tsync();
            // ---- Hu = E^T E + hs Z_a^T T + sigma I,  g = E^T e + hs Z_a^T ta ----
            tgemm_mma<2, true, false>(nu, nu, neq, E, lde, E, lde, [&](int i, int j, double v) { H[i + j * ldh] = v; });
            tsync();
            tgemm_mma<2, true, false>(nu, nu, n, Zw, ldz, T, ldt, [&](int i, int j, double v) { H[i + j * ldh] += hs * v + (i == j ? LQP_SIGMA : 0.0); });
            tgemv<true>(nu, neq, E, lde, ee, [&](int i, double v) { gg[i] = v; });
            tsync();
            tgemv<true>(nu, n, Zw, ldz, ta, [&](int i, double v) { gg[i] += hs * v; });
            // ---- stored constraint rows G = A_ineq Z (normalised) and right-hand sides ub = -(A y + a) + v_frozen ----
            for (int r = tid; r < m; r += nt) {
                double s2 = 0.0, tq = ta[6 + r] + Bnl[6 + r];
                for (int j = 0; j < n; j++) s2 = fma(A[6 + r + j * n], A[6 + r + j * n], s2);
                for (int j = 0; j < c; j++) {
                    const double v = Jc[j + (6 + r) * CMAX];
                    s2 = fma(v, v, s2);
                    tq = fma(v, y[n + j], tq);
                }
                const double rn = s2 > 0.0 ? rsqrt(s2) : 0.0;
                u[r] = rn;  // scratch: 1 / row norm of the torque rows
                ub[r] = (prm.tlim[r] - tq) * rn;
                ub[m + r] = (prm.tlim[r] + tq) * rn;
                const int ia = 3 * m + 10 * kc;
                ub[2 * m + 10 * kc + r] = LQP_ACC_LIMIT - y[6 + r] + vfro[2 * m + 10 * kc + r];
                ub[ia + r] = LQP_ACC_LIMIT + y[6 + r] + vfro[ia + r];
            }
            for (int k = tid; k < 10 * kc; k += nt) {
                const int kk = k / 10;
                double d = 0.0;
                for (int a = 0; a < 6; a++) d = fma(cm[k * 7 + a], y[n + 6 * kk + a], d);
                ub[2 * m + k] = -d * cm[k * 7 + 6] + vfro[2 * m + k];
            }
            tsync();
            // torque rows: (T[6+r,:] + Jc^T[6+r,:] Z_f) / |row|;  friction / ZMP rows: C_k Z_f[6k:6k+6,:] / |row|;  accel rows: Z[6+r,:]
            tgemm_mma<2, true, false>(m, nu, c, Jc + 6 * CMAX, CMAX, Zw + n, ldz, [&](int i, int j, double v) { G[i + j * ldg] = (v + T[6 + i + j * ldt]) * u[i]; });
            for (int k = 0; k < kc; k++)
                tgemm_mma<2, true, false>(10, nu, 6, cm + 70 * k, 7, Zw + n + 6 * k, ldz,
                                          [&](int i, int j, double v) { G[m + 10 * k + i + j * ldg] = v * cm[(10 * k + i) * 7 + 6]; }, k + 1);
            for (int e = tid; e < m * nu; e += nt) {
                const int j = e / m, r = e - j * m;
                G[m + 10 * kc + r + j * ldg] = Zw[6 + r + j * ldz];
            }
            tsync();  // T (in the R arena) is dead from here on
            // ---- Lc Lc^T = Hu;  gh = Lc^{-1} g;  G <- G Lc^{-T};  ub += G gh ----
            double *rdiag = u;  // 1 / Lc(i,i)   (the torque-row norms are no longer needed after the barrier inside tchol_lower)
            const double piv = tchol_lower<LQP_CHOL_OWN>(nu, H, ldh, rdiag);
            if (!(piv > 0.0)) {
                solved = 0;
                break;
            }
            // Xi = Lc^{-1} (R arena, free until the solve);  gh = Xi g;  G <- G Xi^T in place on the tensor cores: an item is an
            // 8-row band times ALL column tiles, so it only reads the rows it later overwrites
            double *Xi = R;
            const int ldx = L.ldh;
            ttri_inverse(nu, H, ldh, rdiag, Xi, ldx);
            tsync();
            tgemv<false>(nu, nu, Xi, ldx, gg, [&](int i, double v) { gh[i] = v; });
            tgemm_mma<LQP_NU_TILES, false, true>(2 * m + 10 * kc, nu, nu, G, ldg, Xi, ldx, [&](int i, int j, double v) { G[i + j * ldg] = v; });
            tsync();
            for (int i = tid; i < rows; i += nt) {
                int ur;
                double sg;
                RW.map(i, ur, sg);
                ub[i] += sg * lq_rowdot(G, ldg, ur, nu, gh);
            }
            tsync();
            const int q = lqp_ldp_solve(sm, n, m, tmax, kc, nu, lev == 1 ? RW.level1_start() : rows, &iters);
            if (q < 0) {
                solved = 0;
                break;
            }
            // ---- u = Lc^{-T} (xh - gh);  y += Z u;  freeze the level's slack ----
            if (warp == 0) {
                for (int j = lane; j < nu; j += 32) gg[j] = xh[j] - gh[j];
                for (int k = nu - 1; k >= 0; k--) {
                    __syncwarp();
                    const double rk = gg[k] * rdiag[k];
                    __syncwarp();
                    if (lane == 0) gg[k] = rk;
                    for (int j = lane; j < k; j += 32) gg[j] = fma(-H[k + j * ldh], rk, gg[j]);
                }
                __syncwarp();
            }
            if (lev == 1)
                for (int k = tid; k < q; k += nt)
                    if (act[k] >= RW.level1_start()) {
                        vfro[act[k]] = lam[k];
                        out.v1[s * L.urmax + act[k] - RW.level1_start()] = lam[k];
                    }
            tsync();
            tgemv<false>(nvar, nu, Zw, ldz, gg, [&](int i, double v) {
                const double yn = y[i] + v;  // y[i] is only touched by the thread that owns row i
                y[i] = yn;
                yout[lev * nvs + i] = yn;
            });
            solved = lev + 1;
        }
        if (lev + 1 == nlev) break;
        // ---- Z[:, win:] <- Z[:, win:] H_0 ... H_{neq-1}  (Householder reflections of (B Z)^T); the next window drops neq columns ----
        tsync();
        bool rank_ok = true;
        for (int k = 0; k < neq; k++) {
            // reflection from row k of E, entries j >= k (every warp computes the same scalars)
            double s2 = 0.0;
            for (int j = k + lane; j < nu; j += 32) s2 = fma(E[k + j * lde], E[k + j * lde], s2);
            s2 = warp_sum(s2);
            const double x0 = E[k + k * lde];
            if (!(s2 > 1e-20)) {
                rank_ok = false;
                break;
            }
            const double alpha = x0 > 0.0 ? -sqrt(s2) : sqrt(s2);
            const double tau = 1.0 / (s2 - alpha * x0);  // 2 / |v|^2, v = x - alpha e_k
            tsync();
            if (tid == 0) E[k + k * lde] = x0 - alpha, htau[k] = tau;
            tsync();
            for (int r = k + 1 + warp; r < neq; r += nw) {
                double d = 0.0;
                for (int j = k + lane; j < nu; j += 32) d = fma(E[r + j * lde], E[k + j * lde], d);
                d = warp_sum(d) * tau;
                for (int j = k + lane; j < nu; j += 32) E[r + j * lde] = fma(-d, E[k + j * lde], E[r + j * lde]);
            }
            tsync();
        }
        if (!rank_ok) {
            solved = lev > 0 ? solved : 0;
            break;
        }
        for (int i = tid; i < nvar; i += nt) {
            double *zr = Zw + i;
            for (int k = 0; k < neq; k++) {
                double d = 0.0;
                for (int j = k; j < nu; j++) d = fma(zr[j * ldz], E[k + j * lde], d);
                d *= htau[k];
                for (int j = k; j < nu; j++) zr[j * ldz] = fma(-d, E[k + j * lde], zr[j * ldz]);
            }
        }
        win += neq;
    }
    tsync();
    // failed levels keep the previous level's answer (oracle / reference: the sequence stops at the first failure)
    for (int lev = solved; lev < nlev; lev++)
        for (int i = tid; i < nvar; i += nt) yout[lev * nvs + i] = y[i];
    // ---- tau = A[6:,:] qdd + Jc^T[6:,:] Fc + B[6:]  (torque rows of level 0, reference src/dwbc.cpp:4355-4370) ----
    tgemv<false>(m, n, A + 6, n, y, [&](int i, double v) {
        double t = v + Bnl[6 + i];
        for (int j = 0; j < c; j++) t = fma(Jc[j + (6 + i) * CMAX], y[n + j], t);
        out.tau[s * m + i] = t;
    });
    if (tid == 0) out.status[s] = solved == nlev ? nlev : (solved > 0 ? -solved : 0), out.iters[s] = iters;
}

// ---------------------------------------------------------------------------------------------------------------------------
// ConfigureLQP_R_NC + CalcControlTorqueLQP_R_NC (reference src/dwbc.cpp:4634-4760, 4479-4502): the non-contact chain's own
// two-level hierarchy over its nc joint accelerations (no contact force), fed by the answer of the reduced LQP:
//   level 0 (HQP::solvefirst, src/dwbc_hqp.cpp:222-289)   eq  J_I_nc qdd = fstar_gnc (last 6 entries of the reduced qdd)
//                                                         ineq |A[nc,nc] qdd + G[nc]| <= 200, own slack; cost 5 A[nc,nc] / |A[nc,nc]|_F
//   level 1 (solveSequentialSingle)                       eq  J_task(ts_[1])[:, nc] qdd = Ja (f*_1 - fstar_base), Ja = [1 skew(p_link - p_pelvis); 0 1]
//                                                         ineq |qdd| <= 5, own slack; the torque rows of level 0 frozen; same cost
// ts_[1] is hard-coded by the reference (:4738-4745) and the rows are NOT normalised here (:4677, 4755).  A_NC[nc,nc] = A[nc,nc]
// (the subtree of a non-contact joint is entirely non-contact, kernels_reduced.cuh).  Same exact level solve as lqp_kernel (the
// stored-row layout of LqpRows with kc = 0, "m" = nc: torque rows, then acceleration rows), on an arena LqpSmem(nc, nc, 6).
// Outputs: ync [B][2][nc] (answer after each level), tau_nc [B][nc] = A[nc,nc] qdd + G[nc], and the stitched joint torque of the
// only caller (reference tests/sp_test/jacc_compare.cpp:543-569): tau [B][m] = [tau_reduced_lqp[0:co_dof]; tau_nc].
struct LqpNcArgs {
    const double *yR;      // [B][stride_yR] final answer of the reduced LQP (qdd_R first)
    int stride_yR, rn;
    const double *tauR;    // [B][rm] torque of the reduced LQP
    int rm, co_dof, vc_dof, nc;
    const int *statusR;    // [B] status of the reduced LQP (> 0: every level solved)
    int levelsR;
    double *ync, *tau_nc, *tau;  // outputs
    int *status, *iters;         // [B] 2 = both levels solved, -k = level k failed, 0 = no valid input
};

__global__ void __launch_bounds__(TEAM, 3) lqp_nc_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, const double *JI_all,
                                                         LqpNcArgs a) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n(), lane = lane_id(), warp = tid >> 5, nw = nt >> 5;
    const long long s = blockIdx.x;
    const int n = w.n, m = w.m, nb = w.nb, nc = a.nc, vc = a.vc_dof;
    const LqpSmem L(nc, nc, 6);
    double *sm = (double *)smem;
    double *Z = sm + L.Z, *G = sm + L.G, *H = sm + L.H, *E = sm + L.E, *R = sm + L.R;
    double *y = sm + L.y, *u = sm + L.u, *xh = sm + L.xh, *gh = sm + L.gh, *gg = sm + L.gg, *ee = sm + L.ee, *htau = sm + L.hv_tau, *ta = sm + L.ta;
    double *ub = sm + L.ub, *vfro = sm + L.vfro, *lam = sm + L.lam, *red = sm + L.red, *cm = sm + L.cm;
    int *act = (int *)(sm + L.ints);
    const int ldz = L.ldz, ldg = L.ldg, ldh = L.ldh, lde = L.lde, ldt = L.ldt;
    double *yout = a.ync + s * 2 * (long long)nc;
    for (int e = tid; e < 2 * nc; e += nt) yout[e] = 0.0;
    for (int i = tid; i < nc; i += nt) a.tau_nc[s * nc + i] = 0.0;
    for (int i = tid; i < m; i += nt) a.tau[s * m + i] = 0.0;
    if (!(a.statusR[s] == a.levelsR)) {  // uniform over the CTA: the reduced LQP did not produce an answer
        if (tid == 0) a.status[s] = 0, a.iters[s] = 0;
        return;
    }
    const LqpRows RW{nc, 0};
    const int rows = RW.total();  // 4 nc: torque +, torque -, accel +, accel -
    const double *A = w.A + s * (long long)n * n + vc + (long long)vc * n;  // A[nc, nc], leading dimension n
    const double *Gnc = w.G + s * n + vc;
    const double *JI = JI_all + s * 6 * (MAXB - 1);  // 6 x nc, ld 6 (records of NCMAX = MAXB - 1 columns, kernels_reduced.cuh)
    const double *qR = a.yR + s * (long long)a.stride_yR;
    const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
    int iters = 0;
    // cost scale 5 / |A[nc,nc]|_F
    double an = 0.0;
    for (int e = tid; e < nc * nc; e += nt) {
        const int j = e / nc, i = e - j * nc;
        an = fma(A[i + (long long)j * n], A[i + (long long)j * n], an);
    }
    an = sqrt(lq_team_sum(an, red));
    const double hs = 5.0 / an;
    for (int e = tid; e < nc * nc; e += nt) {
        const int j = e / nc, i = e - j * nc;
        Z[i + j * ldz] = i == j ? 1.0 : 0.0;
    }
    for (int i = tid; i < nc; i += nt) y[i] = 0.0;
    for (int i = tid; i < L.rtmax; i += nt) vfro[i] = 0.0;
    // fstar_local = Ja (f*_1 - fstar_base)   (cm is free scratch here: no friction rows)
    if (tid < 6) {
        const TaskCfg &tc = cfg->tasks[1];
        const int link = cfg->tlinks[tc.link_start].link;
        const double *pl = pose + link * POSE_STRIDE + 9, *p0 = pose + 9;
        const double d[3] = {pl[0] - p0[0], pl[1] - p0[1], pl[2] - p0[2]};
        double df[6];
        for (int k = 0; k < 6; k++) df[k] = w.fstar[1][s * 6 + k] - qR[k];
        double v = df[tid];
        if (tid == 0) v += -d[2] * df[4] + d[1] * df[5];
        if (tid == 1) v += d[2] * df[3] - d[0] * df[5];
        if (tid == 2) v += -d[1] * df[3] + d[0] * df[4];
        cm[tid] = v;
    }
    int win = 0, solved = 0;
    for (int lev = 0; lev < 2; lev++) {
        const int nu = nc - win, neq = 6;
        double *Zw = Z + win * ldz;
        tsync();
        // ---- E = B_lev Z (6 x nu), ee = B_lev y + b_lev ----
        const double *Bm;
        int ldb;
        if (lev == 0) {
            Bm = JI;
            ldb = 6;
        } else {
            build_task_jacobian(md, cfg, w, s, 1, R, lde);  // 6 x n; its last nc columns are J_task_.rightCols(nc_dof)
            Bm = R + vc * lde;
            ldb = lde;
            tsync();
        }
        for (int r = tid; r < neq; r += nt) {
            double d = 0.0;
            for (int j = 0; j < nc; j++) d = fma(Bm[r + j * ldb], y[j], d);
            ee[r] = d + (lev == 0 ? -qR[a.rn - 6 + r] : -cm[r]);
            htau[r] = 1.0;
        }
        tgemm<1, 2, false, false>(neq, nu, nc, Bm, ldb, Zw, ldz, [&](int i, int j, double v) { E[i + j * lde] = v; });
        tsync();
        // ---- T = A[nc,nc] Z_a (nc x nu), ta = A[nc,nc] y ----
        double *T = R;
        tgemm_mma<2, false, false>(nc, nu, nc, A, n, Zw, ldz, [&](int i, int j, double v) { T[i + j * ldt] = v; });
        tgemv<false>(nc, nc, A, n, y, [&](int i, double v) { ta[i] = v; });
        tsync();
        // ---- Hu = E^T E + hs Z_a^T T + sigma I,  g = E^T e + hs Z_a^T ta ----
        tgemm_mma<2, true, false>(nu, nu, neq, E, lde, E, lde, [&](int i, int j, double v) { H[i + j * ldh] = v; });
        tsync();
        tgemm_mma<2, true, false>(nu, nu, nc, Zw, ldz, T, ldt, [&](int i, int j, double v) { H[i + j * ldh] += hs * v + (i == j ? LQP_SIGMA : 0.0); });
        tgemv<true>(nu, neq, E, lde, ee, [&](int i, double v) { gg[i] = v; });
        tsync();
        tgemv<true>(nu, nc, Zw, ldz, ta, [&](int i, double v) { gg[i] += hs * v; });
        // ---- stored rows: torque rows T (raw), acceleration rows Z (absent at level 0: zero rows with a positive right-hand side) ----
        for (int r = tid; r < nc; r += nt) {
            const double tq = ta[r] + Gnc[r];
            ub[r] = (LQP_TORQUE_LIMIT - tq) + vfro[r];
            ub[nc + r] = (LQP_TORQUE_LIMIT + tq) + vfro[nc + r];
            ub[2 * nc + r] = lev == 0 ? 1.0 : LQP_ACC_LIMIT - y[r];
            ub[3 * nc + r] = lev == 0 ? 1.0 : LQP_ACC_LIMIT + y[r];
        }
        for (int e = tid; e < nc * nu; e += nt) {
            const int j = e / nc, r = e - j * nc;
            G[r + j * ldg] = T[r + j * ldt];
            G[nc + r + j * ldg] = lev == 0 ? 0.0 : Zw[r + j * ldz];
        }
        tsync();  // T (in the R arena) is dead from here on
        double *rdiag = u;
        const double piv = tchol_lower<LQP_CHOL_OWN>(nu, H, ldh, rdiag);
        if (!(piv > 0.0)) break;
        double *Xi = R;
        const int ldx = L.ldh;
        ttri_inverse(nu, H, ldh, rdiag, Xi, ldx);
        tsync();
        tgemv<false>(nu, nu, Xi, ldx, gg, [&](int i, double v) { gh[i] = v; });
        tgemm_mma<LQP_NU_TILES, false, true>(2 * nc, nu, nu, G, ldg, Xi, ldx, [&](int i, int j, double v) { G[i + j * ldg] = v; });
        tsync();
        for (int i = tid; i < rows; i += nt) {
            int ur;
            double sg;
            RW.map(i, ur, sg);
            ub[i] += sg * lq_rowdot(G, ldg, ur, nu, gh);
        }
        tsync();
        const int q = lqp_ldp_solve(sm, nc, nc, 6, 0, nu, lev == 0 ? 0 : RW.level1_start(), &iters);
        if (q < 0) break;
        if (warp == 0) {
            for (int j = lane; j < nu; j += 32) gg[j] = xh[j] - gh[j];
            for (int k = nu - 1; k >= 0; k--) {
                __syncwarp();
                const double rk = gg[k] * rdiag[k];
                __syncwarp();
                if (lane == 0) gg[k] = rk;
                for (int j = lane; j < k; j += 32) gg[j] = fma(-H[k + j * ldh], rk, gg[j]);
            }
            __syncwarp();
        }
        if (lev == 0)  // freeze the slack of the torque rows (the acceleration rows are not part of level 0)
            for (int k = tid; k < q; k += nt)
                if (act[k] < 2 * nc) vfro[act[k]] = lam[k];
        tsync();
        tgemv<false>(nc, nu, Zw, ldz, gg, [&](int i, double v) {
            const double yn = y[i] + v;
            y[i] = yn;
            yout[lev * nc + i] = yn;
        });
        solved = lev + 1;
        if (lev == 1) break;
        // ---- Z <- Z H_0 ... H_5 (Householder reflections of (B Z)^T) ----
        tsync();
        bool rank_ok = true;
        for (int k = 0; k < neq; k++) {
            double s2 = 0.0;
            for (int j = k + lane; j < nu; j += 32) s2 = fma(E[k + j * lde], E[k + j * lde], s2);
            s2 = warp_sum(s2);
            const double x0 = E[k + k * lde];
            if (!(s2 > 1e-20)) {
                rank_ok = false;
                break;
            }
            const double alpha = x0 > 0.0 ? -sqrt(s2) : sqrt(s2);
            const double tau = 1.0 / (s2 - alpha * x0);
            tsync();
            if (tid == 0) E[k + k * lde] = x0 - alpha, htau[k] = tau;
            tsync();
            for (int r = k + 1 + warp; r < neq; r += nw) {
                double d = 0.0;
                for (int j = k + lane; j < nu; j += 32) d = fma(E[r + j * lde], E[k + j * lde], d);
                d = warp_sum(d) * tau;
                for (int j = k + lane; j < nu; j += 32) E[r + j * lde] = fma(-d, E[k + j * lde], E[r + j * lde]);
            }
            tsync();
        }
        if (!rank_ok) break;
        for (int i = tid; i < nc; i += nt) {
            double *zr = Zw + i;
            for (int k = 0; k < neq; k++) {
                double d = 0.0;
                for (int j = k; j < nu; j++) d = fma(zr[j * ldz], E[k + j * lde], d);
                d *= htau[k];
                for (int j = k; j < nu; j++) zr[j * ldz] = fma(-d, E[k + j * lde], zr[j * ldz]);
            }
        }
        win += neq;
    }
    tsync();
    if (solved == 1)
        for (int i = tid; i < nc; i += nt) yout[nc + i] = y[i];  // a failed level keeps the previous answer
    // tau_nc = A[nc,nc] qdd + G[nc]; stitched joint torque
    tgemv<false>(nc, nc, A, n, y, [&](int i, double v) {
        const double t = v + Gnc[i];
        a.tau_nc[s * nc + i] = t;
        a.tau[s * m + (m - nc) + i] = t;
    });
    for (int i = tid; i < a.co_dof; i += nt) a.tau[s * m + i] = a.tauR[s * a.rm + i];
    if (tid == 0) a.status[s] = solved == 2 ? 2 : -solved, a.iters[s] = iters;
}

}  // namespace dwbc
