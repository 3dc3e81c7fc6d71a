// Stage a5: task-space quantities of every hierarchy level, and the closed-form task torque of the
// hqp=false branch.  One warp per robot state.
//
// Reference: RobotData::UpdateTaskSpace / CalcTaskSpace (src/dwbc.cpp:685-816), TaskSpace::CalcJKT /
// CalcNullMatrix (src/task.cpp:90-93, 202-211), CalculateJKT / CalculateTaskNullSpace
// (src/wbd.cpp:207-213, 257-261), CalcTaskControlTorque hqp=false branch (src/dwbc.cpp:856-871).
//
// Restructuring (same results up to rounding):
//   T = J (A^{-1}N_c) is formed once per level and reused for Lambda_t^{-1} = T J^T, Q = Lambda_t T[:,6:]
//   and the null-space update; Lambda_t and (Q W^+ Q^T) are SPD and inverted by Cholesky;
//   Ntorque_l = Null_{l-1} J_kt Lambda_t is what both the torque sum and the level QP need, and
//   Null_l = Null_{l-1} - Ntorque_l T[:,6:]   replaces the reference's 33^3 product per level.
#pragma once
#include "batch.h"
#include "kernels_contact.cuh"
#include "kernels_rbd.cuh"
#include "team_linalg.cuh"
#include "warp_linalg.cuh"

namespace dwbc {

struct TaskSmem {
    int ldt, ldm, X1, X2, T, Nul, S1, S2, S3, vec, total;
    // tmax: largest task dof of the configured hierarchy (the arena is sized for it at launch time)
    __host__ __device__ TaskSmem(int n, int m, int tmax) {
        ldt = tmax | 1;
        ldm = m | 1;
        int x1 = ldt * n > ldm * tmax ? ldt * n : ldm * tmax;
        int o = 0;
        X1 = o; o += x1;            // J_task -> Q -> J_kt -> Ntorque
        X2 = o; o += ldm * tmax;    // W^+ Q^T -> J_kt Lambda_t
        T = o; o += ldt * n;        // J (A^{-1} N_c)
        Nul = o; o += ldm * m;      // Null_{l-1}
        S1 = o; o += ldt * tmax;    // Lambda_t
        S2 = o; o += ldt * tmax;    // (Q W^+ Q^T)^{-1}
        S3 = -1;                    // scratch of the SPD inverses: X2 for Lambda_t (W^+ Q^T not formed yet), X1 for (Q W^+ Q^T)^{-1} (Q dead)
        vec = o; o += 3 * (n + 1);
        total = o;
    }
};

// J_task of one level into Js (t x n, leading dimension ld), all threads of the team.  reference src/dwbc.cpp:685-793
__device__ __forceinline__ void build_task_jacobian(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, const Work &w, long long s,
                                                    int level, double *Js, int ld) {
    const int tid = t_id(), nt = t_n();
    const int n = w.n, nb = w.nb;
    const TaskCfg &tc = cfg->tasks[level];
    const int t = tc.dof;
    if (tc.custom) {
        const double *src = w.jcustom[level] + (w.jcustom_per_state[level] ? s * (long long)t * n : 0);
        tcopy(t, n, src, t, Js, ld);
        return;
    }
    int row = 0;
    for (int k = 0; k < tc.link_count; k++) {
        const TaskLinkCfg &tl = cfg->tlinks[tc.link_start + k];
        const int first = (tl.mode == DWBC_TASK_LINK_ROTATION || tl.mode == DWBC_TASK_LINK_ROTATION_CUSTOM_FRAME) ? 3 : 0;
        if (tl.link == nb) {
            // COM pseudo link: jac_ = SI_body^{-1} CMM_  (reference src/dwbc.cpp:331-356)
            //   linear rows  = A[0:3,:] / M
            //   angular rows = Icom^{-1} (skew(c)^T A[0:3,:] + R0 A[3:6,:]),  c = COM - pelvis origin
            const double *A = w.A + s * (long long)n * n;
            const double *co = w.com + s * 16;
            const double *R0 = w.pose + s * (long long)nb * POSE_STRIDE;
            const Vec3 c = ld3(co + 3);
            const double M = co[15];
            const double *I = co + 6;  // row-major symmetric 3x3
            // inverse of the symmetric 3x3 by cofactors
            const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[5] * I[6] - I[3] * I[8], c02 = I[3] * I[7] - I[4] * I[6];
            const double det = I[0] * c00 + I[1] * c01 + I[2] * c02, rdet = 1.0 / det;
            const double c11 = I[0] * I[8] - I[2] * I[6], c12 = I[1] * I[6] - I[0] * I[7], c22 = I[0] * I[4] - I[1] * I[3];
            const double inv[9] = {c00 * rdet, c01 * rdet, c02 * rdet, c01 * rdet, c11 * rdet, c12 * rdet, c02 * rdet, c12 * rdet, c22 * rdet};
            for (int j = tid; j < n; j += nt) {
                const Vec3 al = v3(A[0 + j * n], A[1 + j * n], A[2 + j * n]), aa = v3(A[3 + j * n], A[4 + j * n], A[5 + j * n]);
                const Vec3 lin = (1.0 / M) * al;
                // skew(c)^T al = -c x al = al x c
                const Vec3 h = cross(al, c) + mv(R0, aa);
                const Vec3 ang = mv(inv, h);
                const double v[6] = {lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
                for (int r = 0; r < tl.dof; r++) Js[row + r + j * ld] = v[first + r];
            }
        } else {
            Vec3 pl = v3(0, 0, 0);
            if (tl.mode == DWBC_TASK_LINK_6D_COM_FRAME || tl.mode == DWBC_TASK_LINK_POSITION_COM_FRAME)
                pl = ld3(md->com[tl.link]);
            else if (tl.mode == DWBC_TASK_LINK_6D_CUSTOM_FRAME || tl.mode == DWBC_TASK_LINK_POSITION_CUSTOM_FRAME)
                pl = v3(tl.point[0], tl.point[1], tl.point[2]);
            point_jacobian_t(md, w.pose + s * (long long)nb * POSE_STRIDE, tl.link, pl, Js, ld, row, first, tl.dof, tid, nt);
        }
        row += tl.dof;
    }
}

// SPD inverse of the t x t matrix in S (ld), in place; Scr is scratch of the same size.  Returns min pivot / max diagonal.
// Synchronises on entry and exit.
__device__ __forceinline__ double spd_inverse_small(int t, double *S, int ld, double *Scr, double *rd) {
    tsync();
    double dmax = 0.0;
    for (int i = 0; i < t; i++) {
        double d = S[i + i * ld];
        dmax = d > dmax ? d : dmax;
    }
    const double piv = team_spd_inverse(t, S, ld, Scr, ld, rd);
    return dmax > 0.0 ? piv / dmax : 0.0;
}

__global__ void __launch_bounds__(TEAM, 6) task_space_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, int hqp, int tmax,
                                                             int null_last) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n();
    const long long s = blockIdx.x;
    const int n = w.n, m = w.m;
    const TaskSmem L(n, m, tmax);
    double *sm = (double *)smem;
    double *X1 = sm + L.X1, *X2 = sm + L.X2, *T = sm + L.T, *Nul = sm + L.Nul, *S1 = sm + L.S1, *S2 = sm + L.S2;
    double *rd = sm + L.vec, *fs = rd + n + 1, *tt = fs + n + 1;
    const int ldm = L.ldm;
    double *tau_task = w.tau_task + s * m;
    for (int i = tid; i < m; i += nt) {
        tau_task[i] = 0.0;
        w.tau_contact[s * m + i] = 0.0;
        tt[i] = 0.0;
    }
    if (!(w.status[s] & 1)) return;  // contact stage failed: leave zero torques, DWBC_ST_TASK_OK stays clear (uniform over the CTA)
    const double *P = w.AinvNc + s * (long long)n * n;  // read through L1/L2, used once per level
    const double *Winv = w.Winv + s * (long long)m * m;
    int ok = 1;  // every thread computes the same value
    const int nl = cfg->nlevels;
    for (int l = 0; l < nl; l++) {
        const int t = cfg->tasks[l].dof;
        const int ldt = t | 1;
        double *Js = X1;
        tsync();  // X1 / X2 / T of the previous level are dead
        build_task_jacobian(md, cfg, w, s, l, Js, ldt);
        for (int i = tid; i < t; i += nt) fs[i] = w.fstar[l][s * t + i];
        tsync();
        tcopy(t, n, Js, ldt, w.Jt[l] + s * (long long)t * n, t);
        // T = J (A^{-1} N_c)      (t x n)
        tgemm_mma<2, false, false>(t, n, n, Js, ldt, P, n, [&](int i, int j, double v) { T[i + j * ldt] = v; });
        tsync();
        // Lambda_t = (T J^T)^{-1}  (t x t)
        tgemm_mma<1, false, true>(t, t, n, T, ldt, Js, ldt, [&](int i, int j, double v) { S1[i + j * ldt] = v; });
        const double cnd1 = spd_inverse_small(t, S1, ldt, X2, rd);  // S1 = Lambda_t
        if (!(cnd1 > 1e-14)) ok = 0;
        tcopy(t, t, S1, ldt, w.Lamt[l] + s * (long long)t * t, t);
        // Q = Lambda_t T[:, 6:]    (t x m) -> X1 (J_task is dead)
        double *Qm = X1;
        tgemm_mma<2, false, false>(t, m, t, S1, ldt, T + 6 * ldt, ldt, [&](int i, int j, double v) { Qm[i + j * ldt] = v; });
        tsync();
        // WQ = W^+ Q^T             (m x t)
        double *WQ = X2;
        tgemm_mma<2, false, true>(m, t, m, Winv, m, Qm, ldt, [&](int i, int j, double v) { WQ[i + j * ldm] = v; });
        tsync();
        // (Q W^+ Q^T)^{-1}         (t x t)
        tgemm_mma<1, false, false>(t, t, m, Qm, ldt, WQ, ldm, [&](int i, int j, double v) { S2[i + j * ldt] = v; });
        const double cnd2 = spd_inverse_small(t, S2, ldt, X1, rd);
        if (!(cnd2 > 1e-14)) ok = 0;
        // J_kt = WQ (Q W^+ Q^T)^{-1}   (m x t) -> X1 (Q is dead)
        double *Jkt = X1;
        tgemm_mma<2, false, false>(m, t, t, WQ, ldm, S2, ldt, [&](int i, int j, double v) { Jkt[i + j * ldm] = v; });
        tsync();
        tcopy(m, t, Jkt, ldm, w.Jkt[l] + s * (long long)m * t, m);
        // JL = J_kt Lambda_t       (m x t) -> X2 (WQ is dead)
        double *JL = X2;
        tgemm_mma<2, false, false>(m, t, t, Jkt, ldm, S1, ldt, [&](int i, int j, double v) { JL[i + j * ldm] = v; });
        tsync();
        // torque_h = JL f*  (closed-form branch; the HQP kernel overwrites it with f* + f*_qp)
        tgemv<false>(m, t, JL, ldm, fs, [&](int i, double v) { w.torque_h[l][s * m + i] = v; });
        // Ntorque = Null_{l-1} JL  (m x t) -> X1 (J_kt already stored)
        double *Nt = JL;
        if (l > 0) {
            Nt = X1;
            tgemm_mma<2, false, false>(m, t, m, Nul, ldm, JL, ldm, [&](int i, int j, double v) { Nt[i + j * ldm] = v; });
            tsync();
        }
        tcopy(m, t, Nt, ldm, w.Ntq[l] + s * (long long)m * t, m);
        // closed-form torque: torque_task += Null_{l-1} torque_h = Ntorque f*
        tgemv<false>(m, t, Nt, ldm, fs, [&](int i, double v) { tt[i] += v; });
        // Null_l = Null_{l-1} - Ntorque T[:, 6:]     (not needed for the last level)
        if (l + 1 < nl || null_last) {  // null_last: the reduced variant needs Null of its last contact-chain level
            if (l == 0) {
                tgemm_mma<2, false, false>(m, m, t, Nt, ldm, T + 6 * ldt, ldt,
                                          [&](int i, int j, double v) { Nul[i + j * ldm] = (i == j ? 1.0 : 0.0) - v; });
            } else {
                tgemm_mma<2, false, false>(m, m, t, Nt, ldm, T + 6 * ldt, ldt, [&](int i, int j, double v) { Nul[i + j * ldm] -= v; });
            }
            tsync();
            tcopy(m, m, Nul, ldm, w.Null[l] + s * (long long)m * m, m);
        }
    }
    tsync();
    if (!hqp) {
        for (int i = tid; i < m; i += nt) tau_task[i] = tt[i];
    }
    if (tid == 0) {
        // hqp = 0: CalcTaskControlTorque always returns 1 on this branch (reference src/dwbc.cpp:871); a singular
        // task space is still reported.  hqp = 1: the bit is decided by the QP kernel, cleared here on failure.
        int st = w.status[s] & 1;
        if (ok && !hqp) st |= 2;
        if (ok && hqp) st |= 8;  // internal: task space valid, consumed by hqp_kernel
        w.status[s] = st;
    }
}

}  // namespace dwbc
