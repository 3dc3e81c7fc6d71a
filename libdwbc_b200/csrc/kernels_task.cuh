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

// Shared-memory arena of one state.  The level-independent quantities of ALL hierarchy levels are stacked (rows of level l
// start at row r_l = sum of the dofs above it), so that every phase between two CTA barriers works on the whole hierarchy.
struct TaskSmem {
    int tsum, ldt, ldm, X1, X2, T, S1, S2, vec, total, pstage;
    __host__ __device__ static int sblock(const TaskCfg *tasks, int nl) {
        int o = 0;
        for (int l = 0; l < nl; l++) o += (tasks[l].dof | 1) * tasks[l].dof;
        return o;
    }
    __host__ __device__ TaskSmem(int n, int m, const TaskCfg *tasks, int nl) {
        tsum = 0;
        for (int l = 0; l < nl; l++) tsum += tasks[l].dof;
        ldt = tsum | 1;
        ldm = m | 1;
        const int sb = sblock(tasks, nl);
        int x1 = ldt * n > ldm * tsum ? ldt * n : ldm * tsum;
        int x2 = ldm * tsum > sb ? ldm * tsum : sb;
        x1 = x1 > sb ? x1 : sb;
        int o = 0;
        X1 = o; o += x1;   // J_task (all levels) -> Q -> scratch of the second inverses -> J_kt
        T = o; o += ldt * n;   // J (A^{-1} N_c), all levels
        X2 = o; o += x2;   // scratch of the first inverses -> W^+ Q^T -> J_kt Lambda_t -> Ntorque
        S1 = o; o += sb;   // Lambda_t blocks
        int cs = 0, rr = 0;  // the null-space recursion keeps T[0:r_l, 6:] JL_l ((r_l | 1) x t_l) where the S2 blocks were
        for (int l = 0; l < nl; l++) {
            cs = (rr | 1) * tasks[l].dof > cs ? (rr | 1) * tasks[l].dof : cs;
            rr += tasks[l].dof;
        }
        S2 = o; o += sb > cs ? sb : cs;   // (Q W^+ Q^T)^{-1} blocks
        pstage = (S2 + sb - X2 >= n * n) ? X2 : -1;  // X2..S2 are free until T exists: room to stage A^{-1}N_c for the first product
        vec = o; o += 2 * (tsum + 1) + (m + 1) + MAX_TASK_DOF + 4;  // f*, condition numbers, closed-form torque, 1/diag of a team Cholesky
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

// SPD inverse of the t x t matrix in S (ld), in place, by the whole team; Scr is scratch of the same size.  Returns min pivot /
// max diagonal.  Synchronises on entry and exit.  (Used by the reduced-dimension kernels.)
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

constexpr int WARP_INV_MAX = 8;  // largest t x t inverse given to a single warp

// One CTA (team) per state.  Phases (one CTA barrier each) work on the whole hierarchy at once; only the null-space recursion
// (Ntorque_l = Null_{l-1} J_kt Lambda_t, Null_l = Null_{l-1} - Ntorque_l T_l[:,6:]) is sequential over the levels, and Null itself
// lives in HBM/L2 (it is a public RobotData field that is written there anyway).  The t x t inverses of different levels run on
// different warps at the same time (warp_spd_inverse).  The per-level version of this kernel went through ~126 CTA barriers
// per state (2 (t + 5) of them inside the SPD inverses) and ncu attributed most of its warp time to waiting at them
// (profiles/r01_ncu_dmma.md).
__global__ void __launch_bounds__(TEAM, 6) task_space_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, int hqp, int null_last) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n(), warp = tid >> 5, nw = nt >> 5;
    const long long s = blockIdx.x;
    const int n = w.n, m = w.m;
    const int nl = cfg->nlevels;
    const TaskSmem L(n, m, cfg->tasks, nl);
    double *sm = (double *)smem;
    double *X1 = sm + L.X1, *X2 = sm + L.X2, *T = sm + L.T, *S1 = sm + L.S1, *S2 = sm + L.S2;
    double *fs = sm + L.vec, *cnd = fs + L.tsum + 1, *tt = cnd + L.tsum + 1, *rdv = tt + m + 1;
    const int ldm = L.ldm, ldt = L.ldt;
    double *tau_task = w.tau_task + s * m;
    for (int i = tid; i < m; i += nt) {
        tau_task[i] = 0.0;
        w.tau_contact[s * m + i] = 0.0;
        tt[i] = 0.0;
    }
    if (!(w.status[s] & 1)) return;  // contact stage failed: leave zero torques, DWBC_ST_TASK_OK stays clear (uniform over the CTA)
    const double *P = w.AinvNc + s * (long long)n * n;  // read through L1/L2
    const double *Winv = w.Winv + s * (long long)m * m;
    // per-level offsets: rows r0[l] in the stacked matrices, so[l] in the S1 / S2 block arenas
    int r0[MAX_LEVELS], so[MAX_LEVELS], td[MAX_LEVELS];
    {
        int r = 0, o = 0;
        for (int l = 0; l < nl; l++) {
            td[l] = cfg->tasks[l].dof;
            r0[l] = r;
            so[l] = o;
            r += td[l];
            o += (td[l] | 1) * td[l];
        }
    }
    // ---- phase 0: J_task of every level (stacked), f* ----
    double *Js = X1;
    for (int l = 0; l < nl; l++) {
        build_task_jacobian(md, cfg, w, s, l, Js + r0[l], ldt);
        for (int i = tid; i < td[l]; i += nt) fs[r0[l] + i] = w.fstar[l][s * td[l] + i];
    }
    // A^{-1}N_c is read by every warp of the first product: one coalesced copy into shared memory instead of 8-byte fragment
    // loads from L2 in the k loop (long_scoreboard was this kernel's top stall)
    const double *Pop = P;
    if (L.pstage >= 0) {
        double *Ps = sm + L.pstage;
        for (int e = tid; e < n * n; e += nt) Ps[e] = P[e];
        Pop = Ps;
    }
    tsync();
    for (int l = 0; l < nl; l++) tcopy(td[l], n, Js + r0[l], ldt, w.Jt[l] + s * (long long)td[l] * n, td[l]);
    // ---- phase 1: T = J (A^{-1} N_c)   (tsum x n) ----
    tgemm_mma<2, false, false>(L.tsum, n, n, Js, ldt, Pop, n, [&](int i, int j, double v) { T[i + j * ldt] = v; });
    tsync();
    // ---- phase 2: Lambda_t^{-1} = T_l J_l^T   (t x t per level) ----
    for (int l = 0; l < nl; l++) {
        const int t = td[l], lds = t | 1;
        double *S = S1 + so[l];
        tgemm_mma<1, false, true>(t, t, n, T + r0[l], ldt, Js + r0[l], ldt, [&](int i, int j, double v) { S[i + j * lds] = v; }, l);
    }
    tsync();
    // ---- phase 3: Lambda_t, one level per warp ----
    // small levels: one per warp, concurrently (only __syncwarp inside); large levels: the whole team, one after another
    for (int l = warp; l < nl; l += nw) {
        const int t = td[l], lds = t | 1;
        if (t > WARP_INV_MAX) continue;
        const double c1 = warp_spd_inverse(t, S1 + so[l], lds, X2 + so[l], lds);
        if (lane_id() == 0) cnd[l] = c1;
    }
    for (int l = 0; l < nl; l++) {
        const int t = td[l], lds = t | 1;
        if (t <= WARP_INV_MAX) continue;
        const double c1 = spd_inverse_small(t, S1 + so[l], lds, X2 + so[l], rdv);
        if (tid == 0) cnd[l] = c1;
    }
    tsync();
    // ---- phase 4: Q_l = Lambda_t T_l[:, 6:]   (stacked tsum x m) -> X1 (J_task is dead) ----
    double *Qm = X1;
    for (int l = 0; l < nl; l++) {
        const int t = td[l], lds = t | 1;
        tcopy(t, t, S1 + so[l], lds, w.Lamt[l] + s * (long long)t * t, t);
        tgemm_mma<2, false, false>(t, m, t, S1 + so[l], lds, T + r0[l] + 6 * ldt, ldt, [&](int i, int j, double v) { Qm[r0[l] + i + j * ldt] = v; }, l);
    }
    tsync();
    // ---- phase 5: WQ = W^+ Q^T   (m x tsum) ----
    double *WQ = X2;
    tgemm_mma<2, false, true>(m, L.tsum, m, Winv, m, Qm, ldt, [&](int i, int j, double v) { WQ[i + j * ldm] = v; });
    tsync();
    // ---- phase 6: Q_l W^+ Q_l^T   (t x t per level) ----
    for (int l = 0; l < nl; l++) {
        const int t = td[l], lds = t | 1;
        double *S = S2 + so[l];
        tgemm_mma<1, false, false>(t, t, m, Qm + r0[l], ldt, WQ + r0[l] * ldm, ldm, [&](int i, int j, double v) { S[i + j * lds] = v; }, l);
    }
    tsync();
    // ---- phase 7: (Q W^+ Q^T)^{-1}, one level per warp (scratch: X1, Q is dead) ----
    for (int l = warp; l < nl; l += nw) {
        const int t = td[l], lds = t | 1;
        if (t > WARP_INV_MAX) continue;
        const double c2 = warp_spd_inverse(t, S2 + so[l], lds, X1 + so[l], lds);
        if (lane_id() == 0) cnd[l] = fmin(cnd[l], c2);
    }
    for (int l = 0; l < nl; l++) {
        const int t = td[l], lds = t | 1;
        if (t <= WARP_INV_MAX) continue;
        const double c2 = spd_inverse_small(t, S2 + so[l], lds, X1 + so[l], rdv);
        if (tid == 0) cnd[l] = fmin(cnd[l], c2);
    }
    tsync();
    int ok = 1;  // every thread computes the same value
    for (int l = 0; l < nl; l++)
        if (!(cnd[l] > 1e-14)) ok = 0;
    // ---- phase 8: J_kt = WQ_l (Q W^+ Q^T)^{-1}   (m x t per level, stacked m x tsum) -> X1 ----
    double *Jkt = X1;
    for (int l = 0; l < nl; l++) {
        const int t = td[l], lds = t | 1;
        tgemm_mma<2, false, false>(m, t, t, WQ + r0[l] * ldm, ldm, S2 + so[l], lds, [&](int i, int j, double v) { Jkt[i + (r0[l] + j) * ldm] = v; }, l);
    }
    tsync();
    // ---- phase 9: JL = J_kt Lambda_t   (m x t per level) -> X2 (WQ is dead) ----
    double *JL = X2;
    for (int l = 0; l < nl; l++) {
        const int t = td[l], lds = t | 1;
        tcopy(m, t, Jkt + r0[l] * ldm, ldm, w.Jkt[l] + s * (long long)m * t, m);
        tgemm_mma<2, false, false>(m, t, t, Jkt + r0[l] * ldm, ldm, S1 + so[l], lds, [&](int i, int j, double v) { JL[i + (r0[l] + j) * ldm] = v; }, l);
    }
    tsync();
    // torque_h = JL f*  (closed-form branch; the HQP kernel overwrites it with f* + f*_qp)
    for (int l = 0; l < nl; l++)
        tgemv<false>(m, td[l], JL + r0[l] * ldm, ldm, fs + r0[l], [&](int i, double v) { w.torque_h[l][s * m + i] = v; });
    // ---- sequential part, without ever reading Null back: with Null_{l-1} = I - sum_{k<l} Ntorque_k T_k[:,6:],
    //   Ntorque_l = Null_{l-1} JL_l = JL_l - Nta[:, 0:r_l] (T[0:r_l, 6:] JL_l)      (Nta: Ntorque of the levels above, stacked)
    //   Null_l    = I - Nta[:, 0:r_{l+1}] T[0:r_{l+1}, 6:]
    // Every operand is in shared memory and the inner dimensions are the (small) task dofs instead of m. ----
    double *Nta = X1;  // J_kt is stored
    double *Cs = S2;   // (Q W^+ Q^T)^{-1} is dead
    for (int l = 0; l < nl; l++) {
        const int t = td[l], rl = r0[l];
        const double *Nt = JL + rl * ldm;
        if (l == 0) {
            tcopy(m, t, JL, ldm, Nta, ldm);
        } else {
            const int ldc = rl | 1;
            tgemm_mma<1, false, false>(rl, t, m, T + 6 * ldt, ldt, JL + rl * ldm, ldm, [&](int i, int j, double v) { Cs[i + j * ldc] = v; });
            tsync();
            tgemm_mma<2, false, false>(m, t, rl, Nta, ldm, Cs, ldc, [&](int i, int j, double v) { Nta[i + (rl + j) * ldm] = JL[i + (rl + j) * ldm] - v; });
            tsync();
            Nt = Nta + rl * ldm;
        }
        tcopy(m, t, Nt, ldm, w.Ntq[l] + s * (long long)m * t, m);
        // closed-form torque: torque_task += Null_{l-1} torque_h = Ntorque f*
        tgemv<false>(m, t, Nt, ldm, fs + rl, [&](int i, double v) { tt[i] += v; });
        if (l + 1 < nl || null_last) {  // null_last: the reduced variant needs Null of its last contact-chain level
            double *Nout = w.Null[l] + s * (long long)m * m;
            tgemm_mma<2, false, false>(m, m, rl + t, l == 0 ? JL : Nta, ldm, T + 6 * ldt, ldt,
                                       [&](int i, int j, double v) { Nout[i + j * m] = (i == j ? 1.0 : 0.0) - v; });
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
