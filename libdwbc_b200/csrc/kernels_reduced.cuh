// Stage a9: reduced-dimension variant (Ahn 2024).  One CTA (team) per robot state.
//
// Reference: RobotData::ReducedDynamicsCalculate (src/dwbc.cpp:2752-2989), ReducedCalcContactConstraint (:3077-3142),
// ReducedCalcGravCompensation (:3144-3150), ReducedCalcTaskSpace (:3152-3253), TaskSpace::CalcJKT_R (src/task.cpp:95-142),
// CalculateJKT_R (src/wbd.cpp:220-226), ReducedCalcTaskControlTorque hqp=false branch (:3374-3417) and the expansion
// back to the 33 joints (:3442-3443).
//
// The system is split into the virtual-contact coordinates vc = [base 6 | contact-chain joints] and the non-contact
// chain nc, which is condensed to a 6-dof rigid body seen from the pelvis: n 39 -> 24, m 33 -> 18 in double support.
// Like the reference (its block indexing `rightCols(nc_dof)`, `A_inv_.block(0, vc_dof, ...)`), this needs the q order
// [base | contact chain | non-contact chain] to be contiguous: double support or left-foot single support.
//
// Restructuring (same results up to rounding):
//   * the nc-restricted CRBA of the reference is not repeated: the subtree of an nc joint is entirely nc, so
//     A_NC_O[0:6, nc] = A[0:6, nc] and A_NC_O[nc, nc] = A[nc, nc]; only the composite inertia of the nc chain at the
//     pelvis (SI_nc_b_) is new, and it is a plain sum over the nc bodies in the pelvis frame;
//   * J_CR = [J_C[:, :vc] | 0] makes ReducedCalcContactConstraint the generic contact stage on (A_R_inv, J_CR): the
//     existing contact_dynamics_kernel runs on a second workspace with n = 24, m = 18 (Lambda_CR = Lambda_contact,
//     J_CR_INV_T = J_C_INV_T J_R^T hold identically), and the contact-chain / COM task levels run through the
//     existing task_space_kernel with their reduced Jacobians as per-state custom Jacobians;
//   * non-contact tasks need (J A^{-1} N_C J^T)^{-1} on the FULL model; A^{-1} N_C is never formed:
//     J A^{-1} N_C J^T = (J A^{-1}) J^T - (J X) Lambda_c (J X)^T with X = A^{-1} J_C^T, and
//     A^{-1} N_C v = y - A^{-1} J_C^T Lambda_c J_C y with y = A^{-1} v.
#pragma once
#include "batch.h"
#include "kernels_contact.cuh"
#include "kernels_task.cuh"
#include "team_linalg.cuh"

namespace dwbc {

constexpr int NCMAX = MAXB - 1;  // capacity of the non-contact chain (joints)

// host-decided structure of the reduced model for one contact pattern (shared by every state of the call)
struct ReducedCfg {
    int vc_dof, nc_dof, co_dof, rn, rm;  // rn = reduced_system_dof_, rm = reduced_model_dof_
    unsigned long long nc_mask;           // bit b set <=> body b belongs to the non-contact chain
    int ncontact_active;
    int pattern[MAX_ACTIVE_CONTACTS];     // registered index of each active contact of the pattern
    int nlevels;                          // all hierarchy levels
    int k_contact;                        // levels [0, k) are COM / contact-chain tasks, [k, nlevels) non-contact tasks
    int cls[MAX_LEVELS];                  // 0 cmm_task, 1 reduced_task_, 2 noncont_task
};

// per-state extras of the reduced model (the second Work carries everything that has a full-model twin)
struct ReducedX {
    double *JI;       // [B][6*NCMAX]  J_I_nc_       6 x nc
    double *JIinvT;   // [B][6*NCMAX]  J_I_nc_inv_T  6 x nc
    double *SInc;     // [B][16]       mass_nc, com_pos_nc (3), inertia_nc_ (9 row-major), pad
    double *AR;       // [B][rn*rn]    A_R
    double *Jbkt;     // [B][rm*6]     J_base_R_kt_
    double *JtR[MAX_LEVELS];    // [B][t*rn]  J_task_R_ of the COM / contact-chain levels
    double *LamNC[MAX_LEVELS];  // [B][t*t]   Lambda_task_ of the non-contact levels
    double *tauR;     // [B][rm]  torque_task_R_
    double *tauNC;    // [B][NCMAX] torque_task_NC_
    // CalcSingleTaskTorqueWithQP_R_NC (reference src/dwbc.cpp:3601-3756): inputs prepared by reduced_nc_kernel for hqp_kernel
    double *fonR;     // [B][6]     force_on_nc_r_   (plays f*)
    double *tprevNC;  // [B][rm]    torque_grav_R_ + torque_task_R_ of the COM / contact-chain levels  (plays torque_grav_)
    double *NtqNC;    // [B][rm*6]  Null_{first_nc-1} J_base_R_kt_   (plays Ntorque)
    int *stNC;        // [B]        status word of the base-wrench QP
    double *fqpNC;    // [B][6]     force_on_nc_R_qp_
};

__host__ __device__ inline int reduced_dyn_smem_doubles(int rn, int nc) {
    const int ld = rn | 1;
    return 2 * ld * rn + 3 * 7 * nc + MAXB * 10 + 64 + rn + 1;
}

// ---------------------------------------------------------------- ReducedDynamicsCalculate
__global__ void __launch_bounds__(TEAM, 4) reduced_dynamics_kernel(const DevModel *__restrict__ md, Work w, Work wr, ReducedX rx, ReducedCfg rc) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n();
    const long long s = blockIdx.x;
    const int n = w.n, nb = w.nb, rn = rc.rn, vc = rc.vc_dof, nc = rc.nc_dof;
    const int ld = rn | 1, ldj = 7;
    double *sm = (double *)smem;
    double *L = sm, *Xs = L + ld * rn, *JI = Xs + ld * rn, *K = JI + ldj * nc, *JT = K + ldj * nc;
    double *acc = JT + ldj * nc, *si = acc + MAXB * 10, *rd = si + 64;
    // the state must show the contact pattern the reduced model was built for
    int ok = w.status[s] != 0 && w.cdof[s] == 6 * rc.ncontact_active;
    for (int k = 0; k < rc.ncontact_active; k++) ok = ok && (w.cidx[s * MAX_ACTIVE_CONTACTS + k] == rc.pattern[k]);
    if (!ok) {
        if (tid == 0) wr.status[s] = 0;
        return;
    }
    const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
    const double *A = w.A + s * (long long)n * n, *Ainv = w.Ainv + s * (long long)n * n;
    // composite inertia of the nc chain about the pelvis origin, pelvis frame (model_.Ic[2], reference :2828-2887)
    for (int b = tid; b < nb; b += nt) {
        double *a = acc + b * 10;
        if ((rc.nc_mask >> b) & 1ull) {
            const double *pb = pose + b * POSE_STRIDE;
            const double mb = md->mass[b];
            const Vec3 cw = ld3(pb + 9) + mv(pb, ld3(md->com[b])) - ld3(pose + 9);
            const Vec3 c = mtv(pose, cw);
            double Rrel[9], tmp[9], Ip[9], R0t[9];
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) R0t[i * 3 + j] = pose[j * 3 + i];
            mm3(R0t, pb, Rrel);
            mm3(Rrel, md->inertia[b], tmp);
            for (int i = 0; i < 3; i++)
                for (int j = 0; j < 3; j++) Ip[i * 3 + j] = tmp[i * 3] * Rrel[j * 3] + tmp[i * 3 + 1] * Rrel[j * 3 + 1] + tmp[i * 3 + 2] * Rrel[j * 3 + 2];
            const double cc = dot(c, c);
            a[0] = mb;
            a[1] = mb * c.x; a[2] = mb * c.y; a[3] = mb * c.z;
            a[4] = Ip[0] + mb * (cc - c.x * c.x); a[5] = Ip[1] - mb * c.x * c.y; a[6] = Ip[2] - mb * c.x * c.z;
            a[7] = Ip[4] + mb * (cc - c.y * c.y); a[8] = Ip[5] - mb * c.y * c.z; a[9] = Ip[8] + mb * (cc - c.z * c.z);
        } else {
            for (int k = 0; k < 10; k++) a[k] = 0.0;
        }
    }
    tsync();
    if (tid < 10) {
        double v = 0.0;
        for (int b = 0; b < nb; b++) v += acc[b * 10 + tid];
        si[tid] = v;
    }
    tsync();
    // InertiaMatrixSegment (reference src/math.cpp:304-311): mass, COM, inertia about the COM and its inverse
    const double M = si[0];
    const Vec3 c = v3(si[1] / M, si[2] / M, si[3] / M);
    const double cc = dot(c, c);
    const double I[9] = {si[4] - M * (cc - c.x * c.x), si[5] + M * c.x * c.y, si[6] + M * c.x * c.z,
                         si[5] + M * c.x * c.y, si[7] - M * (cc - c.y * c.y), si[8] + M * c.y * c.z,
                         si[6] + M * c.x * c.z, si[8] + M * c.y * c.z, si[9] - M * (cc - c.z * c.z)};
    double inv[9];
    {
        const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[5] * I[6] - I[3] * I[8], c02 = I[3] * I[7] - I[4] * I[6];
        const double det = I[0] * c00 + I[1] * c01 + I[2] * c02, rdet = 1.0 / det;
        const double c11 = I[0] * I[8] - I[2] * I[6], c12 = I[1] * I[6] - I[0] * I[7], c22 = I[0] * I[4] - I[1] * I[3];
        inv[0] = c00 * rdet; inv[1] = c01 * rdet; inv[2] = c02 * rdet;
        inv[3] = c01 * rdet; inv[4] = c11 * rdet; inv[5] = c12 * rdet;
        inv[6] = c02 * rdet; inv[7] = c12 * rdet; inv[8] = c22 * rdet;
    }
    if (tid == 0) {
        double *o = rx.SInc + s * 16;
        o[0] = M; o[1] = c.x; o[2] = c.y; o[3] = c.z;
        for (int k = 0; k < 9; k++) o[4 + k] = I[k];
    }
    // J_I_nc_ = SI_nc_l^{-1} [1 0; skew(c)^T 1] [R0^T A[0:3, nc]; A[3:6, nc]]   (6 x nc), reference :2906-2916
    for (int j = tid; j < nc; j += nt) {
        const double *col = A + (long long)(vc + j) * n;
        const Vec3 al = mtv(pose, v3(col[0], col[1], col[2])), aa = v3(col[3], col[4], col[5]);
        const Vec3 lin = (1.0 / M) * al;
        const Vec3 ang = mv(inv, cross(al, c) + aa);  // skew(c)^T al = al x c
        double *o = JI + j * ldj;
        o[0] = lin.x; o[1] = lin.y; o[2] = lin.z; o[3] = ang.x; o[4] = ang.y; o[5] = ang.z;
    }
    tsync();
    tcopy(6, nc, JI, ldj, rx.JI + s * 6 * NCMAX, 6);
    // K = J_I A^{-1}[nc, nc]  (6 x nc)
    tgemm<1, 1, false, false>(6, nc, nc, JI, ldj, Ainv + vc + (long long)vc * n, n, [&](int i, int j, double v) { K[i + j * ldj] = v; });
    // A_R_inv = J_R A^{-1} J_R^T assembled blockwise (reference :2936-2956)
    tcopy(vc, vc, Ainv, n, L, ld);
    tgemm<1, 1, false, true>(vc, 6, nc, Ainv + (long long)vc * n, n, JI, ldj, [&](int i, int j, double v) {
        L[i + (vc + j) * ld] = v;
        L[(vc + j) + i * ld] = v;
    });
    tsync();
    tgemm<1, 1, false, true>(6, 6, nc, K, ldj, JI, ldj, [&](int i, int j, double v) { L[(vc + i) + (vc + j) * ld] = v; });
    tsync();
    tcopy(rn, rn, L, ld, wr.Ainv + s * (long long)rn * rn, rn);
    // A_R = A_R_inv^{-1} (reference :2958, LLT)
    const double minpiv = team_spd_inverse(rn, L, ld, Xs, ld, rd);
    tcopy(rn, rn, L, ld, rx.AR + s * (long long)rn * rn, rn);
    // J_I_nc_inv_T = A_R[vc:, :vc] A^{-1}[:vc, nc] + A_R[vc:, vc:] K   (6 x nc), reference :2971
    tgemm<1, 1, false, false>(6, nc, vc, L + vc, ld, Ainv + (long long)vc * n, n, [&](int i, int j, double v) { JT[i + j * ldj] = v; });
    tsync();
    tgemm<1, 1, false, false>(6, nc, 6, L + vc + vc * ld, ld, K, ldj, [&](int i, int j, double v) { JT[i + j * ldj] += v; });
    tsync();
    tcopy(6, nc, JT, ldj, rx.JIinvT + s * 6 * NCMAX, 6);
    // G_R = [G[:vc]; J_I_nc_inv_T G[nc]], J_CR = [J_C[:, :vc] | 0]   (reference :2982-2987)
    const double *G = w.G + s * n;
    for (int i = tid; i < vc; i += nt) wr.G[s * rn + i] = G[i];
    tgemv<false>(6, nc, JT, ldj, G + vc, [&](int i, double v) { wr.G[s * rn + vc + i] = v; });
    {
        const double *Jc = w.Jc + s * (long long)CMAX * n;
        double *o = wr.Jc + s * (long long)CMAX * rn;
        for (int e = tid; e < CMAX * rn; e += nt) o[e] = (e < CMAX * vc) ? Jc[e] : 0.0;
    }
    if (tid == 0) wr.status[s] = (minpiv > 0.0) ? -1 : 0;
}

// ---------------------------------------------------------------- reduced Jacobians of the COM / contact-chain levels
// J_task_R = [J[:, :vc] | J[:, nc] J_I_nc_inv_T^T] (cmm_task) or [J[:, :vc] | 0] (reduced_task_), reference src/task.cpp:107-121
__global__ void __launch_bounds__(TEAM, 4) reduced_task_jacobian_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, Work wr,
                                                                        ReducedX rx, ReducedCfg rc, int tmax) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n();
    const long long s = blockIdx.x;
    const int n = w.n, rn = rc.rn, vc = rc.vc_dof, nc = rc.nc_dof;
    if (!(wr.status[s] & 1)) return;
    double *Js = (double *)smem;  // (tmax|1) x n
    const double *JT = rx.JIinvT + s * 6 * NCMAX;
    for (int l = 0; l < rc.k_contact; l++) {
        const int t = cfg->tasks[l].dof, ldt = t | 1;
        tsync();
        build_task_jacobian(md, cfg, w, s, l, Js, ldt);
        tsync();
        tcopy(t, n, Js, ldt, w.Jt[l] + s * (long long)t * n, t);
        double *o = rx.JtR[l] + s * (long long)t * rn;
        for (int e = tid; e < t * vc; e += nt) {
            int j = e / t, i = e - j * t;
            o[e] = Js[i + j * ldt];
        }
        if (rc.cls[l] == 0) {
            tgemm<1, 1, false, true>(t, 6, nc, Js + vc * ldt, ldt, JT, 6, [&](int i, int j, double v) { o[i + (vc + j) * t] = v; });
        } else {
            for (int e = tid; e < t * 6; e += nt) o[t * vc + e] = 0.0;
        }
    }
}

// ---------------------------------------------------------------- non-contact tasks, J_base_R_kt_, expansion to 33 joints
struct ReducedNcSmem {
    int ldt, J, Jp, JA, JX, M2, Lam, Lamp, Scr, Jbkt, S6, Q6, vec, vlen, total;
    __host__ __device__ ReducedNcSmem(int n, int rm, int tmax) {
        ldt = tmax | 1;
        const int ldr = rm | 1;
        int o = 0;
        J = o; o += ldt * n;
        Jp = o; o += ldt * n;
        JA = o; o += (ldt * n > ldr * 6 ? ldt * n : ldr * 6);  // J A^{-1}; also W_R^+ Q^T of the base task
        JX = o; o += ldt * CMAX;
        M2 = o; o += ldt * CMAX;
        Lam = o; o += ldt * tmax;
        Lamp = o; o += ldt * tmax;
        Scr = o; o += ldt * tmax;
        Jbkt = o; o += ldr * 6;
        S6 = o; o += 3 * 7 * 6;
        Q6 = o; o += 2 * 7 * rm;
        vlen = n + 1;  // >= MAX_TASK_DOF, CMAX
        vec = o; o += 16 * vlen;
        total = o;
    }
};

// qp_mode bit 0: the COM / contact-chain levels went through hqp_kernel (CalcSingleTaskTorqueWithQP_R); bit 1: so did the
// contact redistribution (CalcContactRedistributeR): torque_contact_[0:contact_dof_] = torque_contact_R_[0:contact_dof_].
__global__ void __launch_bounds__(TEAM, 4) reduced_nc_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, Work wr, ReducedX rx,
                                                             ReducedCfg rc, int tmax, int qp_mode) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n();
    const long long s = blockIdx.x;
    const int n = w.n, m = w.m, nb = w.nb, rn = rc.rn, rm = rc.rm, vc = rc.vc_dof, nc = rc.nc_dof, co = rc.co_dof;
    const ReducedNcSmem Ls(n, rm, tmax);
    double *sm = (double *)smem;
    double *Jcur = sm + Ls.J, *Jprev = sm + Ls.Jp, *JA = sm + Ls.JA, *JX = sm + Ls.JX, *M2 = sm + Ls.M2;
    double *Lam = sm + Ls.Lam, *Lamp = sm + Ls.Lamp, *Scr = sm + Ls.Scr, *Jbkt = sm + Ls.Jbkt;
    double *S6 = sm + Ls.S6, *S6b = S6 + 42, *S6c = S6b + 42, *Q6 = sm + Ls.Q6, *Q6a = Q6 + 7 * rm;
    const int VL = Ls.vlen;
    double *rd = sm + Ls.vec, *tmp = rd + VL, *y1 = tmp + VL, *uu = y1 + VL, *tmp2 = uu + VL, *tauR = tmp2 + VL, *tauNC = tauR + VL;
    double *thR = tauNC + VL, *ntr = thR + VL, *z = ntr + VL, *z2 = z + VL, *lf = z2 + VL, *fsv = lf + VL, *nf = fsv + VL, *fon = nf + VL;
    const int ldr = rm | 1;
    double *fonr = fon + 8;  // force_on_nc_r_ (6 entries in the tail of fon's slot)
    if (tid < 6) fonr[tid] = 0.0;
    double *tau_task = w.tau_task + s * m, *tau_grav = w.tau_grav + s * m;
    for (int i = tid; i < m; i += nt) tau_task[i] = 0.0, w.tau_contact[s * m + i] = 0.0;
    const int st = wr.status[s];
    if (!(st & 1)) {  // reduced contact stage failed or the state does not show the pattern
        for (int i = tid; i < m; i += nt) tau_grav[i] = 0.0;
        if (tid == 0) w.status[s] = 0;
        return;
    }
    const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
    const double *Ainv = w.Ainv + s * (long long)n * n, *Jc = w.Jc + s * (long long)CMAX * n;
    const double *LamC = wr.Lam + s * CMAX * CMAX;  // Lambda_CR = Lambda_contact
    const double *JI = rx.JI + s * 6 * NCMAX, *JT = rx.JIinvT + s * 6 * NCMAX;
    const int c = w.cdof[s];
    // torque_grav_ = [torque_grav_R_[0:co]; G[nc]]   (reference :3146-3148: the second assignment overwrites entries co..rm-1)
    for (int i = tid; i < m; i += nt) tau_grav[i] = i < co ? wr.tau_grav[s * rm + i] : w.G[s * n + 6 + i];
    for (int i = tid; i < rm; i += nt) tauR[i] = wr.tau_task[s * rm + i];  // closed-form sum over the COM / contact-chain levels
    for (int i = tid; i < nc; i += nt) tauNC[i] = 0.0;
    int ok = (st & 2) ? 1 : 0;
    if (rc.nlevels == 0) ok = 0;  // gravity compensation only: DWBC_ST_TASK_OK stays clear
    const int nl = rc.nlevels, k0 = rc.k_contact;
    if (k0 < nl) {
        // ---- J_base_R_kt_: CalculateJKT_R on the pelvis 6D Jacobian (reference :3161-3162).  J_base_R = [blkdiag(1, R0) | 0]
        const double *PR = wr.AinvNc + s * (long long)rn * rn, *WRinv = wr.Winv + s * (long long)rm * rm;
        // S6 = Jb P[0:6,0:6] Jb^T with Jb = blkdiag(1, R0): rotate rows and columns 3..5
        auto jb_left = [&](const double *src, int lds, int cols, double *dst, int ldd) {  // dst = Jb * src (6 x cols)
            for (int e = tid; e < 6 * cols; e += nt) {
                int j = e / 6, i = e - j * 6;
                double v;
                if (i < 3) v = src[i + j * lds];
                else v = pose[(i - 3) * 3] * src[3 + j * lds] + pose[(i - 3) * 3 + 1] * src[4 + j * lds] + pose[(i - 3) * 3 + 2] * src[5 + j * lds];
                dst[i + j * ldd] = v;
            }
        };
        jb_left(PR, rn, 6, S6b, 7);  // Jb P[0:6, 0:6]
        tsync();
        for (int e = tid; e < 36; e += nt) {  // (..) Jb^T
            int j = e / 6, i = e - j * 6;
            double v;
            if (j < 3) v = S6b[i + j * 7];
            else v = pose[(j - 3) * 3] * S6b[i + 3 * 7] + pose[(j - 3) * 3 + 1] * S6b[i + 4 * 7] + pose[(j - 3) * 3 + 2] * S6b[i + 5 * 7];
            S6[i + j * 7] = v;
        }
        const double cb1 = spd_inverse_small(6, S6, 7, S6c, rd);  // lambda_base_R_
        if (!(cb1 > 1e-14)) ok = 0;
        jb_left(PR + 6 * rn, rn, rm, Q6a, 7);  // Jb P[0:6, 6:]  (6 x rm)
        tsync();
        // Q6 = lambda_base (Jb P[0:6, 6:])   (6 x rm)
        tgemm<1, 1, false, false>(6, rm, 6, S6, 7, Q6a, 7, [&](int i, int j, double v) { Q6[i + j * 7] = v; });
        tsync();
        // WQ = W_R^+ Q6^T  (rm x 6)
        tgemm<1, 1, false, true>(rm, 6, rm, WRinv, rm, Q6, 7, [&](int i, int j, double v) { JA[i + j * ldr] = v; });
        tsync();
        tgemm<1, 1, false, false>(6, 6, rm, Q6, 7, JA, ldr, [&](int i, int j, double v) { S6b[i + j * 7] = v; });
        const double cb2 = spd_inverse_small(6, S6b, 7, S6c, rd);
        if (!(cb2 > 1e-14)) ok = 0;
        tgemm<1, 1, false, false>(rm, 6, 6, JA, ldr, S6b, 7, [&](int i, int j, double v) { Jbkt[i + j * ldr] = v; });
        tsync();
        tcopy(rm, 6, Jbkt, ldr, rx.Jbkt + s * (long long)rm * 6, rm);
        const double *NulR = (k0 > 0) ? wr.Null[k0 - 1] + s * (long long)rm * rm : nullptr;
        int tp = 0;
        for (int l = k0; l < nl; l++) {
            const int t = cfg->tasks[l].dof, ldt = t | 1;
            tsync();
            build_task_jacobian(md, cfg, w, s, l, Jcur, ldt);
            for (int i = tid; i < t; i += nt) fsv[i] = w.fstar[l][s * t + i];
            tsync();
            tcopy(t, n, Jcur, ldt, w.Jt[l] + s * (long long)t * n, t);
            // Lambda_task_ = (J A^{-1} N_C J^T)^{-1} = (JA J^T - JX Lambda_c JX^T)^{-1},  JA = J A^{-1},  JX = JA J_C^T
            tgemm<1, 2, false, false>(t, n, n, Jcur, ldt, Ainv, n, [&](int i, int j, double v) { JA[i + j * ldt] = v; });
            tsync();
            tgemm<1, 1, false, true>(t, c, n, JA, ldt, Jc, CMAX, [&](int i, int j, double v) { JX[i + j * ldt] = v; });
            tsync();
            tgemm<1, 1, false, false>(t, c, c, JX, ldt, LamC, CMAX, [&](int i, int j, double v) { M2[i + j * ldt] = v; });
            tgemm<1, 1, false, true>(t, t, n, JA, ldt, Jcur, ldt, [&](int i, int j, double v) { Lam[i + j * ldt] = v; });
            tsync();
            tgemm<1, 1, false, true>(t, t, c, M2, ldt, JX, ldt, [&](int i, int j, double v) { Lam[i + j * ldt] -= v; });
            const double cn = spd_inverse_small(t, Lam, ldt, Scr, rd);
            if (!(cn > 1e-14)) ok = 0;
            tcopy(t, t, Lam, ldt, rx.LamNC[l] + s * (long long)t * t, t);
            // lf = Lambda f*,  tmp = J^T lf  (n)
            tgemv<false>(t, t, Lam, ldt, fsv, [&](int i, double v) { lf[i] = v; });
            tsync();
            tgemv<true>(n, t, Jcur, ldt, lf, [&](int i, double v) { tmp[i] = v; });
            tsync();
            // force_on_nc_ = [tmp[0:3]; R0 tmp[3:6]],  torque_h_R_ = [J_base_R_kt_[:co] force; J_I_nc_inv_T tmp[nc]]
            if (tid < 6) {
                double v;
                if (tid < 3) v = tmp[tid];
                else v = pose[(tid - 3) * 3] * tmp[3] + pose[(tid - 3) * 3 + 1] * tmp[4] + pose[(tid - 3) * 3 + 2] * tmp[5];
                fon[tid] = v;
                fonr[tid] += v;  // force_on_nc_r_ += force_on_nc_   (reference :3302; only used under hqp)
            }
            tsync();
            tgemv<false>(co, 6, Jbkt, ldr, fon, [&](int i, double v) { thR[i] = v; });
            tgemv<false>(6, nc, JT, 6, tmp + vc, [&](int i, double v) { thR[co + i] = v; });
            tsync();
            if (l == k0) {
                // first non-contact task: torque_null_h_R_ = Null_R torque_h_R_, torque_null_h_nc_ = torque_nc_
                for (int i = tid; i < nc; i += nt) tauNC[i] += tmp[vc + i];
                for (int i = tid; i < rm; i += nt) ntr[i] = thR[i];
            } else {
                // null_force = Lambda_p J_p (A^{-1} N_C tmp)   (reference :3393)
                tgemv<false>(n, n, Ainv, n, tmp, [&](int i, double v) { y1[i] = v; });
                tsync();
                tgemv<false>(c, n, Jc, CMAX, y1, [&](int i, double v) { z[i] = v; });
                tsync();
                tgemv<false>(c, c, LamC, CMAX, z, [&](int i, double v) { z2[i] = v; });
                tsync();
                tgemv<true>(n, c, Jc, CMAX, z2, [&](int i, double v) { uu[i] = v; });
                tsync();
                tgemv<false>(n, n, Ainv, n, uu, [&](int i, double v) { y1[i] -= v; });
                tsync();
                const int ldp = tp | 1;
                tgemv<false>(tp, n, Jprev, ldp, y1, [&](int i, double v) { z[i] = v; });
                tsync();
                tgemv<false>(tp, tp, Lamp, ldp, z, [&](int i, double v) { nf[i] = v; });
                tsync();
                tgemv<true>(n, tp, Jprev, ldp, nf, [&](int i, double v) { tmp2[i] = v; });
                tsync();
                // torque_null_h_nc_ = torque_nc_ - tmp2[nc]; null_torque_h_r = [thR[:co] - Jbkt[:co] [tmp2[0:3]; R0 tmp2[3:6]]; J_I_nc_inv_T torque_null_h_nc_]
                for (int i = tid; i < nc; i += nt) {
                    const double v = tmp[vc + i] - tmp2[vc + i];
                    uu[i] = v;
                    tauNC[i] += v;
                }
                if (tid < 6) {
                    double v;
                    if (tid < 3) v = tmp2[tid];
                    else v = pose[(tid - 3) * 3] * tmp2[3] + pose[(tid - 3) * 3 + 1] * tmp2[4] + pose[(tid - 3) * 3 + 2] * tmp2[5];
                    fon[tid] = v;
                    fonr[tid] -= v;  // force_on_nc_r_ -= temp_torque_[0:6]   (reference :3325)
                }
                tsync();
                tgemv<false>(co, 6, Jbkt, ldr, fon, [&](int i, double v) { ntr[i] = thR[i] - v; });
                tgemv<false>(6, nc, JT, 6, uu, [&](int i, double v) { ntr[co + i] = v; });
            }
            tsync();
            // torque_task_R_ += Null_R ntr   (a non-contact task cannot be the first level: the reference reads ts_[-1])
            if (NulR) {
                tgemv<false>(rm, rm, NulR, rm, ntr, [&](int i, double v) { tauR[i] += v; });
            } else {
                ok = 0;
            }
            tsync();
            // keep this task as "previous" for the next non-contact level
            double *sw = Jcur; Jcur = Jprev; Jprev = sw;
            sw = Lam; Lam = Lamp; Lamp = sw;
            tp = t;
        }
    }
    tsync();
    if ((qp_mode & 1) && k0 < nl && k0 > 0) {
        // inputs of CalcSingleTaskTorqueWithQP_R_NC for hqp_kernel: f* := force_on_nc_r_, Ntorque := Null_{k0-1} J_base_R_kt_,
        // torque_prev := torque_grav_R_ + torque_task_R_ of the contact-chain levels (their sum is still in wr.tau_task)
        const double *NulR = wr.Null[k0 - 1] + s * (long long)rm * rm;
        double *g = rx.NtqNC + s * (long long)rm * 6;
        tgemm<1, 1, false, false>(rm, 6, rm, NulR, rm, Jbkt, ldr, [&](int i, int j, double v) { g[i + j * rm] = v; });
        for (int i = tid; i < rm; i += nt) rx.tprevNC[s * rm + i] = wr.tau_grav[s * rm + i] + wr.tau_task[s * rm + i];
        if (tid < 6) rx.fonR[s * 6 + tid] = fonr[tid];
        if (tid == 0) rx.stNC[s] = ok ? (1 | 8) : 0;
    }
    // ---- expansion (reference :3442-3443): tau[0:co] = tau_R[0:co];  tau[co:] = J_I_nc^T tau_R[co:co+6] + N_I_nc tau_NC,
    //      N_I_nc tau_NC = tau_NC - J_I_nc^T (J_I_nc_inv_T tau_NC)
    tgemv<false>(6, nc, JT, 6, tauNC, [&](int i, double v) { z[i] = tauR[co + i] - v; });
    tsync();
    for (int i = tid; i < rm; i += nt) rx.tauR[s * rm + i] = tauR[i];
    for (int i = tid; i < nc; i += nt) rx.tauNC[s * NCMAX + i] = tauNC[i];
    for (int i = tid; i < co; i += nt) tau_task[i] = tauR[i];
    for (int i = tid; i < nc; i += nt) {
        double v = tauNC[i];
        for (int k = 0; k < 6; k++) v = fma(JI[k + i * 6], z[k], v);
        tau_task[co + i] = v;
    }
    if (qp_mode & 2)  // reference src/dwbc.cpp:3765-3766
        for (int i = tid; i < c && i < rm; i += nt) w.tau_contact[s * m + i] = wr.tau_contact[s * rm + i];
    if (tid == 0) w.status[s] = 1 | (ok ? 2 : 0) | (((qp_mode & 2) && c > 6) ? (st & 4) : 0);  // reference: ret stays 0 for contact_dof_ <= 6
}

// After the base-wrench QP: torque_task_[0:co] += (J_base_R_kt_ force_on_nc_R_qp_)[0:co]  (reference src/dwbc.cpp:3420-3424); a failed
// QP makes ReducedCalcTaskControlTorque return 0 before torque_task_ is assembled (:3399-3403): zero torque, task bit cleared.
__global__ void reduced_ncqp_apply_kernel(Work w, ReducedX rx, ReducedCfg rc) {
    const long long s = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x, m = w.m, rm = rc.rm, co = rc.co_dof;
    const int st = w.status[s];
    if (!(st & 2)) return;  // already failed (or the state does not show the pattern)
    const bool okq = (rx.stNC[s] & 2) != 0;
    const double *Jb = rx.Jbkt + s * (long long)rm * 6, *x = rx.fqpNC + s * 6;
    for (int i = tid; i < m; i += nt) {
        if (!okq) {
            w.tau_task[s * m + i] = 0.0;
        } else if (i < co) {
            double v = 0.0;
            for (int k = 0; k < 6; k++) v = fma(Jb[i + k * rm], x[k], v);
            w.tau_task[s * m + i] += v;
        }
    }
    if (tid == 0 && !okq) w.status[s] = st & ~2;
}

// ReducedCalcContactRedistribute after a hierarchy with non-contact levels: torque_contact_[0:contact_dof_] = torque_contact_R_[0:contact_dof_]
// (reference src/dwbc.cpp:3765-3766) and the redistribution bit of the status word.
__global__ void reduced_redist_apply_kernel(Work w, Work wr, ReducedCfg rc) {
    const long long s = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x, m = w.m, rm = rc.rm;
    const int st = w.status[s];
    if (!(st & 1)) return;
    const int c = w.cdof[s];
    for (int i = tid; i < m; i += nt) w.tau_contact[s * m + i] = (i < c && i < rm) ? wr.tau_contact[s * rm + i] : 0.0;
    if (tid == 0 && c > 6) w.status[s] = st | (wr.status[s] & 4);
}

}  // namespace dwbc
