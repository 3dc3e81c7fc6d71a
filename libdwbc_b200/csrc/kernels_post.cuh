// Post-path outputs of a control cycle: contact wrench, ZMP and the analytic two-foot contact redistribution.
// One warp per robot state (a few hundred flops per state; the operands are the J_C_INV_T / V2 / pose records the
// contact stage left in HBM).
//
// Reference: RobotData::getContactForce -> CalculateContactForce (src/dwbc.cpp:891-896, src/wbd.cpp:268-271),
// RobotData::getZMP (src/dwbc.cpp:898-939), CalcContactRedistribute hqp=false branch (src/dwbc.cpp:1570-1619) with
// ContactRedistributetwomod (src/wbd.cpp:273-404).
#pragma once
#include "batch.h"
#include "kernels_rbd.cuh"
#include "warp_linalg.cuh"

namespace dwbc {

// ContactForce = J_C_INV_T[:, 6:] tau - P_C for the warp's state; tau (m entries) may live in global or shared memory.
// Lane i < c returns row i in `mine` and the rows are also written to out[0..CMAX) (zero padded) when out != null.
__device__ __forceinline__ void warp_contact_force(const Work &w, long long s, int c, const double *tau, double *out) {
    const int lane = lane_id(), n = w.n, m = w.m;
    const double *Jb = w.Jbar + s * (long long)CMAX * n + 6 * CMAX;
    const double *pc = w.PC + s * CMAX;
    for (int i = lane; i < CMAX; i += 32) {
        double a0 = 0.0, a1 = 0.0;
        if (i < c) {
            int j = 0;
            for (; j + 1 < m; j += 2) {
                a0 = fma(Jb[i + j * CMAX], tau[j], a0);
                a1 = fma(Jb[i + (j + 1) * CMAX], tau[j + 1], a1);
            }
            if (j < m) a0 = fma(Jb[i + j * CMAX], tau[j], a0);
            a0 = (a0 + a1) - pc[i];
        }
        out[i] = a0;
    }
}

// mode bit 0: contact force of `tau_in` (null: tau_grav + tau_task + tau_contact) -> cf_out [B x CMAX]
// mode bit 1: ZMP of cf_in (null: the force computed by bit 0) -> zmp_out [B x 3]
__global__ void contact_force_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, int mode,
                                     const double *__restrict__ tau_in, const double *__restrict__ cf_in, double *__restrict__ cf_out,
                                     double *__restrict__ zmp_out) {
    DWBC_DYN_SMEM(smem);
    const int wpb = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = lane_id();
    const long long s = (long long)blockIdx.x * wpb + wid;
    if (s >= w.B) return;
    const int m = w.m, nb = w.nb;
    double *tau = (double *)smem + wid * (MAXB + CMAX), *cf = tau + MAXB;
    const int c = (w.status[s] & 1) ? w.cdof[s] : 0;
    if (mode & 1) {
        for (int i = lane; i < m; i += 32)
            tau[i] = tau_in ? tau_in[s * m + i] : w.tau_grav[s * m + i] + w.tau_task[s * m + i] + w.tau_contact[s * m + i];
        __syncwarp();
        warp_contact_force(w, s, c, tau, cf);
        __syncwarp();
        if (cf_out)
            for (int i = lane; i < CMAX; i += 32) cf_out[s * CMAX + i] = cf[i];
    } else {
        for (int i = lane; i < CMAX; i += 32) cf[i] = cf_in[s * CMAX + i];
        __syncwarp();
    }
    if ((mode & 2) && lane == 0) {
        // reference src/dwbc.cpp:898-939; the k-th ACTIVE contact reads contact_force(6k ..) (the reference indexes with the
        // registered index, i.e. it assumes the active contacts are a prefix of the registered ones)
        const int kc = c / 6;
        const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
        double total_fz = 0.0;
        for (int k = 0; k < kc; k++) total_fz += cf[6 * k + 2];
        Vec3 zmp = v3(0, 0, 0);
        for (int k = 0; k < kc; k++) {
            const ContactCfg &cc = cfg->contacts[w.cidx[s * MAX_ACTIVE_CONTACTS + k]];
            const double *po = pose + cc.link * POSE_STRIDE;
            Vec3 z = ld3(po + 9) + mv(po, v3(cc.point[0], cc.point[1], cc.point[2]));  // xc_pos
            const double fz = cf[6 * k + 2];
            if (!(fz > -1.0e-3)) {
                z.x += -cf[6 * k + 4] / fz;
                z.y += cf[6 * k + 3] / fz;
            }
            zmp = zmp + (fz / total_fz) * z;
        }
        st3(zmp_out + s * 3, zmp);
    }
}

// Closed-form split of the resultant wrench between two feet, reference src/wbd.cpp:273-404 (ContactRedistributetwomod).
// F12 (in) / FR (out): [f1 m1 f2 m2]; returns eta.
__device__ __forceinline__ double redistribute_two(double eta_cust, double footlength, double footwidth, double mu_s, double ratio_x,
                                                   double ratio_y, Vec3 P1, Vec3 P2, const double *F12, double *FR) {
    const Vec3 f1 = ld3(F12), m1 = ld3(F12 + 3), f2 = ld3(F12 + 6), m2 = ld3(F12 + 9);
    const Vec3 fs = f1 + f2, ms = cross(P1, f1) + m1 + cross(P2, f2) + m2;
    const double RF[6] = {fs.x, fs.y, fs.z, ms.x, ms.y, ms.z};
    double eta_lb = 1.0 - eta_cust, eta_ub = eta_cust;
    auto clip = [&](double A, double B, double C) {
        const double a = A * A, b = 2.0 * A * B, c = B * B - C * C;
        const double disc = sqrt(b * b - 4.0 * a * c);
        const double s1 = (-b + disc) / 2.0 / a, s2 = (-b - disc) / 2.0 / a;
        if (s1 > s2) {
            if (s1 < eta_ub) eta_ub = s1;
            if (s2 > eta_lb) eta_lb = s2;
        } else {
            if (s2 < eta_ub) eta_ub = s2;
            if (s1 > eta_lb) eta_lb = s1;
        }
    };
    const double Ax = (P1.z - P2.z) * RF[1] - (P1.y - P2.y) * RF[2], Bx = RF[3] + P2.z * RF[1] - P2.y * RF[2];
    const double Ay = -(P1.z - P2.z) * RF[0] + (P1.x - P2.x) * RF[2], By = RF[4] - P2.z * RF[0] + P2.x * RF[2];
    const double Az = -(P1.x - P2.x) * RF[1] + (P1.y - P2.y) * RF[0], Bz = RF[5] + P2.y * RF[0] - P2.x * RF[1];
    clip(Ax, Bx, ratio_y * footwidth / 2.0 * fabs(RF[2]));
    clip(Ay, By, ratio_x * footlength / 2.0 * fabs(RF[2]));
    clip(Az, Bz, mu_s * fabs(RF[2]));
    const double eta_s = (-RF[3] - P2.z * RF[1] + P2.y * RF[2]) / ((P1.z - P2.z) * RF[1] - (P1.y - P2.y) * RF[2]);
    double eta = eta_s;
    if (eta_s > eta_ub) eta = eta_ub;
    else if (eta_s < eta_lb) eta = eta_lb;
    if ((eta > eta_cust) || (eta < 1.0 - eta_cust)) eta = 0.5;
    FR[0] = eta * RF[0], FR[1] = eta * RF[1], FR[2] = eta * RF[2];
    FR[3] = Ax * eta * eta + Bx * eta;
    FR[4] = Ay * eta * eta + By * eta;
    FR[5] = Az * eta * eta + Bz * eta;
    FR[6] = (1.0 - eta) * RF[0], FR[7] = (1.0 - eta) * RF[1], FR[8] = (1.0 - eta) * RF[2];
    FR[9] = (1.0 - eta) * (Ax * eta + Bx);
    FR[10] = (1.0 - eta) * (Ay * eta + By);
    FR[11] = (1.0 - eta) * (Az * eta + Bz);
    return eta;
}

// CalcContactRedistribute(hqp = false): torque_contact_ = V2^T (J_C_INV_T[6:12, 6:] V2^T)^{-1} (fc_redist - ContactForce)[6:12]
// for contact_dof_ == 12, else torque_contact_ = 0 and the call "returns 0" (DWBC_ST_REDIST_OK stays clear).
// Uses cc_[0] and cc_[1], the first two REGISTERED contacts, like the reference.
__global__ void redistribute_analytic_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w) {
    DWBC_DYN_SMEM(smem);
    const int wpb = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = lane_id();
    const long long s = (long long)blockIdx.x * wpb + wid;
    if (s >= w.B) return;
    const int n = w.n, m = w.m, nb = w.nb;
    constexpr int lda = 7;
    double *tau = (double *)smem + wid * (MAXB + CMAX + 6 * lda + 8), *cf = tau + MAXB, *Aug = cf + CMAX;
    double *tau_contact = w.tau_contact + s * m;
    int st = w.status[s] & ~4;
    const int c = (st & 1) ? w.cdof[s] : 0;
    if (c != 12 || cfg->ncontact < 2 || RMAX < 6) {
        for (int i = lane; i < m; i += 32) tau_contact[i] = 0.0;
        if (lane == 0) w.status[s] = st;
        return;
    }
    for (int i = lane; i < m; i += 32) tau[i] = w.tau_grav[s * m + i] + w.tau_task[s * m + i] + tau_contact[i];
    __syncwarp();
    warp_contact_force(w, s, c, tau, cf);
    __syncwarp();
    // every lane evaluates the closed form redundantly (no divergence, no broadcast needed afterwards)
    const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
    const double *R0 = pose;  // pelvis, row-major world <- link
    // link_[0].rpy(2) with rpy = rotm.eulerAngles(2, 1, 0) (reference src/link.cpp:81; Eigen's documented algorithm)
    double e0 = atan2(R0[3], R0[0]);
    if (e0 < 0.0) e0 += 3.14159265358979323846;
    const double s1 = sin(e0), c1 = cos(e0);
    const double e2 = atan2(s1 * R0[2] - c1 * R0[5], c1 * R0[4] - s1 * R0[1]);
    const double ang = -e2, ca = cos(ang), sa = sin(ang);
    const double Rz[9] = {ca, -sa, 0, sa, ca, 0, 0, 0, 1};  // rotateWithZ, reference src/math.cpp:55-74
    const Vec3 com = ld3(w.com + s * 16);
    Vec3 P[2];
    for (int k = 0; k < 2; k++) {
        const ContactCfg &cc = cfg->contacts[k];
        const double *po = pose + cc.link * POSE_STRIDE;
        P[k] = mv(Rz, ld3(po + 9) + mv(po, v3(cc.point[0], cc.point[1], cc.point[2])) - com);
    }
    double F12[12], FR[12];
    for (int k = 0; k < 4; k++) st3(F12 + 3 * k, mv(Rz, ld3(cf + 3 * k)));
    redistribute_two(0.99, 0.26, 0.1, 1.0, 0.9, 0.9, P[0], P[1], F12, FR);
    const Vec3 d0 = mtv(Rz, ld3(FR + 6)) - ld3(cf + 6), d1 = mtv(Rz, ld3(FR + 9)) - ld3(cf + 9);
    // Aug = [J_C_INV_T[6:12, 6:] V2^T | desired]   (6 x 7)
    const double *Jb = w.Jbar + s * (long long)CMAX * n + 6 * CMAX, *V2 = w.V2 + s * (long long)RMAX * m;
    for (int e = lane; e < 36; e += 32) {
        const int j = e / 6, i = e - j * 6;
        double a0 = 0.0;
        for (int a = 0; a < m; a++) a0 = fma(Jb[6 + i + a * CMAX], V2[j + a * RMAX], a0);
        Aug[i + j * lda] = a0;
    }
    if (lane == 0) {
        st3(Aug + 6 * lda, d0);
        st3(Aug + 6 * lda + 3, d1);
    }
    __syncwarp();
    const double cond = wgauss_jordan(6, 1, Aug, lda);
    __syncwarp();
    for (int a = lane; a < m; a += 32) {
        double a0 = 0.0;
        for (int j = 0; j < 6; j++) a0 = fma(V2[j + a * RMAX], Aug[j + 6 * lda], a0);
        tau_contact[a] = a0;
    }
    if (lane == 0) w.status[s] = (cond > 1e-14) ? (st | 4) : st;
}

}  // namespace dwbc
