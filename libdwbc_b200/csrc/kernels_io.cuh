// The steps either side of the hot path, as batched element-wise kernels:
//   * f* generation: task PD on quintic / slerp trajectories -- TaskLink::GetFstarPosPD / GetFstarRotPD
//     (reference src/task.cpp:268-334), QuinticSpline, GetPhi (src/math.cpp:127-183, 272-287), hooked into
//     UpdateTaskSpace per LINK_MODE (src/dwbc.cpp:714-785);
//   * the wire format of the RL caller: MuJoCo qpos/qvel layout -> RBDL layout (src/pybind/rl_wbc_bridge.cpp:88-107) and
//     the float32 torque command torque_grav_ + torque_task_ + torque_contact_ (getTorqueCommand, :131-139).
// HBM-bound byte shuffling: one thread per element / per (state, task link), coalesced state-major rows.
#pragma once
#include "batch.h"
#include "kernels_rbd.cuh"

namespace dwbc {

// ---- MuJoCo <-> RBDL state layout ----
// qpos = [x y z | qw qx qy qz | joints]  ->  q = [x y z | qx qy qz | joints | qw];  qvel is copied.
__global__ void mujoco_to_rbdl_kernel(int B, int n, const double *__restrict__ qpos, const double *__restrict__ qvel, double *__restrict__ q,
                                      double *__restrict__ qd) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int np = n + 1;
    if (e < (long long)B * np) {
        const long long s = e / np;
        const int j = (int)(e - s * np);
        int src;
        if (j < 3) src = j;
        else if (j < 6) src = j + 1;
        else if (j < n) src = j + 1;
        else src = 3;
        q[e] = qpos[s * np + src];
    }
    if (e < (long long)B * n) qd[e] = qvel[e];
}

// ctrl[b][i] = float(torque_grav_ + torque_task_ + torque_contact_)
__global__ void torque_command_f32_kernel(long long count, const double *__restrict__ tg, const double *__restrict__ tt, const double *__restrict__ tc,
                                          float *__restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < count) out[e] = (float)(tg[e] + tt[e] + tc[e]);
}

// ---- f* from trajectories ----
struct TrajLink {           // one task link of the level
    int mode, link, row, dof;   // LINK_MODE, link id (nb = COM pseudo link), first row in the level's f*, rows
    double point[3];
    double gains[18];       // pos_p, pos_d, pos_a, rot_p, rot_d, rot_a  (SetTaskGain, reference src/task.cpp:253-264)
    // per-state trajectory parameters (device, caller-owned); null = trajectory not set (f* rows left untouched)
    const double *pos_time, *pos_init, *vel_init, *pos_des, *vel_des;  // [B][2], [B][3] ...
    const double *rot_time, *rot_init, *rot_des;                       // [B][2], [B][9] row-major
};
struct TrajLevel {
    int nlinks;
    TrajLink links[MAX_TASK_LINKS];
};

// QuinticSpline (reference src/math.cpp:127-183) with the 3x3 solve written out: returns (position, velocity, acceleration)
__device__ __forceinline__ Vec3 quintic_spline(double time, double t0, double tf, double x0, double v0, double a0, double xf, double vf, double af) {
    if (time < t0) return v3(x0, v0, a0);
    if (time > tf) return v3(xf, vf, af);
    const double T = tf - t0;
    const double R1 = xf - x0 - v0 * T - a0 * T * T / 2.0, R2 = vf - v0 - a0 * T, R3 = af - a0;
    const double T2 = T * T, T3 = T2 * T;
    const double a4 = (10.0 * R1 - 4.0 * R2 * T + 0.5 * R3 * T2) / T3;
    const double a5 = (-15.0 * R1 + 7.0 * R2 * T - R3 * T2) / (T3 * T);
    const double a6 = (6.0 * R1 - 3.0 * R2 * T + 0.5 * R3 * T2) / (T3 * T2);
    const double a3 = a0 / 2.0, s = time - t0, s2 = s * s, s3 = s2 * s;
    return v3(x0 + v0 * s + a3 * s2 + a4 * s3 + a5 * s3 * s + a6 * s3 * s2,
              v0 + 2.0 * a3 * s + 3.0 * a4 * s2 + 4.0 * a5 * s3 + 5.0 * a6 * s3 * s,
              2.0 * a3 + 6.0 * a4 * s + 12.0 * a5 * s2 + 20.0 * a6 * s3);
}

struct Quat {
    double w, x, y, z;
};
// Eigen::Quaterniond(Matrix3d) (Eigen/src/Geometry/Quaternion.h, quaternionbase_assign_impl): R row-major
__device__ __forceinline__ Quat quat_from_rot(const double *R) {
    Quat q;
    double t = R[0] + R[4] + R[8];
    if (t > 0.0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[7] - R[5]) * t;
        q.y = (R[2] - R[6]) * t;
        q.z = (R[3] - R[1]) * t;
    } else {
        int i = 0;
        if (R[4] > R[0]) i = 1;
        if (R[8] > R[4 * i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[3 * k + j] - R[3 * j + k]) * t;
        v[j] = (R[3 * j + i] + R[3 * i + j]) * t;
        v[k] = (R[3 * k + i] + R[3 * i + k]) * t;
        q.x = v[0], q.y = v[1], q.z = v[2];
    }
    return q;
}
__device__ __forceinline__ Quat quat_mul(Quat a, Quat b) {
    return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
// Eigen QuaternionBase::slerp
__device__ __forceinline__ Quat quat_slerp(Quat a, double t, Quat b) {
    const double one = 1.0 - 2.220446049250313e-16;
    const double d = a.w * b.w + a.x * b.x + a.y * b.y + a.z * b.z, ad = fabs(d);
    double s0, s1;
    if (ad >= one) {
        s0 = 1.0 - t;
        s1 = t;
    } else {
        const double th = acos(ad), st = sin(th);
        s0 = sin((1.0 - t) * th) / st;
        s1 = sin(t * th) / st;
    }
    if (d < 0.0) s1 = -s1;
    return Quat{s0 * a.w + s1 * b.w, s0 * a.x + s1 * b.x, s0 * a.y + s1 * b.y, s0 * a.z + s1 * b.z};
}
// Eigen QuaternionBase::toRotationMatrix, row-major out
__device__ __forceinline__ void quat_to_rot(Quat q, double *R) {
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y, tyz = tz * q.y,
                 tzz = tz * q.z;
    R[0] = 1.0 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1.0 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1.0 - (txx + tyy);
}

// One thread per (state, task link).  fstar [B][t] is updated in place on the rows of links whose trajectory is set.
__global__ void fstar_pd_kernel(const DevModel *__restrict__ md, Work w, const double *__restrict__ qd, TrajLevel lv, int t, const double *__restrict__ time,
                                double *__restrict__ fstar) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)w.B * lv.nlinks) return;
    const long long s = e / lv.nlinks;
    const TrajLink &tl = lv.links[e - s * lv.nlinks];
    if (!tl.pos_time && !tl.rot_time) return;
    const int nb = w.nb, n = w.n;
    const double now = time[s];
    // current pose / velocity of the controlled frame (reference src/dwbc.cpp:714-785, src/link.cpp:76-96)
    Vec3 xpos, v, wv;
    double Rm[9];
    if (tl.link == nb) {
        // COM pseudo link: xpos = com_pos, rotm = pelvis rotm, [v; w] = jac_ qdot  (reference src/dwbc.cpp:320-365)
        const double *co = w.com + s * 16;
        const double *A = w.A + s * (long long)n * n;
        const double *R0 = w.pose + s * (long long)nb * POSE_STRIDE;
        const double *qv = qd + s * n;
        const Vec3 c = ld3(co + 3);
        Vec3 lin = v3(0, 0, 0), h = v3(0, 0, 0);
        for (int j = 0; j < n; j++) {
            const Vec3 al = v3(A[0 + j * n], A[1 + j * n], A[2 + j * n]), aa = v3(A[3 + j * n], A[4 + j * n], A[5 + j * n]);
            lin = lin + qv[j] * al;
            h = h + qv[j] * (cross(al, c) + mv(R0, aa));
        }
        const double *I = co + 6;
        const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[5] * I[6] - I[3] * I[8], c02 = I[3] * I[7] - I[4] * I[6];
        const double rdet = 1.0 / (I[0] * c00 + I[1] * c01 + I[2] * c02);
        const double c11 = I[0] * I[8] - I[2] * I[6], c12 = I[1] * I[6] - I[0] * I[7], c22 = I[0] * I[4] - I[1] * I[3];
        const double inv[9] = {c00 * rdet, c01 * rdet, c02 * rdet, c01 * rdet, c11 * rdet, c12 * rdet, c02 * rdet, c12 * rdet, c22 * rdet};
        xpos = ld3(co);
        v = (1.0 / co[15]) * lin;
        wv = mv(inv, h);
        for (int k = 0; k < 9; k++) Rm[k] = R0[k];
    } else {
        const double *po = w.pose + (s * nb + tl.link) * POSE_STRIDE;
        const double *lvl = w.linkvel + (s * nb + tl.link) * 6;
        for (int k = 0; k < 9; k++) Rm[k] = po[k];
        xpos = ld3(po + 9);
        v = ld3(lvl);
        wv = ld3(lvl + 3);
        if (tl.mode == DWBC_TASK_LINK_6D_COM_FRAME || tl.mode == DWBC_TASK_LINK_POSITION_COM_FRAME) {
            const Vec3 r = mv(Rm, ld3(md->com[tl.link]));  // xipos = xpos + rotm com, vi = v + w x (rotm com)
            xpos = xpos + r;
            v = v + cross(wv, r);
        } else if (tl.mode == DWBC_TASK_LINK_6D_CUSTOM_FRAME || tl.mode == DWBC_TASK_LINK_POSITION_CUSTOM_FRAME) {
            const Vec3 pt = v3(tl.point[0], tl.point[1], tl.point[2]);
            xpos = xpos + mv(Rm, pt);
            v = v + cross(wv, pt);  // reference quirk (src/dwbc.cpp:741, 768): w x task_point_ with the LOCAL point
        }
    }
    double *f = fstar + s * t + tl.row;
    const bool has_pos = tl.mode <= DWBC_TASK_LINK_POSITION_CUSTOM_FRAME;
    const bool has_rot = tl.mode <= DWBC_TASK_LINK_6D_CUSTOM_FRAME || tl.mode >= DWBC_TASK_LINK_ROTATION;
    const int rot_row = tl.mode <= DWBC_TASK_LINK_6D_CUSTOM_FRAME ? 3 : 0;
    if (has_pos && tl.pos_time) {
        const double t0 = tl.pos_time[2 * s], t1 = tl.pos_time[2 * s + 1];
        const double cur[3] = {xpos.x, xpos.y, xpos.z}, cv[3] = {v.x, v.y, v.z};
        for (int j = 0; j < 3; j++) {
            const Vec3 qn = quintic_spline(now, t0, t1, tl.pos_init[3 * s + j], tl.vel_init[3 * s + j], 0.0, tl.pos_des[3 * s + j], tl.vel_des[3 * s + j], 0.0);
            f[j] = tl.gains[6 + j] * qn.z + tl.gains[j] * (qn.x - cur[j]) + tl.gains[3 + j] * (qn.y - cv[j]);
        }
    }
    if (has_rot && tl.rot_time) {
        const double t0 = tl.rot_time[2 * s], t1 = tl.rot_time[2 * s + 1];
        const Vec3 qs = quintic_spline(now, t0, t1, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0);
        const Quat qi = quat_from_rot(tl.rot_init + 9 * s), qdz = quat_from_rot(tl.rot_des + 9 * s);
        double Rt[9];
        quat_to_rot(quat_slerp(qi, qs.x, qdz), Rt);
        // AngleAxisd(rq_desired * rq_init.inverse())   (Eigen AngleAxis::operator=(QuaternionBase))
        const double ni = qi.w * qi.w + qi.x * qi.x + qi.y * qi.y + qi.z * qi.z;
        const Quat qinv{qi.w / ni, -qi.x / ni, -qi.y / ni, -qi.z / ni};
        const Quat dq = quat_mul(qdz, qinv);
        double nrm = sqrt(dq.x * dq.x + dq.y * dq.y + dq.z * dq.z);
        double angle = 0.0;
        Vec3 axis = v3(1, 0, 0);
        if (nrm != 0.0) {
            angle = 2.0 * atan2(nrm, fabs(dq.w));
            if (dq.w < 0.0) nrm = -nrm;
            axis = v3(dq.x / nrm, dq.y / nrm, dq.z / nrm);
        }
        const Vec3 wt = (angle * qs.y) * axis;
        // GetPhi(current_rot, rot_traj) = 1/2 sum_i cur[:, i] x traj[:, i]   (reference src/math.cpp:272-287)
        Vec3 phi = v3(0, 0, 0);
        for (int i = 0; i < 3; i++) phi = phi + cross(v3(Rm[i], Rm[3 + i], Rm[6 + i]), v3(Rt[i], Rt[3 + i], Rt[6 + i]));
        phi = 0.5 * phi;
        const double pe[3] = {phi.x, phi.y, phi.z}, de[3] = {wt.x - wv.x, wt.y - wv.y, wt.z - wv.z};
        for (int j = 0; j < 3; j++) f[rot_row + j] = tl.gains[9 + j] * pe[j] + tl.gains[12 + j] * de[j];
    }
}

}  // namespace dwbc
