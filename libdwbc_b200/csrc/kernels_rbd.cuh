// Stage a1/a2 of the hot path: rigid-body dynamics and contact Jacobians, one warp per robot state.
//
// Computes what RobotData::UpdateKinematics (reference src/dwbc.cpp:279-371) gets from RBDL
// (UpdateKinematicsCustom, CompositeRigidBodyAlgorithm, NonlinearEffects) and what
// ContactConstraint::Update / Link::UpdateJac (reference src/contact_constraint.cpp:51-77,
// src/link.cpp:76-107) get from CalcPointJacobian6D -- but formulated for a warp:
//
//   * forward kinematics sweeps the tree LEVEL by level (all bodies of one depth in parallel);
//   * every spatial quantity is expressed in ONE inertial frame: world axes, origin at the pelvis
//     position p0 of this state.  In a common frame composite inertias are plain sums over subtrees
//     (mass, first moment, inertia about p0: 10 numbers per body), joint motion vectors need no
//     transforms, and the mass matrix is   A(i,j) = s_j . (Ic_i s_i)   for j on the path of i.
//     Anchoring the frame at p0 (not at the world origin) keeps the parallel-axis terms small,
//     so nothing is lost to cancellation when the robot is far from the world origin;
//   * the centre of mass, the gravity vector G = 9.81 * A(2,:) and the centroidal quantities follow
//     from A and the whole-body composite inertia, no extra pass.
#pragma once
#include "model.h"
#include "warp_linalg.cuh"

namespace dwbc {

// per-body record kept in shared memory (doubles)
enum { RB_R = 0, RB_P = 9, RB_AX = 12, RB_IC = 15 /* M, h(3), Ixx,Ixy,Ixz,Iyy,Iyz,Izz */, RB_V = 25, RB_A = 31, RB_F = 37, RB_STRIDE = 43 };
// per-body pose record written to global memory for later stages: R (row-major 9), p (3), world axis (3)
constexpr int POSE_STRIDE = 15;

struct Vec3 {
    double x, y, z;
};
__device__ __forceinline__ Vec3 v3(double x, double y, double z) { return Vec3{x, y, z}; }
__device__ __forceinline__ Vec3 operator+(Vec3 a, Vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ Vec3 operator-(Vec3 a, Vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ Vec3 operator*(double s, Vec3 a) { return v3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ Vec3 cross(Vec3 a, Vec3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ Vec3 ld3(const double *p) { return v3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(double *p, Vec3 a) { p[0] = a.x, p[1] = a.y, p[2] = a.z; }
// row-major 3x3 times vector
__device__ __forceinline__ Vec3 mv(const double *R, Vec3 a) {
    return v3(R[0] * a.x + R[1] * a.y + R[2] * a.z, R[3] * a.x + R[4] * a.y + R[5] * a.z, R[6] * a.x + R[7] * a.y + R[8] * a.z);
}
__device__ __forceinline__ Vec3 mtv(const double *R, Vec3 a) {
    return v3(R[0] * a.x + R[3] * a.y + R[6] * a.z, R[1] * a.x + R[4] * a.y + R[7] * a.z, R[2] * a.x + R[5] * a.y + R[8] * a.z);
}
__device__ __forceinline__ void mm3(const double *A, const double *B, double *C) {
    DWBC_UNROLL(3)
    for (int i = 0; i < 3; i++) {
        DWBC_UNROLL(3)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
    }
}
// symmetric 3x3 stored as xx,xy,xz,yy,yz,zz times vector
__device__ __forceinline__ Vec3 symv(const double *S, Vec3 a) {
    return v3(S[0] * a.x + S[1] * a.y + S[2] * a.z, S[1] * a.x + S[3] * a.y + S[4] * a.z, S[2] * a.x + S[4] * a.y + S[5] * a.z);
}

// spatial (common frame) inertia times motion vector [w; v]: returns moment n and force f
__device__ __forceinline__ void inertia_apply(const double *ic, Vec3 w, Vec3 v, Vec3 &n, Vec3 &f) {
    const double M = ic[0];
    const Vec3 h = ld3(ic + 1);
    n = symv(ic + 4, w) + cross(h, v);
    f = M * v - cross(h, w);
}

// Motion vector of velocity column `col` (0..n-1) of a state, in the common frame: angular part w,
// linear part v (velocity of the point instantaneously at p0).
__device__ __forceinline__ void motion_col(const double *rb, int col, Vec3 &w, Vec3 &v) {
    if (col < 3) {
        w = v3(0, 0, 0);
        v = v3(col == 0, col == 1, col == 2);
    } else if (col < 6) {
        const double *R0 = rb + RB_R;
        w = v3(R0[col - 3], R0[3 + col - 3], R0[6 + col - 3]);  // column (col-3) of R0 = R0 * e_k
        v = v3(0, 0, 0);
    } else {
        const double *b = rb + (col - 5) * RB_STRIDE;
        w = ld3(b + RB_AX);
        v = cross(ld3(b + RB_P), w);  // (p_j - p0) x a_j, RB_P is stored relative to p0
    }
}

struct RbdOut {
    double *A;        // [B][n*n]
    double *G;        // [B][n]
    double *Bnl;      // [B][n]   nonlinear effects (may be null)
    double *pose;     // [B][nb][POSE_STRIDE]  (p is absolute world position)
    double *com;      // [B][16]: com(3), com-from-pelvis(3), Icom world (9), pad
    double *linkvel;  // [B][nb][6]  world [v; w] of every link origin (may be null)
};

// shared-memory doubles needed per warp
__host__ __device__ inline int rbd_smem_doubles(int nb) { return nb * RB_STRIDE + 16; }

// Forward / backward sweeps of one state by ONE warp: pose, world axis, velocity, bias acceleration, common-frame inertia,
// composite inertia and subtree bias force of every body, left in the per-body records rb (RB_*).  Only __syncwarp inside.
__device__ __forceinline__ void rbd_sweeps(const DevModel *__restrict__ md, const double *__restrict__ q, const double *__restrict__ qd, double *rb) {
    const int lane = lane_id();
    const int nb = md->nb, n = md->n;
    // ---- base pose ----
    if (lane == 0) {
        const double x = q[3], y = q[4], z = q[5], w = q[n];
        double *R = rb + RB_R;
        R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
        R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
        R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
        rb[RB_P] = rb[RB_P + 1] = rb[RB_P + 2] = 0.0;  // positions are kept relative to p0
        rb[RB_AX] = rb[RB_AX + 1] = rb[RB_AX + 2] = 0.0;
        // spatial velocity of the base: w = R0 * omega_body, v = pdot
        Vec3 wb = mv(R, v3(qd[3], qd[4], qd[5]));
        st3(rb + RB_V, wb);
        st3(rb + RB_V + 3, v3(qd[0], qd[1], qd[2]));
        // RNEA start: gravity as an upward acceleration of the frame, plus v0 x (S_rot omega) = [0; pdot x w]
        st3(rb + RB_A, v3(0, 0, 0));
        st3(rb + RB_A + 3, cross(v3(qd[0], qd[1], qd[2]), wb) + v3(0, 0, 9.81));
    }
    __syncwarp();
    // ---- forward sweep by depth: pose, world axis, velocity, acceleration (qddot = 0) ----
    for (int d = 1; d <= md->max_depth; d++) {
        const int e0 = md->level_start[d], e1 = md->level_start[d + 1];
        for (int e = e0 + lane; e < e1; e += 32) {
            const int b = md->order[e], l = md->parent[b];
            const double *pb = rb + l * RB_STRIDE;
            double *bb = rb + b * RB_STRIDE;
            // joint rotation about the (unit) axis: Rodrigues
            const Vec3 a = ld3(md->axis[b]);
            double sn, cs;
            sincos(q[5 + b], &sn, &cs);
            const double c1 = 1.0 - cs;
            double Rq[9] = {cs + c1 * a.x * a.x,       c1 * a.x * a.y - sn * a.z, c1 * a.x * a.z + sn * a.y,
                            c1 * a.x * a.y + sn * a.z, cs + c1 * a.y * a.y,       c1 * a.y * a.z - sn * a.x,
                            c1 * a.x * a.z - sn * a.y, c1 * a.y * a.z + sn * a.x, cs + c1 * a.z * a.z};
            double Rj[9], Rw[9];
            mm3(md->E_T[b], Rq, Rj);
            mm3(pb + RB_R, Rj, Rw);
            DWBC_UNROLL(9)
            for (int k = 0; k < 9; k++) bb[RB_R + k] = Rw[k];
            const Vec3 p = ld3(pb + RB_P) + mv(pb + RB_R, ld3(md->r_T[b]));
            st3(bb + RB_P, p);
            const Vec3 aw = mv(Rw, a);
            st3(bb + RB_AX, aw);
            // v_b = v_parent + s_b qd_b ; a_b = a_parent + (v_b x s_b) qd_b
            const double qdb = qd[5 + b];
            const Vec3 sw = aw, sv = cross(p, aw);
            const Vec3 vw = ld3(pb + RB_V) + qdb * sw, vv = ld3(pb + RB_V + 3) + qdb * sv;
            st3(bb + RB_V, vw);
            st3(bb + RB_V + 3, vv);
            const Vec3 cw = cross(vw, sw), cv = cross(vw, sv) + cross(vv, sw);
            st3(bb + RB_A, ld3(pb + RB_A) + qdb * cw);
            st3(bb + RB_A + 3, ld3(pb + RB_A + 3) + qdb * cv);
        }
        __syncwarp();
    }
    // ---- per-body inertia in the common frame, and the RNEA body force ----
    for (int b = lane; b < nb; b += 32) {
        double *bb = rb + b * RB_STRIDE;
        const double *R = bb + RB_R;
        const double mass = md->mass[b];
        const Vec3 c = ld3(bb + RB_P) + mv(R, ld3(md->com[b]));  // COM relative to p0
        double T[9], Iw[9], Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
        mm3(R, md->inertia[b], T);
        mm3(T, Rt, Iw);
        const double cc = dot(c, c);
        double *ic = bb + RB_IC;
        ic[0] = mass;
        st3(ic + 1, mass * c);
        ic[4] = Iw[0] + mass * (cc - c.x * c.x);
        ic[5] = Iw[1] - mass * c.x * c.y;
        ic[6] = Iw[2] - mass * c.x * c.z;
        ic[7] = Iw[4] + mass * (cc - c.y * c.y);
        ic[8] = Iw[5] - mass * c.y * c.z;
        ic[9] = Iw[8] + mass * (cc - c.z * c.z);
        // f = I a + v x* (I v)
        Vec3 na, fa, nv, fv;
        inertia_apply(ic, ld3(bb + RB_A), ld3(bb + RB_A + 3), na, fa);
        const Vec3 vw = ld3(bb + RB_V), vv = ld3(bb + RB_V + 3);
        inertia_apply(ic, vw, vv, nv, fv);
        st3(bb + RB_F, na + cross(vw, nv) + cross(vv, fv));
        st3(bb + RB_F + 3, fa + cross(vw, fv));
    }
    __syncwarp();
    // ---- backward sweep by depth: composite inertia and subtree force (parents gather children) ----
    for (int d = md->max_depth - 1; d >= 0; d--) {
        const int e0 = md->level_start[d], e1 = md->level_start[d + 1];
        for (int e = e0 + lane; e < e1; e += 32) {
            const int b = md->order[e];
            double *bb = rb + b * RB_STRIDE;
            for (int ci = md->child_start[b]; ci < md->child_start[b + 1]; ci++) {
                const double *cb = rb + md->child_list[ci] * RB_STRIDE;
                DWBC_UNROLL(10)
                for (int k = 0; k < 10; k++) bb[RB_IC + k] += cb[RB_IC + k];
                DWBC_UNROLL(6)
                for (int k = 0; k < 6; k++) bb[RB_F + k] += cb[RB_F + k];
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void rbd_state(const DevModel *__restrict__ md, const double *__restrict__ q, const double *__restrict__ qd,
                                          double *rb, const RbdOut &out, long long s) {
    const int lane = lane_id();
    const int nb = md->nb, n = md->n;
    rbd_sweeps(md, q, qd, rb);
    // ---- mass matrix: row i per lane, all columns; A(i,j) = s_j . (Ic_deeper s_deeper) ----
    double *A = out.A + s * (long long)n * n;
    for (int i = lane; i < n; i += 32) {
        Vec3 wi, vi;
        motion_col(rb, i, wi, vi);
        const int bi = i < 6 ? 0 : i - 5;
        Vec3 ni, fi;
        inertia_apply(rb + bi * RB_STRIDE + RB_IC, wi, vi, ni, fi);
        const uint64_t anc_i = md->anc_mask[bi];
        for (int j = 0; j < n; j++) {
            const int bj = j < 6 ? 0 : j - 5;
            double val = 0.0;
            Vec3 wj, vj;
            motion_col(rb, j, wj, vj);
            if (bj == bi || bj == 0 || ((anc_i >> bj) & 1ull)) {
                // body of column j is the body of column i or one of its ancestors: the deeper body is i's
                val = dot(wj, ni) + dot(vj, fi);
            } else if (bi == 0 || ((md->anc_mask[bj] >> bi) & 1ull)) {
                // i's body is a proper ancestor of j's: the deeper body is j's
                Vec3 nj, fj;
                inertia_apply(rb + bj * RB_STRIDE + RB_IC, wj, vj, nj, fj);
                val = dot(wi, nj) + dot(vi, fj);
            }
            A[i + (long long)j * n] = val;
        }
    }
    // ---- gravity vector, nonlinear effects, COM quantities ----
    const double Mtot = rb[RB_IC];
    const Vec3 cfp = (1.0 / Mtot) * ld3(rb + RB_IC + 1);  // COM relative to the pelvis origin, world axes
    for (int j = lane; j < n; j += 32) {
        Vec3 wj, vj;
        motion_col(rb, j, wj, vj);
        // A(2,j) = e_z-translation column against column j = f-part of Ic_base-or-j applied ... = M * (v_com of column j).z
        const int bj = j < 6 ? 0 : j - 5;
        Vec3 nj, fj;
        inertia_apply(rb + bj * RB_STRIDE + RB_IC, wj, vj, nj, fj);
        out.G[s * n + j] = 9.81 * fj.z;  // reference src/dwbc.cpp:358 (G_ = -Jcom_lin^T m g), identity G = 9.81 A(2,:)
        if (out.Bnl) {
            const double *F = rb + bj * RB_STRIDE + RB_F;
            out.Bnl[s * n + j] = dot(wj, ld3(F)) + dot(vj, ld3(F + 3));
        }
    }
    if (lane == 0) {
        double *co = out.com + s * 16;
        st3(co, cfp + v3(q[0], q[1], q[2]));
        st3(co + 3, cfp);
        // inertia about the COM in world axes (reference src/dwbc.cpp:346): Io - M (|c|^2 1 - c c^T)
        const double *io = rb + RB_IC + 4;
        const double cc = dot(cfp, cfp);
        co[6] = io[0] - Mtot * (cc - cfp.x * cfp.x);
        co[7] = co[9] = io[1] + Mtot * cfp.x * cfp.y;
        co[8] = co[12] = io[2] + Mtot * cfp.x * cfp.z;
        co[10] = io[3] - Mtot * (cc - cfp.y * cfp.y);
        co[11] = co[13] = io[4] + Mtot * cfp.y * cfp.z;
        co[14] = io[5] - Mtot * (cc - cfp.z * cfp.z);
        co[15] = Mtot;
    }
    // ---- poses for the Jacobian consumers ----
    const Vec3 p0 = v3(q[0], q[1], q[2]);
    for (int b = lane; b < nb; b += 32) {
        const double *bb = rb + b * RB_STRIDE;
        double *po = out.pose + (s * nb + b) * POSE_STRIDE;
        DWBC_UNROLL(9)
        for (int k = 0; k < 9; k++) po[k] = bb[RB_R + k];
        st3(po + 9, ld3(bb + RB_P) + p0);
        st3(po + 12, ld3(bb + RB_AX));
        if (out.linkvel) {
            // velocity of the link origin: v(p0) + w x (p_b - p0)   (reference src/link.cpp:85-88)
            const Vec3 w = ld3(bb + RB_V);
            double *lv = out.linkvel + (s * nb + b) * 6;
            st3(lv, ld3(bb + RB_V + 3) + cross(w, ld3(bb + RB_P)));
            st3(lv + 3, w);
        }
    }
    __syncwarp();
}

// 6 x n point Jacobian (rows [linear; angular], reference src/contact_constraint.cpp:56-62) of a
// point fixed in body `body` at local offset `pl`, one column per thread (`tid` of `nt` cooperating threads),
// written to dst with leading dimension ldd starting at row `row0`.  `pose` = this state's pose records.
__device__ __forceinline__ void point_jacobian_t(const DevModel *__restrict__ md, const double *__restrict__ pose, int body, Vec3 pl,
                                                 double *dst, int ldd, int row0, int rows_first, int rows_count, int tid, int nt) {
    const int n = md->n;
    const double *pb = pose + body * POSE_STRIDE;
    const Vec3 pw = ld3(pb + 9) + mv(pb, pl);
    const double *R0 = pose;
    const Vec3 p0 = ld3(pose + 9);
    const uint64_t anc = md->anc_mask[body];
    for (int j = tid; j < n; j += nt) {
        Vec3 lin = v3(0, 0, 0), ang = v3(0, 0, 0);
        if (j < 3) {
            lin = v3(j == 0, j == 1, j == 2);
        } else if (j < 6) {
            const Vec3 a = v3(R0[j - 3], R0[3 + j - 3], R0[6 + j - 3]);
            ang = a;
            lin = cross(a, pw - p0);
        } else if ((anc >> (j - 5)) & 1ull) {
            const double *pj = pose + (j - 5) * POSE_STRIDE;
            const Vec3 a = ld3(pj + 12);
            ang = a;
            lin = cross(a, pw - ld3(pj + 9));
        }
        const double v[6] = {lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
        for (int r = 0; r < rows_count; r++) dst[row0 + r + j * ldd] = v[rows_first + r];
    }
}
__device__ __forceinline__ void point_jacobian(const DevModel *__restrict__ md, const double *__restrict__ pose, int body, Vec3 pl,
                                               double *dst, int ldd, int row0, int rows_first, int rows_count) {
    point_jacobian_t(md, pose, body, pl, dst, ldd, row0, rows_first, rows_count, lane_id(), 32);
}

}  // namespace dwbc
