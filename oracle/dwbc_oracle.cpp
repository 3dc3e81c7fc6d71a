// TEST INFRASTRUCTURE -- CPU oracle for the batched whole-body-control hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
// load the library built from this file (oracle/build/liboracle.so).  The product
// (libdwbc_b200/csrc) never links, includes or calls anything in oracle/.
//
// This is a plain, single-state, FP64 restatement of the reference's per-control-cycle path,
// following the reference AS WRITTEN (same products, LU inverses, COD pseudo-inverse):
//   UpdateKinematics             reference src/dwbc.cpp:279-371   (+ RBDL FK / CRBA / RNEA restated
//                                from Featherstone, "Rigid Body Dynamics Algorithms", ch. 5-6;
//                                RBDL = saga0619/rbdl-orb fork, unpinned, not available offline)
//   SetContact / UpdateContactConstraint   include/dwbc.h:432-474, src/dwbc.cpp:433-454,
//                                src/contact_constraint.cpp:51-77
//   CalcContactConstraint        src/dwbc.cpp:456-478 -> src/wbd.cpp:108-143, 32-53
//   CalcGravCompensation         src/dwbc.cpp:875-889 -> src/wbd.cpp:186-192
//   SetTaskSpace/UpdateTaskSpace/CalcTaskSpace   src/dwbc.cpp:662-816, src/task.cpp:90-93,202-211,
//                                src/wbd.cpp:207-213,257-261
//   CalcTaskControlTorque / CalcSingleTaskTorqueWithQP   src/dwbc.cpp:818-873, 941-1127
//   CalcContactRedistribute      src/dwbc.cpp:1372-1568
// Parity is pinned by the reference's golden vectors tests/cases/{1,2} (tests/test_oracle_golden.py).
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "linalg.hpp"
#include "qp.hpp"

#include <atomic>
#include <thread>

namespace orc {

struct V3 {
    double x = 0, y = 0, z = 0;
};
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
struct M3 {
    double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
};
static inline M3 eye3() {
    M3 r;
    r.m[0][0] = r.m[1][1] = r.m[2][2] = 1;
    return r;
}
static inline M3 mul(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
static inline M3 tr(const M3 &a) {
    M3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
    return r;
}
static inline V3 mul(const M3 &a, V3 v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z, a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
static inline M3 skew3(V3 v) {
    M3 r;
    r.m[0][1] = -v.z; r.m[0][2] = v.y;
    r.m[1][0] = v.z;  r.m[1][2] = -v.x;
    r.m[2][0] = -v.y; r.m[2][1] = v.x;
    return r;
}
static inline M3 axis_rot(V3 a, double th) {
    double n = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    a = (1.0 / n) * a;
    M3 K = skew3(a), K2 = mul(K, K), R = eye3();
    double s = std::sin(th), c = 1 - std::cos(th);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R.m[i][j] += s * K.m[i][j] + c * K2.m[i][j];
    return R;
}
static inline M3 quat_rot(double x, double y, double z, double w) {
    M3 R;
    R.m[0][0] = 1 - 2 * (y * y + z * z); R.m[0][1] = 2 * (x * y - z * w);     R.m[0][2] = 2 * (x * z + y * w);
    R.m[1][0] = 2 * (x * y + z * w);     R.m[1][1] = 1 - 2 * (x * x + z * z); R.m[1][2] = 2 * (y * z - x * w);
    R.m[2][0] = 2 * (x * z - y * w);     R.m[2][1] = 2 * (y * z + x * w);     R.m[2][2] = 1 - 2 * (x * x + y * y);
    return R;
}

// 6x6 helpers on spatial quantities ordered [angular; linear] (Featherstone).
struct M6 {
    double m[6][6];
    M6() { std::memset(m, 0, sizeof(m)); }
};
struct V6 {
    double v[6] = {0, 0, 0, 0, 0, 0};
};
static inline V6 mul(const M6 &a, const V6 &x) {
    V6 r;
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += a.m[i][j] * x.v[j];
        r.v[i] = s;
    }
    return r;
}
static inline V6 mulT(const M6 &a, const V6 &x) {
    V6 r;
    for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int j = 0; j < 6; j++) s += a.m[j][i] * x.v[j];
        r.v[i] = s;
    }
    return r;
}
static inline M6 mul(const M6 &a, const M6 &b) {
    M6 r;
    for (int i = 0; i < 6; i++)
        for (int k = 0; k < 6; k++) {
            double aik = a.m[i][k];
            if (aik == 0) continue;
            for (int j = 0; j < 6; j++) r.m[i][j] += aik * b.m[k][j];
        }
    return r;
}
static inline M6 tr(const M6 &a) {
    M6 r;
    for (int i = 0; i < 6; i++)
        for (int j = 0; j < 6; j++) r.m[i][j] = a.m[j][i];
    return r;
}
// motion transform parent -> child; R = rotation child->parent, r = child origin in parent coords
static inline M6 xform(const M3 &R, V3 r) {
    M6 X;
    M3 Rt = tr(R), Rr = mul(Rt, skew3(r));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            X.m[i][j] = Rt.m[i][j];
            X.m[3 + i][3 + j] = Rt.m[i][j];
            X.m[3 + i][j] = -Rr.m[i][j];
        }
    return X;
}
static inline M6 spatial_inertia(double mass, V3 c, const M3 &Ic) {
    M6 I;
    M3 cx = skew3(c), cxcx = mul(cx, cx);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            I.m[i][j] = Ic.m[i][j] - mass * cxcx.m[i][j];
            I.m[i][3 + j] = mass * cx.m[i][j];
            I.m[3 + i][j] = -mass * cx.m[i][j];
        }
    for (int i = 0; i < 3; i++) I.m[3 + i][3 + i] = mass;
    return I;
}
static inline V6 crm(const V6 &v, const V6 &u) {  // v x u (motion)
    V3 w{v.v[0], v.v[1], v.v[2]}, vl{v.v[3], v.v[4], v.v[5]}, uw{u.v[0], u.v[1], u.v[2]}, ul{u.v[3], u.v[4], u.v[5]};
    V3 a = cross(w, uw), b = cross(w, ul) + cross(vl, uw);
    V6 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = b.x; r.v[4] = b.y; r.v[5] = b.z;
    return r;
}
static inline V6 crf(const V6 &v, const V6 &f) {  // v x* f (force)
    V3 w{v.v[0], v.v[1], v.v[2]}, vl{v.v[3], v.v[4], v.v[5]}, n{f.v[0], f.v[1], f.v[2]}, fl{f.v[3], f.v[4], f.v[5]};
    V3 a = cross(w, n) + cross(vl, fl), b = cross(w, fl);
    V6 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = b.x; r.v[4] = b.y; r.v[5] = b.z;
    return r;
}

enum { TASK_LINK_6D = 0, TASK_LINK_6D_COM_FRAME, TASK_LINK_6D_CUSTOM_FRAME, TASK_LINK_POSITION, TASK_LINK_POSITION_COM_FRAME,
       TASK_LINK_POSITION_CUSTOM_FRAME, TASK_LINK_ROTATION, TASK_LINK_ROTATION_CUSTOM_FRAME };
static const double INFTY = 1.0e20;  // qpOASES INFTY, confirmed by the golden lbA*mat
static const double COD_THRESHOLD = 1.0e-6;  // reference include/dwbc_wbd.h:10

struct Contact {
    int link = 0;
    V3 point, normal;
    double lx = 0, ly = 0, mu = 0.2, muz = 0.2;  // reference src/contact_constraint.cpp:42
    bool on = false;
    Mat J;    // 6 x n, rows [linear; angular]
    M3 rotm;  // world <- link
    V3 xc;
};
struct TaskLink {
    int mode = 0, link = 0, dof = 0;
    V3 point;
};
struct Task {
    bool custom = false;
    int dof = 0;
    std::vector<TaskLink> links;
    Mat J, Lambda, Jkt, Null, fstar, fstar_qp, contact_qp, torque_h;
    Mat qpA, qpub;  // last assembled QP (parity with the golden A*mat / ubA*mat)
    std::vector<int> active, gray;
    int qp_status = 0, prox_iters = 0;
    // a9 reduced-dimension variant (reference include/dwbc_task.h: noncont_task, reduced_task_, cmm_task, ..._R_ members)
    bool noncont = false, reduced = false, cmm = false;
    int nc_h = 0;
    Mat J_R, Lambda_R, Jkt_R, Null_R, J_NC, torque_nc, force_on_nc, torque_h_R, torque_null_h_R, torque_null_h_nc;
};

struct Robot {
    // model
    int nb = 0, n = 0, m = 0;
    std::vector<int> parent;
    std::vector<M3> E_T, inertia;
    std::vector<V3> r_T, axis, com;
    std::vector<double> mass;
    double total_mass = 0;
    // configuration
    std::vector<Contact> cc;
    std::vector<Task> ts;
    bool tlim_set = false;
    Mat tlim;
    // state
    std::vector<double> q, qd;
    std::vector<M3> Rw;      // world <- body
    std::vector<V3> pw;      // body origin, world
    std::vector<M6> Xl;      // parent -> body motion transform
    std::vector<V6> vel;     // spatial velocity, body coords
    Mat A, A_inv, B, G, CMM, J_com, J_C, Lambda_c, J_C_INV_T, N_C, A_inv_N_C, W, W_inv, V2, NwJw, P_C;
    Mat torque_grav, torque_task, torque_contact, cf_redis_qp, redisA, redisub;
    V3 com_pos;
    int contact_dof = 0, contact_links = 0, contact_ok = 0, redis_status = 0;
    double redis_eta = 0;
    std::vector<int> redis_active, redis_gray;

    void update_kinematics(const double *q_, const double *qd_);
    Mat point_jac(int body, V3 pl) const;
    void set_contact(const unsigned char *flags);
    int calc_contact_constraint();
    void calc_grav();
    void update_task_space();
    void calc_task_space();
    int calc_task_torque(bool hqp);
    int single_task_qp(Task &t, const Mat &null_prev, const Mat &torque_prev);
    int single_task_qp_on(Task &t, const Mat &Jkt, int m, const Mat &null_prev, const Mat &torque_prev, const Mat &NwJw, const Mat &J_C_INV_T,
                          const Mat &P_C, bool tlim_set, const Mat &tlim);
    // reduced torque limit (reduced_model_dof_ entries): the reference compares torque_limit_.size() with reduced_model_dof_
    // in CalcSingleTaskTorqueWithQP_R (src/dwbc.cpp:3462-3468), i.e. the caller sets a reduced-size limit before the reduced HQP
    bool tlimR_set = false;
    Mat tlimR, torque_contact_R, cf_redis_qp_R, force_on_nc_r, force_on_nc_R_qp, nc_qp_contact, nc_qpA, nc_qpub;
    std::vector<int> nc_qp_active;
    int reduced_redistribute(bool hqp = true);
    Mat redistribute_analytic(const Mat &JCINVT, const Mat &PC, const Mat &V2m, const Mat &tin, int msize);
    int calc_redistribute(bool hqp);
    Mat get_contact_force(const Mat &command_torque) const;
    V3 get_zmp(const Mat &contact_force, std::vector<V3> *per_contact = nullptr) const;
    void contact_const(Mat &A_const, Mat &A_rot, int &csize) const;

    // ---- a9: reduced-dimension variant (Ahn 2024), reference src/dwbc.cpp:2752-3446 ----
    std::vector<int> co_link, nc_link, co_joint, nc_joint, vc_joint;
    int co_dof = 0, nc_dof = 0, vc_dof = 0, rsys = 0, rmod = 0, reduced_ok = 0;
    double mass_nc = 0;
    V3 com_pos_nc;
    Mat A_NC, SI_nc_b, SI_nc_l, cmm_nc, J_I_nc, J_R, A_R_inv, A_R, J_I_nc_inv_T, N_I_nc, G_R, J_CR;
    Mat Lambda_CR, J_CR_INV_T, N_CR, A_R_inv_N_CR, W_R, W_R_inv, V2_R, NwJw_R, P_CR, torque_grav_R;
    Mat J_base_R, J_base_R_kt, lambda_base_R, torque_task_R, torque_task_NC;
    void reduced_dynamics();
    int reduced_contact_constraint();
    void reduced_grav();
    int reduced_task_space();
    int reduced_task_torque(bool hqp);

    // ---- a10: LQP variant (Ahn 2025), reference src/dwbc.cpp:4304-4502, src/dwbc_hqp.cpp ----
    struct LqpLevel {
        int nineq = 0, neq = 0, nu = 0;  // own inequality rows, equality rows, null-space size after this level
        Mat A, a, B, b;                  // A y + a <= v,  B y + b = w   (rows normalised)
        bool cost = false;
        Mat Z, y, v, w, u;               // Z_: nvar x nu; y_ans_, v_ans_, w_ans_, u_ans_
        Mat qpH, qpg, qpA, qpub;         // the level QP exactly as the reference assembles it
        int status = 0;
    };
    std::vector<LqpLevel> lqp;
    Mat lqp_H, lqp_torque;  // cost Hessian shared by the levels; torque of the final y
    int lqp_acc = 0, lqp_con = 0;
    void configure_lqp(bool reduced = false);
    bool lqp_reduced = false;
    void lqp_prepare();
    int lqp_solve_level(int i);
    int calc_lqp();
    // ConfigureLQP_R_NC + CalcControlTorqueLQP_R_NC (reference src/dwbc.cpp:4634-4760, 4479-4502): the non-contact chain's own hierarchy
    std::vector<LqpLevel> lqp_nc;
    Mat lqp_nc_H, lqp_nc_torque, lqp_stitched;  // cost; A_NC[nc,nc] qdd_nc + G_NC; the 33-joint torque of jacc_compare.cpp:543-569
    void configure_lqp_r_nc();
    int calc_lqp_r_nc();
};

void Robot::update_kinematics(const double *q_, const double *qd_) {
    q.assign(q_, q_ + n + 1);
    qd.assign(qd_, qd_ + n);
    Rw.assign(nb, M3());
    pw.assign(nb, V3());
    Xl.assign(nb, M6());
    vel.assign(nb, V6());
    std::vector<M6> Ic(nb);
    // --- FK (RBDL UpdateKinematicsCustom), reference src/dwbc.cpp:304 ---
    Rw[0] = quat_rot(q[3], q[4], q[5], q[n]);
    pw[0] = {q[0], q[1], q[2]};
    Xl[0] = xform(Rw[0], V3());
    {
        V3 vb = mul(tr(Rw[0]), V3{qd[0], qd[1], qd[2]});
        vel[0].v[0] = qd[3]; vel[0].v[1] = qd[4]; vel[0].v[2] = qd[5];
        vel[0].v[3] = vb.x; vel[0].v[4] = vb.y; vel[0].v[5] = vb.z;
    }
    for (int i = 1; i < nb; i++) {
        int l = parent[i];
        M3 Rj = mul(E_T[i], axis_rot(axis[i], q[5 + i]));
        Rw[i] = mul(Rw[l], Rj);
        pw[i] = pw[l] + mul(Rw[l], r_T[i]);
        Xl[i] = xform(Rj, r_T[i]);
        V6 vp = mul(Xl[i], vel[l]);
        vp.v[0] += axis[i].x * qd[5 + i]; vp.v[1] += axis[i].y * qd[5 + i]; vp.v[2] += axis[i].z * qd[5 + i];
        vel[i] = vp;
    }
    // --- CRBA (RBDL CompositeRigidBodyAlgorithm), reference src/dwbc.cpp:305 ---
    for (int i = 0; i < nb; i++) Ic[i] = spatial_inertia(mass[i], com[i], inertia[i]);
    for (int i = nb - 1; i >= 1; i--) {
        M6 t = mul(tr(Xl[i]), mul(Ic[i], Xl[i]));
        M6 &P = Ic[parent[i]];
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) P.m[a][b] += t.m[a][b];
    }
    A.set_zero(n, n);
    M6 S0;  // base motion subspace in body-0 coordinates; columns: v_world (3), omega_body (3)
    {
        M3 Rt = tr(Rw[0]);
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 3; k++) S0.m[3 + i][k] = Rt.m[i][k];
        for (int k = 0; k < 3; k++) S0.m[k][3 + k] = 1.0;
    }
    for (int i = nb - 1; i >= 1; i--) {
        V6 S;
        S.v[0] = axis[i].x; S.v[1] = axis[i].y; S.v[2] = axis[i].z;
        V6 F = mul(Ic[i], S);
        A(5 + i, 5 + i) = S.v[0] * F.v[0] + S.v[1] * F.v[1] + S.v[2] * F.v[2];
        int j = i;
        while (j > 0) {
            F = mulT(Xl[j], F);
            j = parent[j];
            if (j > 0) {
                double v = axis[j].x * F.v[0] + axis[j].y * F.v[1] + axis[j].z * F.v[2];
                A(5 + i, 5 + j) = v;
                A(5 + j, 5 + i) = v;
            } else {
                V6 b = mulT(S0, F);
                for (int k = 0; k < 6; k++) A(5 + i, k) = A(k, 5 + i) = b.v[k];
            }
        }
    }
    {
        M6 t = mul(tr(S0), mul(Ic[0], S0));
        for (int a = 0; a < 6; a++)
            for (int b = 0; b < 6; b++) A(a, b) = t.m[a][b];
    }
    // A_inv_ = A_.llt().solve(I), reference src/dwbc.cpp:307
    llt_solve_identity(A, A_inv);

    // COM position from A_, reference src/dwbc.cpp:320-324
    M3 A30;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A30.m[i][j] = A(3 + i, j) / total_mass;
    M3 skm = mul(Rw[0], A30);
    V3 cfp{skm.m[2][1], skm.m[0][2], skm.m[1][0]};
    com_pos = cfp + pw[0];
    // CMM_ = cm_rot6 * A_.topRows(6), reference src/dwbc.cpp:331-340
    Mat cm = Mat::identity(6);
    M3 skt = tr(skew3(cfp));
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            cm(3 + i, 3 + j) = Rw[0].m[i][j];
            cm(3 + i, j) = skt.m[i][j];
        }
    CMM = cm * A.top_rows(6);
    // B_ = NonlinearEffects (RNEA with qddot = 0, gravity (0,0,-9.81)), reference src/dwbc.cpp:343-344
    {
        std::vector<V6> acc(nb), f(nb);
        V6 aT;
        aT.v[5] = 9.81;
        V6 vJ0;
        vJ0.v[0] = qd[3]; vJ0.v[1] = qd[4]; vJ0.v[2] = qd[5];
        V6 a0 = mul(Xl[0], aT), c0 = crm(vel[0], vJ0);
        for (int k = 0; k < 6; k++) a0.v[k] += c0.v[k];
        acc[0] = a0;
        for (int i = 1; i < nb; i++) {
            V6 vJ;
            vJ.v[0] = axis[i].x * qd[5 + i]; vJ.v[1] = axis[i].y * qd[5 + i]; vJ.v[2] = axis[i].z * qd[5 + i];
            V6 a = mul(Xl[i], acc[parent[i]]), c = crm(vel[i], vJ);
            for (int k = 0; k < 6; k++) a.v[k] += c.v[k];
            acc[i] = a;
        }
        for (int i = 0; i < nb; i++) {
            M6 I = spatial_inertia(mass[i], com[i], inertia[i]);
            V6 Ia = mul(I, acc[i]), Iv = mul(I, vel[i]), c = crf(vel[i], Iv);
            for (int k = 0; k < 6; k++) f[i].v[k] = Ia.v[k] + c.v[k];
        }
        B.set_zero(n, 1);
        for (int i = nb - 1; i >= 1; i--) {
            B[5 + i] = axis[i].x * f[i].v[0] + axis[i].y * f[i].v[1] + axis[i].z * f[i].v[2];
            V6 fp = mulT(Xl[i], f[i]);
            for (int k = 0; k < 6; k++) f[parent[i]].v[k] += fp.v[k];
        }
        V6 fb = mulT(S0, f[0]);
        for (int k = 0; k < 6; k++) B[k] = fb.v[k];
    }
    // COM inertia, SI_body_, COM Jacobian, G_, reference src/dwbc.cpp:346-358
    M3 A33;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A33.m[i][j] = A(3 + i, 3 + j);
    M3 RAR = mul(Rw[0], mul(A33, tr(Rw[0])));
    M3 sc = skew3(cfp), scsct = mul(sc, tr(sc));
    Mat SI(6, 6);
    for (int i = 0; i < 3; i++) {
        SI(i, i) = total_mass;
        for (int j = 0; j < 3; j++) SI(3 + i, 3 + j) = RAR.m[i][j] - total_mass * scsct.m[i][j];
    }
    J_com = lu_inverse(SI) * CMM;
    G.set_zero(n, 1);
    for (int j = 0; j < n; j++) G[j] = -(J_com(2, j) * total_mass * (-9.81));
}

// RBDL CalcPointJacobian6D followed by the reference's row swap (linear rows on top),
// reference src/contact_constraint.cpp:56-62, src/link.cpp:98-107.
Mat Robot::point_jac(int body, V3 pl) const {
    Mat J(6, n);
    V3 p = pw[body] + mul(Rw[body], pl);
    M3 sk = mul(skew3(p - pw[0]), Rw[0]);
    for (int i = 0; i < 3; i++) {
        J(i, i) = 1.0;
        for (int j = 0; j < 3; j++) {
            J(i, 3 + j) = -sk.m[i][j];
            J(3 + i, 3 + j) = Rw[0].m[i][j];
        }
    }
    int i = body;
    while (i > 0) {
        V3 a = mul(Rw[i], axis[i]);
        V3 l = cross(a, p - pw[i]);
        J(0, 5 + i) = l.x; J(1, 5 + i) = l.y; J(2, 5 + i) = l.z;
        J(3, 5 + i) = a.x; J(4, 5 + i) = a.y; J(5, 5 + i) = a.z;
        i = parent[i];
    }
    return J;
}

// SetContact (reference include/dwbc.h:432-474) + UpdateContactConstraint (src/dwbc.cpp:433-454)
void Robot::set_contact(const unsigned char *flags) {
    contact_dof = 0;
    contact_links = 0;
    for (size_t i = 0; i < cc.size(); i++) {
        cc[i].on = flags[i] != 0;
        if (cc[i].on) contact_dof += 6, contact_links++;
        cc[i].xc = pw[cc[i].link] + mul(Rw[cc[i].link], cc[i].point);
        cc[i].rotm = Rw[cc[i].link];
        cc[i].J = point_jac(cc[i].link, cc[i].point);
    }
    J_C.set_zero(contact_dof, n);
    int row = 0;
    for (auto &c : cc)
        if (c.on) {
            J_C.set_block(row, 0, c.J);
            row += 6;
        }
}

// reference src/wbd.cpp:108-143
int Robot::calc_contact_constraint() {
    int c = contact_dof;
    Lambda_c = lu_inverse(J_C * A_inv * J_C.t());
    J_C_INV_T = Lambda_c * J_C * A_inv;
    N_C = Mat::identity(n) - J_C.t() * J_C_INV_T;
    A_inv_N_C = A_inv * N_C;
    W = A_inv_N_C.block(6, 6, m, m);
    V2.set_zero(std::max(c - 6, 0), m);
    NwJw.set_zero(m, m);  // reference src/dwbc.cpp:471 (33x33 zeros until overwritten)
    P_C.set_zero(c, 1);
    contact_ok = 1;
    Cod f = cod_compute(W, COD_THRESHOLD);
    W_inv = f.pinv;
    if (c > 6) {
        int rank = f.rank;
        V2 = f.Q.t().block(rank, 0, m - rank, m);
        if (c - 6 != V2.r) {  // "Contact Space Factorization Error" (src/wbd.cpp:130-135): the reference inverts a non-square matrix here
            contact_ok = 0;
            return contact_ok;
        }
        Mat tmp = lu_inverse(J_C_INV_T.block(0, 6, c - 6, m) * V2.t());  // topRightCorner(rows-6, cols-6)
        NwJw = V2.t() * tmp;
    }
    return contact_ok;
}

// reference src/wbd.cpp:186-192
void Robot::calc_grav() {
    torque_task.set_zero(m, 1);
    torque_contact.set_zero(m, 1);
    torque_grav = W_inv * (A_inv.bottom_rows(m) * (N_C * G));
    P_C = J_C_INV_T * G;
}

// reference src/dwbc.cpp:685-793 (Jacobian assembly only; the f* PD hooks are outside this path)
void Robot::update_task_space() {
    for (auto &t : ts) {
        if (t.custom || t.links.empty()) continue;
        t.J.set_zero(t.dof, n);
        int row = 0;
        for (auto &tl : t.links) {
            Mat Jl;
            bool is_com = tl.link == nb;
            switch (tl.mode) {
            case TASK_LINK_6D:
            case TASK_LINK_POSITION:
            case TASK_LINK_ROTATION:
            case TASK_LINK_ROTATION_CUSTOM_FRAME:
                Jl = is_com ? J_com : point_jac(tl.link, V3());
                break;
            case TASK_LINK_6D_COM_FRAME:
            case TASK_LINK_POSITION_COM_FRAME:
                Jl = is_com ? J_com : point_jac(tl.link, com[tl.link]);
                break;
            default:  // *_CUSTOM_FRAME
                Jl = is_com ? J_com : point_jac(tl.link, tl.point);
                break;
            }
            switch (tl.mode) {
            case TASK_LINK_6D: case TASK_LINK_6D_COM_FRAME: case TASK_LINK_6D_CUSTOM_FRAME:
                t.J.set_block(row, 0, Jl);
                break;
            case TASK_LINK_POSITION: case TASK_LINK_POSITION_COM_FRAME: case TASK_LINK_POSITION_CUSTOM_FRAME:
                t.J.set_block(row, 0, Jl.top_rows(3));
                break;
            default:
                t.J.set_block(row, 0, Jl.bottom_rows(3));
                break;
            }
            row += tl.dof;
        }
    }
}

// reference src/dwbc.cpp:795-816, src/wbd.cpp:207-213, 257-261
void Robot::calc_task_space() {
    update_task_space();
    for (size_t i = 0; i < ts.size(); i++) {
        Task &t = ts[i];
        t.Lambda = lu_inverse(t.J * A_inv * N_C * t.J.t());
        Mat Q = (t.Lambda * t.J * A_inv * N_C).right_cols(m);
        Mat QWQ = Q * W_inv * Q.t();
        t.Jkt = W_inv * Q.t() * cod_compute(QWQ, COD_THRESHOLD).pinv;
        if (i + 1 != ts.size()) {
            Mat prev = (i == 0) ? Mat::identity(m) : ts[i - 1].Null;
            t.Null = prev * (Mat::identity(m) - t.Jkt * t.Lambda * t.J * A_inv_N_C.right_cols(m));
        }
    }
}

// A_const_a and A_rot of the active contacts, reference src/dwbc.cpp:1055-1075 / src/wbd.cpp:59-97
void Robot::contact_const(Mat &A_const, Mat &A_rot, int &csize) const {
    csize = 10 * contact_links;
    A_const.set_zero(csize, contact_dof);
    A_rot.set_zero(contact_dof, contact_dof);
    int ci = 0, ki = 0;
    for (auto &c : cc) {
        if (!c.on) continue;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) {
                A_rot(ki + a, ki + b) = c.rotm.m[b][a];
                A_rot(ki + 3 + a, ki + 3 + b) = c.rotm.m[b][a];
            }
        // GetZMPConstMatrix
        A_const(ci + 0, ki + 2) = -c.lx; A_const(ci + 0, ki + 4) = -1;
        A_const(ci + 1, ki + 2) = -c.lx; A_const(ci + 1, ki + 4) = 1;
        A_const(ci + 2, ki + 2) = -c.ly; A_const(ci + 2, ki + 3) = -1;
        A_const(ci + 3, ki + 2) = -c.ly; A_const(ci + 3, ki + 3) = 1;
        // GetForceConstMatrix
        int f = ci + 4;
        A_const(f + 0, ki + 0) = 1.0;  A_const(f + 0, ki + 2) = -c.mu;
        A_const(f + 1, ki + 0) = -1.0; A_const(f + 1, ki + 2) = -c.mu;
        A_const(f + 2, ki + 1) = 1.0;  A_const(f + 2, ki + 2) = -c.mu;
        A_const(f + 3, ki + 1) = -1.0; A_const(f + 3, ki + 2) = -c.mu;
        A_const(f + 4, ki + 5) = 1.0;  A_const(f + 4, ki + 2) = -c.muz;
        A_const(f + 5, ki + 5) = -1.0; A_const(f + 5, ki + 2) = -c.muz;
        ci += 10;
        ki += 6;
    }
}

// reference src/dwbc.cpp:941-1127
int Robot::single_task_qp(Task &t, const Mat &null_prev, const Mat &torque_prev) {
    return single_task_qp_on(t, t.Jkt, m, null_prev, torque_prev, NwJw, J_C_INV_T, P_C, tlim_set, tlim);
}

// The QP of CalcSingleTaskTorqueWithQP (src/dwbc.cpp:941-1127) and of its reduced twin CalcSingleTaskTorqueWithQP_R
// (src/dwbc.cpp:3448-3599): the two differ only in the matrices they are handed (model size m / reduced_model_dof_).
int Robot::single_task_qp_on(Task &t, const Mat &Jkt, int m, const Mat &null_prev, const Mat &torque_prev, const Mat &NwJw, const Mat &J_C_INV_T,
                             const Mat &P_C, bool tlim_set, const Mat &tlim) {
    int td = t.dof, r = std::max(contact_dof - 6, 0);
    int nv = td + r;
    int tl_rows = tlim_set ? 2 * m : 0;
    Mat A_const, A_rot;
    int csize;
    contact_const(A_const, A_rot, csize);
    int rows = csize + tl_rows;
    Mat Aq(rows, nv), ub(rows, 1);
    Mat Nt = null_prev * Jkt * t.Lambda;
    Mat base = torque_prev + Nt * t.fstar;
    if (tlim_set) {
        Aq.set_block(0, 0, Nt);
        Aq.set_block(m, 0, -Nt);
        if (r > 0) {
            Aq.set_block(0, td, NwJw);
            Aq.set_block(m, td, -NwJw);
        }
        for (int i = 0; i < m; i++) {
            ub[i] = tlim[i] - base[i];
            ub[m + i] = tlim[i] + base[i];
        }
    }
    if (csize > 0) {
        Mat Atemp = A_const * A_rot * J_C_INV_T.right_cols(m);
        Aq.set_block(tl_rows, 0, -(Atemp * Nt));
        if (r > 0) Aq.set_block(tl_rows, td, -(Atemp * NwJw));
        Mat bA = A_const * (A_rot * P_C) - Atemp * base;
        for (int i = 0; i < csize; i++) ub[tl_rows + i] = -bA[i];
    }
    t.qpA = Aq;
    t.qpub = ub;
    QpResult res = lexqp(Aq, ub.d, td);
    t.qp_status = res.status;
    t.prox_iters = res.prox_iters;
    t.active = res.active;
    t.gray = res.gray;
    t.fstar_qp.set_zero(td, 1);
    t.contact_qp.set_zero(r, 1);
    if (!res.status) return 0;
    for (int i = 0; i < td; i++) t.fstar_qp[i] = res.x[i];
    for (int i = 0; i < r; i++) t.contact_qp[i] = res.x[td + i];
    return 1;
}

// reference src/dwbc.cpp:818-873
int Robot::calc_task_torque(bool hqp) {
    calc_task_space();
    torque_task.set_zero(m, 1);
    torque_contact.set_zero(m, 1);
    for (auto &t : ts) t.qp_status = 0, t.active.clear(), t.gray.clear(), t.fstar_qp.set_zero(t.dof, 1), t.contact_qp.set_zero(std::max(contact_dof - 6, 0), 1);
    for (size_t i = 0; i < ts.size(); i++) {
        Task &t = ts[i];
        if (hqp) {
            Mat prev_null = (i == 0) ? Mat::identity(m) : ts[i - 1].Null;
            Mat tprev = (i == 0) ? torque_grav : torque_grav + torque_task;
            if (!single_task_qp(t, prev_null, tprev)) return 0;
            t.torque_h = t.Jkt * t.Lambda * (t.fstar + t.fstar_qp);
            torque_task = (i == 0) ? t.torque_h : torque_task + ts[i - 1].Null * t.torque_h;
            // quirk 1: with no internal-force space the reference multiplies a
            // 33x33 zero matrix by an empty vector; defined here as zero.
            if (contact_dof > 6) torque_contact = NwJw * t.contact_qp;
            else torque_contact.set_zero(m, 1);
        } else {
            t.torque_h = t.Jkt * t.Lambda * t.fstar;
            torque_task = (i == 0) ? t.torque_h : torque_task + ts[i - 1].Null * t.torque_h;
        }
    }
    return 1;
}

// Eigen MatrixBase::eulerAngles(2, 1, 0) as documented (Eigen 3.3 / 3.4 Geometry/EulerAngles.h; Eigen is an unpinned
// third-party dependency, CMakeLists.txt:32): returns (angle about z, about y, about x) with the first angle in [0, pi].
// The reference stores it as Link::rpy (src/link.cpp:81), so rpy(2) is the angle about x of that decomposition.
static inline std::array<double, 3> eigen_euler_zyx(const M3 &R) {
    const int i = 2, j = 1, k = 0;  // a0 = 2, a1 = 1: odd permutation
    std::array<double, 3> res;
    res[0] = std::atan2(R.m[j][k], R.m[k][k]);
    const double c2 = std::sqrt(R.m[i][i] * R.m[i][i] + R.m[i][j] * R.m[i][j]);
    if (res[0] < 0.0) {
        res[0] += M_PI;
        res[1] = std::atan2(-R.m[i][k], -c2);
    } else {
        res[1] = std::atan2(-R.m[i][k], c2);
    }
    const double s1 = std::sin(res[0]), c1 = std::cos(res[0]);
    res[2] = std::atan2(s1 * R.m[k][i] - c1 * R.m[j][i], c1 * R.m[j][j] - s1 * R.m[k][j]);
    return res;
}

// ContactRedistributetwomod, reference src/wbd.cpp:273-404 (closed-form split of the resultant wrench between two feet)
static void contact_redistribute_two(double eta_cust, double footlength, double footwidth, double staticFrictionCoeff, double ratio_x,
                                     double ratio_y, V3 P1, V3 P2, const double *F12, double *RF, double *FR, double &eta) {
    // ResultantForce = [I 0 I 0; skew(P1) I skew(P2) I] F12
    V3 f1{F12[0], F12[1], F12[2]}, m1{F12[3], F12[4], F12[5]}, f2{F12[6], F12[7], F12[8]}, m2{F12[9], F12[10], F12[11]};
    V3 fs = f1 + f2, ms = cross(P1, f1) + m1 + cross(P2, f2) + m2;
    RF[0] = fs.x; RF[1] = fs.y; RF[2] = fs.z; RF[3] = ms.x; RF[4] = ms.y; RF[5] = ms.z;
    double eta_lb = 1.0 - eta_cust, eta_ub = eta_cust;
    auto clip = [&](double A, double B, double C) {
        double a = A * A, b = 2.0 * A * B, c = B * B - C * C;
        double s1 = (-b + std::sqrt(b * b - 4.0 * a * c)) / 2.0 / a;
        double s2 = (-b - std::sqrt(b * b - 4.0 * a * c)) / 2.0 / a;
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
    clip(Ax, Bx, ratio_y * footwidth / 2.0 * std::fabs(RF[2]));
    clip(Ay, By, ratio_x * footlength / 2.0 * std::fabs(RF[2]));
    clip(Az, Bz, staticFrictionCoeff * std::fabs(RF[2]));
    const double eta_s = (-RF[3] - P2.z * RF[1] + P2.y * RF[2]) / ((P1.z - P2.z) * RF[1] - (P1.y - P2.y) * RF[2]);
    eta = eta_s;
    if (eta_s > eta_ub) eta = eta_ub;
    else if (eta_s < eta_lb) eta = eta_lb;
    if ((eta > eta_cust) || (eta < 1.0 - eta_cust)) eta = 0.5;
    FR[0] = eta * RF[0]; FR[1] = eta * RF[1]; FR[2] = eta * RF[2];
    FR[3] = Ax * eta * eta + Bx * eta;
    FR[4] = Ay * eta * eta + By * eta;
    FR[5] = Az * eta * eta + Bz * eta;
    FR[6] = (1.0 - eta) * RF[0]; FR[7] = (1.0 - eta) * RF[1]; FR[8] = (1.0 - eta) * RF[2];
    FR[9] = (1.0 - eta) * (Ax * eta + Bx);
    FR[10] = (1.0 - eta) * (Ay * eta + By);
    FR[11] = (1.0 - eta) * (Az * eta + Bz);
}

// getContactForce -> CalculateContactForce, reference src/dwbc.cpp:891-896, src/wbd.cpp:268-271
Mat Robot::get_contact_force(const Mat &command_torque) const { return J_C_INV_T.right_cols(m) * command_torque - P_C; }

// getZMP, reference src/dwbc.cpp:898-939 (as written: registered contact i reads contact_force(6 i + ...), i.e. the active
// contacts are assumed to be a prefix of the registered ones)
V3 Robot::get_zmp(const Mat &cf, std::vector<V3> *per_contact) const {
    V3 zmp{0, 0, 0};
    double total_fz = 0;
    for (size_t i = 0; i < cc.size(); i++)
        if (cc[i].on) total_fz += cf[2 + i * 6];
    if (per_contact) per_contact->assign(cc.size(), V3{0, 0, 0});
    for (size_t i = 0; i < cc.size(); i++) {
        if (!cc[i].on) continue;
        V3 z = cc[i].xc;
        if (!(cf[i * 6 + 2] > -1.0e-3)) {
            z.x = cc[i].xc.x + (-cf[i * 6 + 4] / cf[i * 6 + 2]);
            z.y = cc[i].xc.y + (cf[i * 6 + 3] / cf[i * 6 + 2]);
        }
        if (per_contact) (*per_contact)[i] = z;
        zmp = zmp + (cf[i * 6 + 2] / total_fz) * z;
    }
    return zmp;
}

// reference src/dwbc.cpp:1377-1568 (QP branch).  The reference sizes the QP with
// contact_link_num_*contact_constraint_size rows (quirk 5): the extra rows are 0*x in [0,0]
// and are dropped here.
// The closed form shared by CalcContactRedistribute(hqp = false) (reference src/dwbc.cpp:1570-1619) and the analytic branch of
// CalcContactRedistributeR (:4940-4985): ContactRedistributetwomod on the yaw-aligned contact forces, then the torque that
// realises the change of the second foot's wrench through the internal-force space V2.
Mat Robot::redistribute_analytic(const Mat &JCINVT, const Mat &PC, const Mat &V2m, const Mat &tin, int msize) {
    Mat cf = JCINVT.right_cols(msize) * tin - PC;  // ContactForce_
    const double ang = -eigen_euler_zyx(Rw[0])[2];  // -link_[0].rpy(2), rpy = rotm.eulerAngles(2, 1, 0) (src/link.cpp:81)
    M3 Rz = eye3();  // rotateWithZ, reference src/math.cpp:55-74
    Rz.m[0][0] = std::cos(ang); Rz.m[0][1] = -std::sin(ang);
    Rz.m[1][0] = std::sin(ang); Rz.m[1][1] = std::cos(ang);
    V3 P1 = mul(Rz, cc[0].xc - com_pos), P2 = mul(Rz, cc[1].xc - com_pos);
    double F12[12], FR[12], res6[6], eta = 0.95;
    for (int k = 0; k < 4; k++) {
        V3 f = mul(Rz, V3{cf[3 * k], cf[3 * k + 1], cf[3 * k + 2]});
        F12[3 * k] = f.x; F12[3 * k + 1] = f.y; F12[3 * k + 2] = f.z;
    }
    contact_redistribute_two(0.99, 0.26, 0.1, 1.0, 0.9, 0.9, P1, P2, F12, res6, FR, eta);
    Mat des(6, 1);
    M3 Rzt = tr(Rz);
    for (int k = 2; k < 4; k++) {
        V3 f = mul(Rzt, V3{FR[3 * k], FR[3 * k + 1], FR[3 * k + 2]});
        des[3 * (k - 2)] = -cf[3 * k] + f.x;
        des[3 * (k - 2) + 1] = -cf[3 * k + 1] + f.y;
        des[3 * (k - 2) + 2] = -cf[3 * k + 2] + f.z;
    }
    redis_eta = eta;
    return V2m.t() * lu_inverse(JCINVT.right_cols(msize).bottom_rows(6) * V2m.t()) * des;
}

int Robot::calc_redistribute(bool hqp) {
    redis_status = 0;
    redis_active.clear();
    redis_gray.clear();
    Mat tin = torque_grav + torque_task + torque_contact;
    int r = contact_dof - 6;
    if (!hqp) {
        // analytic two-foot branch, reference src/dwbc.cpp:1570-1619: the first two REGISTERED contacts (cc_[0], cc_[1])
        if (contact_dof != 12) {
            torque_contact.set_zero(m, 1);
            return 0;
        }
        torque_contact = redistribute_analytic(J_C_INV_T, P_C, V2, tin, m);
        redis_status = 1;
        return 1;
    }
    if (r <= 0) {
        torque_contact.set_zero(m, 1);
        redis_status = 1;
        return 1;
    }
    Mat A_const, A_rot;
    int csize;
    contact_const(A_const, A_rot, csize);
    int tl_rows = tlim_set ? 2 * m : 0;
    Mat Aq(tl_rows + csize, r), ub(tl_rows + csize, 1);
    if (tlim_set) {
        Aq.set_block(0, 0, NwJw);
        Aq.set_block(m, 0, -NwJw);
        for (int i = 0; i < m; i++) {
            ub[i] = tlim[i] - tin[i];
            ub[m + i] = tlim[i] + tin[i];
        }
    }
    Mat CM = -(A_const * A_rot);
    Mat CJ = CM * J_C_INV_T.right_cols(m);
    Aq.set_block(tl_rows, 0, CJ * NwJw);
    Mat rhs = CM * P_C - CJ * tin;
    for (int i = 0; i < csize; i++) ub[tl_rows + i] = rhs[i];
    redisA = Aq;
    redisub = ub;
    QpResult res = qp_min_norm(Aq, ub.d);
    if (!res.status) {
        torque_contact.set_zero(m, 1);
        return 0;
    }
    fill_active(Aq, ub.d, res);
    redis_active = res.active;
    redis_gray = res.gray;
    cf_redis_qp.set_zero(r, 1);
    for (int i = 0; i < r; i++) cf_redis_qp[i] = res.x[i];
    torque_contact = torque_contact + NwJw * cf_redis_qp;
    redis_status = 1;
    return 1;
}


// ==========================================================================================
// a9: reduced-dimension variant (Ahn 2024).  Literal restatement of reference src/dwbc.cpp:2752-3446,
// src/task.cpp:95-142, src/wbd.cpp:220-226.  No goldens exist for it: parity is pinned by the property
// "reduced torques == full-model torques" printed by reference tests/sp_test/redu_dyn_test.cpp:323-360
// and asserted in tests/test_oracle_reduced.py.  hqp=true (CalcSingleTaskTorqueWithQP_R[_NC]) is not restated yet.
// ==========================================================================================
static inline bool has(const std::vector<int> &v, int x) { return std::find(v.begin(), v.end(), x) != v.end(); }

// ReducedDynamicsCalculate, reference src/dwbc.cpp:2752-2989
void Robot::reduced_dynamics() {
    co_link.clear(); nc_link.clear(); co_joint.clear(); nc_joint.clear(); vc_joint.clear();
    co_link.push_back(0);
    for (auto &c : cc)
        if (c.on) {
            int l = c.link;
            while (l != 0) {
                co_link.push_back(l);
                co_joint.push_back(5 + l);  // joint_[l].joint_id_ = q index of the joint that moves link l
                l = parent[l];
            }
        }
    std::sort(co_link.begin(), co_link.end());
    for (int i = 0; i < nb; i++)
        if (!has(co_link, i)) nc_link.push_back(i);
    std::sort(co_joint.begin(), co_joint.end());
    for (int i = 0; i < 6; i++) vc_joint.push_back(i);
    vc_joint.insert(vc_joint.end(), co_joint.begin(), co_joint.end());
    for (int l0 : nc_link) {
        int l = l0;
        while (l != 0) {
            if (!has(nc_joint, 5 + l)) nc_joint.push_back(5 + l);
            l = parent[l];
            if (has(nc_link, l)) break;
        }
    }
    std::sort(nc_joint.begin(), nc_joint.end());
    nc_dof = (int)nc_joint.size();
    co_dof = (int)co_joint.size();
    vc_dof = co_dof + 6;
    rmod = co_dof + 6;
    rsys = rmod + 6;

    // CRBA restricted to the non-contact bodies, reference :2828-2878
    std::vector<M6> Ic(nb);
    for (int i = 0; i < nb; i++)
        if (has(nc_link, i)) Ic[i] = spatial_inertia(mass[i], com[i], inertia[i]);
    Mat A_NC_O(n, n);
    M6 S0;
    {
        M3 Rt = tr(Rw[0]);
        for (int i = 0; i < 3; i++)
            for (int k = 0; k < 3; k++) S0.m[3 + i][k] = Rt.m[i][k];
        for (int k = 0; k < 3; k++) S0.m[k][3 + k] = 1.0;
    }
    for (int i = nb - 1; i >= 1; i--) {
        if (!has(nc_link, i)) continue;
        {   // the RBDL parent of every real body is non-zero (the pelvis hangs off the translation body)
            M6 t = mul(tr(Xl[i]), mul(Ic[i], Xl[i]));
            M6 &P = Ic[parent[i]];
            for (int a = 0; a < 6; a++)
                for (int b = 0; b < 6; b++) P.m[a][b] += t.m[a][b];
        }
        V6 S;
        S.v[0] = axis[i].x; S.v[1] = axis[i].y; S.v[2] = axis[i].z;
        V6 F = mul(Ic[i], S);
        A_NC_O(5 + i, 5 + i) = S.v[0] * F.v[0] + S.v[1] * F.v[1] + S.v[2] * F.v[2];
        int j = i;
        while (j > 0) {
            F = mulT(Xl[j], F);
            j = parent[j];
            if (j > 0) {
                double v = axis[j].x * F.v[0] + axis[j].y * F.v[1] + axis[j].z * F.v[2];
                A_NC_O(5 + i, 5 + j) = v;
                A_NC_O(5 + j, 5 + i) = v;
            } else {
                V6 b = mulT(S0, F);
                for (int k = 0; k < 6; k++) A_NC_O(5 + i, k) = A_NC_O(k, 5 + i) = b.v[k];
            }
        }
    }
    // SI_nc_b_ from model_.Ic[2] (composite of the nc chain at the pelvis, pelvis frame), reference :2880-2885
    SI_nc_b.set_zero(6, 6);
    const double M = Ic[0].m[3][3];
    for (int i = 0; i < 3; i++) {
        SI_nc_b(i, i) = M;
        for (int j = 0; j < 3; j++) {
            SI_nc_b(3 + i, 3 + j) = Ic[0].m[i][j];
            SI_nc_b(3 + i, j) = Ic[0].m[i][3 + j];
            SI_nc_b(j, 3 + i) = Ic[0].m[i][3 + j];
        }
    }
    // InertiaMatrixSegment / InertiaMatrix, reference src/math.cpp:304-320
    mass_nc = SI_nc_b(0, 0);
    com_pos_nc = {SI_nc_b(5, 1) / mass_nc, SI_nc_b(3, 2) / mass_nc, SI_nc_b(4, 0) / mass_nc};
    M3 sc = skew3(com_pos_nc), scsct = mul(sc, tr(sc));
    SI_nc_l.set_zero(6, 6);
    for (int i = 0; i < 3; i++) {
        SI_nc_l(i, i) = mass_nc;
        for (int j = 0; j < 3; j++) SI_nc_l(3 + i, 3 + j) = SI_nc_b(3 + i, 3 + j) - mass_nc * scsct.m[i][j];
    }
    // A_NC, reference :2890-2906 (only its top-right 6 x nc block is rotated; the bottom-left is left as is)
    A_NC.set_zero(nc_dof + 6, nc_dof + 6);
    A_NC.set_block(0, 0, SI_nc_b);
    for (int i = 0; i < nc_dof; i++) {
        for (int j = 0; j < 6; j++) {
            A_NC(6 + i, j) = A_NC_O(nc_joint[i], j);
            A_NC(j, 6 + i) = A_NC_O(j, nc_joint[i]);
        }
        for (int j = 0; j < nc_dof; j++) A_NC(6 + i, 6 + j) = A_NC_O(nc_joint[i], nc_joint[j]);
    }
    {
        Mat R0t(3, 3);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) R0t(a, b) = Rw[0].m[b][a];
        A_NC.set_block(0, 6, R0t * A_NC.block(0, 6, 3, nc_dof));
    }
    // cmm_nc_, J_I_nc_, reference :2908-2916
    Mat cm_rot6 = Mat::identity(6);
    {
        M3 skt = tr(sc);
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) cm_rot6(3 + i, j) = skt.m[i][j];
    }
    cmm_nc = cm_rot6 * A_NC.block(0, 6, 6, nc_dof);
    J_I_nc = lu_inverse(SI_nc_l) * cmm_nc;
    // J_R, reference :2918-2929
    J_R.set_zero(rsys, n);
    for (int i = 0; i < vc_dof; i++) J_R(i, vc_joint[i]) = 1.0;
    for (int i = 0; i < nc_dof; i++)
        for (int k = 0; k < 6; k++) J_R(vc_dof + k, nc_joint[i]) = J_I_nc(k, i);
    // A_R_inv assembled blockwise, reference :2936-2956
    A_R_inv.set_zero(rsys, rsys);
    Mat A_tr_temp(vc_dof, n);
    for (int i = 0; i < vc_dof; i++) {
        for (int j = 0; j < vc_dof; j++) A_R_inv(i, j) = A_inv(vc_joint[i], vc_joint[j]);
        for (int j = 0; j < n; j++) A_tr_temp(i, j) = A_inv(vc_joint[i], j);
    }
    Mat tr_blk = A_tr_temp.right_cols(nc_dof) * J_I_nc.t();
    A_R_inv.set_block(0, vc_dof, tr_blk);
    A_R_inv.set_block(vc_dof, 0, tr_blk.t());
    A_R_inv.set_block(vc_dof, vc_dof, J_R.bottom_rows(6) * A_inv * J_R.bottom_rows(6).t());
    llt_solve_identity(A_R_inv, A_R);  // reference :2958
    // reference :2971-2989 (block indexing assumes q order [base | contact chain | non-contact chain])
    J_I_nc_inv_T = A_R.block(vc_dof, 0, 6, vc_dof) * A_inv.block(0, vc_dof, vc_dof, nc_dof) +
                   A_R.block(vc_dof, vc_dof, 6, 6) * J_I_nc * A_inv.block(vc_dof, vc_dof, nc_dof, nc_dof);
    N_I_nc = Mat::identity(nc_dof) - J_I_nc.t() * J_I_nc_inv_T;
    G_R.set_zero(rsys, 1);
    for (int i = 0; i < vc_dof; i++) G_R[i] = G[i];
    G_R.set_block(vc_dof, 0, J_I_nc_inv_T * G.block(vc_dof, 0, nc_dof, 1));
    J_CR.set_zero(contact_dof, rsys);
    J_CR.set_block(0, 0, J_C.block(0, 0, contact_dof, vc_dof));
}

// ReducedCalcContactConstraint, reference src/dwbc.cpp:3077-3142
int Robot::reduced_contact_constraint() {
    const int c = contact_dof, cols = rsys;
    J_CR.set_zero(c, rsys);
    J_CR.set_block(0, 0, J_C.block(0, 0, c, vc_dof));
    Mat Jl = J_CR.left_cols(c + 6);
    Lambda_CR = lu_inverse(Jl * A_R_inv.block(0, 0, c + 6, c + 6) * Jl.t());
    J_C_INV_T = (Lambda_CR * J_C) * A_inv;
    J_CR_INV_T.set_zero(c, rsys);
    J_CR_INV_T.set_block(0, 0, J_C_INV_T.block(0, 0, c, vc_dof));
    J_CR_INV_T.set_block(0, vc_dof, J_C_INV_T.block(0, vc_dof, c, nc_dof) * J_I_nc.t());
    N_C = Mat::identity(n) - J_C.t() * J_C_INV_T;
    N_CR = Mat::identity(cols) - J_CR.t() * J_CR_INV_T;
    A_R_inv_N_CR = A_R_inv * N_CR;
    W_R = A_R_inv_N_CR.block(6, 6, cols - 6, cols - 6);
    V2_R.set_zero(std::max(c - 6, 0), rmod);
    NwJw_R.set_zero(rmod, rmod);
    Cod f = cod_compute(W_R, COD_THRESHOLD);
    W_R_inv = f.pinv;
    reduced_ok = 1;
    if (c > 6) {
        const int rank = cols - c;  // force_rank, reference :3119
        V2_R = f.Q.t().block(rank, 0, rmod - rank, rmod);
        NwJw_R = V2_R.t() * lu_inverse(J_CR_INV_T.right_cols(cols - 6).top_rows(c - 6) * V2_R.t());
        if (c - 6 != V2_R.r) reduced_ok = 0;
    }
    return reduced_ok;
}

// ReducedCalcGravCompensation, reference src/dwbc.cpp:3144-3150
void Robot::reduced_grav() {
    torque_grav_R = W_R_inv * (A_R_inv.bottom_rows(rmod) * (N_CR * G_R));
    if (torque_grav.r != m) torque_grav.set_zero(m, 1);
    for (int i = 0; i < rmod; i++) torque_grav[i] = torque_grav_R[i];
    for (int i = 0; i < nc_dof; i++) torque_grav[co_dof + i] = G[vc_dof + i];
    P_CR = J_CR_INV_T * G_R;
}

// CalculateJKT_R, reference src/wbd.cpp:220-226
static void calculate_jkt_r(const Mat &J_task_R, const Mat &A_R_inv_N_CR, const Mat &W_R_inv, Mat &J_kt_R, Mat &lambda_R) {
    lambda_R = lu_inverse(J_task_R * A_R_inv_N_CR * J_task_R.t());
    Mat Q = (lambda_R * J_task_R * A_R_inv_N_CR).right_cols(J_task_R.c - 6);
    J_kt_R = W_R_inv * Q.t() * cod_compute(Q * W_R_inv * Q.t(), COD_THRESHOLD).pinv;
}

// ReducedCalcTaskSpace, reference src/dwbc.cpp:3152-3253 + TaskSpace::CalcJKT_R src/task.cpp:95-142.
// Returns 0 for the reference's "UNDEFINED TASK TYPE" (a level that mixes contact-chain and non-contact links).
int Robot::reduced_task_space() {
    A_inv_N_C = A_inv * N_C;
    update_task_space();
    J_base_R = point_jac(0, V3()).left_cols(rsys);
    calculate_jkt_r(J_base_R, A_R_inv_N_CR, W_R_inv, J_base_R_kt, lambda_base_R);
    int nch = 0;
    for (auto &t : ts) {
        if (t.custom) {
            t.noncont = t.reduced = t.cmm = false;
            if (t.J.left_cols(vc_dof).norm() > 1.0e-4) t.reduced = true;
            if (t.J.right_cols(nc_dof).norm() > 1.0e-4) {
                t.noncont = true;
                t.nc_h = nch++;
            }
        } else {
            for (auto &tl : t.links) {
                if (tl.link == nb) t.cmm = true;
                else if (has(co_link, tl.link)) t.reduced = true;
                else {
                    t.noncont = true;
                    t.nc_h = nch++;
                }
            }
        }
    }
    for (size_t i = 0; i < ts.size(); i++) {
        Task &t = ts[i];
        const int td = t.dof;
        t.J_R.set_zero(td, rsys);
        t.Lambda_R.set_zero(td, td);
        t.Jkt_R.set_zero(rsys - 6, td);
        t.force_on_nc.set_zero(6, 1);
        if (t.cmm) {
            t.J_R.set_block(0, 0, t.J.block(0, 0, td, vc_dof));
            t.J_R.set_block(0, vc_dof, t.J.block(0, vc_dof, td, nc_dof) * J_I_nc_inv_T.t());
            calculate_jkt_r(t.J_R, A_R_inv_N_CR, W_R_inv, t.Jkt_R, t.Lambda_R);
            t.Lambda = t.Lambda_R;
        } else if (t.reduced && !t.noncont) {
            t.J_R.set_block(0, 0, t.J.block(0, 0, td, vc_dof));
            calculate_jkt_r(t.J_R, A_R_inv_N_CR, W_R_inv, t.Jkt_R, t.Lambda_R);
            t.Lambda = t.Lambda_R;
        } else if (!t.reduced && t.noncont) {
            t.Lambda = lu_inverse(t.J * A_inv_N_C * t.J.t());
            t.J_NC = t.J.block(0, vc_dof, td, nc_dof);
            t.J_R.set_block(0, 0, t.J.block(0, 0, td, vc_dof));
            t.J_R.set_block(0, vc_dof, t.J.block(0, vc_dof, td, nc_dof) * J_I_nc_inv_T.t());
        } else {
            return 0;  // "UNDEFINED TASK TYPE"
        }
        if (i + 1 != ts.size()) {
            if (!t.noncont) {
                Mat prev = (i == 0) ? Mat::identity(rmod) : ts[i - 1].Null_R;
                t.Null_R = prev * (Mat::identity(rmod) - t.Jkt_R * t.Lambda_R * t.J_R * A_R_inv_N_CR.right_cols(rmod));
            } else {
                if (i == 0) return 0;  // the reference reads ts_[-1]
                t.Null_R = ts[i - 1].Null_R;
            }
        }
    }
    return 1;
}

// ReducedCalcTaskControlTorque(hqp=false), reference src/dwbc.cpp:3255-3446 (else-branch :3374-3417, expansion :3442-3443)
int Robot::reduced_task_torque(bool hqp) {
    int first_nc = -1;
    force_on_nc_r.set_zero(6, 1);
    force_on_nc_R_qp.set_zero(6, 1);
    torque_task_R.set_zero(rmod, 1);
    torque_task.set_zero(m, 1);
    torque_task_NC.set_zero(nc_dof, 1);
    Mat R0(3, 3);
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) R0(a, b) = Rw[0].m[a][b];
    for (size_t i = 0; i < ts.size(); i++) {
        Task &t = ts[i];
        t.torque_nc.set_zero(nc_dof, 1);
        t.torque_h_R.set_zero(rmod, 1);
        t.torque_null_h_nc.set_zero(nc_dof, 1);
        t.torque_null_h_R.set_zero(rmod, 1);
        if (t.noncont) {
            if (i == 0) return 0;
            Mat temp = t.J.t() * (t.Lambda * t.fstar);
            t.torque_nc = temp.block(vc_dof, 0, nc_dof, 1);
            t.force_on_nc.set_block(0, 0, temp.block(0, 0, 3, 1));
            t.force_on_nc.set_block(3, 0, R0 * temp.block(3, 0, 3, 1));
            if (hqp) force_on_nc_r = force_on_nc_r + t.force_on_nc;  // reference src/dwbc.cpp:3302
            t.torque_h_R.set_block(0, 0, J_base_R_kt.top_rows(co_dof) * t.force_on_nc);
            t.torque_h_R.set_block(co_dof, 0, J_I_nc_inv_T * t.torque_nc);
            if (t.nc_h == 0) {
                first_nc = (int)i;
                t.torque_null_h_R = ts[i - 1].Null_R * t.torque_h_R;
                t.torque_null_h_nc = t.torque_nc;
            } else {
                Task &p = ts[i - 1];
                Mat null_force = p.Lambda * (p.J * (A_inv_N_C * (t.J.t() * (t.Lambda * t.fstar))));
                Mat null_torque_h_r(rmod, 1);
                temp = p.J.t() * null_force;
                t.torque_null_h_nc = t.torque_nc - temp.block(vc_dof, 0, nc_dof, 1);
                temp.set_block(3, 0, R0 * temp.block(3, 0, 3, 1));
                null_torque_h_r.set_block(0, 0, t.torque_h_R.block(0, 0, co_dof, 1) - J_base_R_kt.top_rows(co_dof) * temp.block(0, 0, 6, 1));
                null_torque_h_r.set_block(co_dof, 0, J_I_nc_inv_T * (t.torque_nc - p.J_NC.t() * null_force));
                t.torque_null_h_R = p.Null_R * null_torque_h_r;
                if (hqp) force_on_nc_r = force_on_nc_r - temp.block(0, 0, 6, 1);  // :3325
            }
            if (hqp) continue;  // under hqp the non-contact levels are summed after the base-wrench QP (:3405-3418)
        } else if (hqp) {
            // reference src/dwbc.cpp:3336-3362
            t.fstar_qp.set_zero(t.dof, 1);
            t.contact_qp.set_zero(std::max(contact_dof - 6, 0), 1);
            Mat prev_null = (i == 0) ? Mat::identity(rmod) : ts[i - 1].Null_R;
            Mat tprev = (i == 0) ? torque_grav_R : torque_grav_R + torque_task_R;
            if (!single_task_qp_on(t, t.Jkt_R, rmod, prev_null, tprev, NwJw_R, J_CR_INV_T, P_CR, tlimR_set, tlimR)) return 0;
            t.torque_h_R = t.Jkt_R * t.Lambda * (t.fstar + t.fstar_qp);
            t.torque_null_h_R = (i == 0) ? t.torque_h_R : ts[i - 1].Null_R * t.torque_h_R;
        } else {
            t.torque_h_R = t.Jkt_R * t.Lambda * t.fstar;
            t.torque_null_h_R = (i == 0) ? t.torque_h_R : ts[i - 1].Null_R * t.torque_h_R;
        }
        torque_task_R = torque_task_R + t.torque_null_h_R;
        torque_task_NC = torque_task_NC + t.torque_null_h_nc;
    }
    Mat torque_task_R_qp(rmod, 1);
    if (hqp && first_nc >= 0) {
        // CalcSingleTaskTorqueWithQP_R_NC (reference src/dwbc.cpp:3601-3756): the QP of CalcSingleTaskTorqueWithQP_R with the "task"
        // Ntorque = Null_{first_nc-1} J_base_R_kt (6 variables: a correction of the wrench the non-contact levels put on the
        // condensed body) and f* replaced by force_on_nc_r_
        Task ncq;
        ncq.dof = 6;
        ncq.Lambda = Mat::identity(6);
        ncq.fstar = force_on_nc_r;
        const int okq = single_task_qp_on(ncq, J_base_R_kt, rmod, ts[first_nc - 1].Null_R, torque_grav_R + torque_task_R, NwJw_R, J_CR_INV_T, P_CR, tlimR_set, tlimR);
        nc_qpA = ncq.qpA;
        nc_qpub = ncq.qpub;
        if (!okq) return 0;
        force_on_nc_R_qp = ncq.fstar_qp;
        nc_qp_contact = ncq.contact_qp;
        nc_qp_active = ncq.active;
        for (size_t i = first_nc; i < ts.size(); i++)
            if (ts[i].noncont) {
                torque_task_R = torque_task_R + ts[i].torque_null_h_R;
                torque_task_NC = torque_task_NC + ts[i].torque_null_h_nc;
            }
        torque_task_R_qp = J_base_R_kt * force_on_nc_R_qp;  // only rows 0..co_dof-1 are used (:3420)
    }
    torque_task.set_block(0, 0, torque_task_R.block(0, 0, co_dof, 1) + torque_task_R_qp.block(0, 0, co_dof, 1));
    torque_task.set_block(co_dof, 0, J_I_nc.t() * torque_task_R.block(co_dof, 0, 6, 1) + N_I_nc * torque_task_NC);
    return 1;
}

// ReducedCalcContactRedistribute(hqp = true) -> CalcContactRedistributeR, reference src/dwbc.cpp:3758-3770, 4762-4960: the QP of
// CalcContactRedistribute on the reduced matrices with torque_input = torque_grav_R_ + torque_task_R_; then
// torque_contact_[0:contact_dof_] = torque_contact_R_[0:contact_dof_].
int Robot::reduced_redistribute(bool hqp) {
    torque_contact.set_zero(m, 1);
    torque_contact_R.set_zero(rmod, 1);
    const int r = contact_dof - 6;
    if (r <= 0) return 0;  // reference: ret stays 0 when contact_dof_ <= 6
    Mat tin = torque_grav_R + torque_task_R;
    if (!hqp) {  // analytic branch, reference src/dwbc.cpp:4938-4990
        if (co_dof != 12) return 0;
        torque_contact_R = redistribute_analytic(J_CR_INV_T, P_CR, V2_R, tin, rmod);
        for (int i = 0; i < contact_dof; i++) torque_contact[i] = torque_contact_R[i];
        return 1;
    }
    Mat A_const, A_rot;
    int csize;
    contact_const(A_const, A_rot, csize);
    const int tl_rows = tlimR_set ? 2 * rmod : 0;
    Mat Aq(tl_rows + csize, r), ub(tl_rows + csize, 1);
    if (tlimR_set) {
        Aq.set_block(0, 0, NwJw_R);
        Aq.set_block(rmod, 0, -NwJw_R);
        for (int i = 0; i < rmod; i++) {
            ub[i] = tlimR[i] - tin[i];
            ub[rmod + i] = tlimR[i] + tin[i];
        }
    }
    Mat CM = -(A_const * A_rot);
    Mat CJ = CM * J_CR_INV_T.right_cols(rmod);
    Aq.set_block(tl_rows, 0, CJ * NwJw_R);
    Mat rhs = CM * P_CR - CJ * tin;
    for (int i = 0; i < csize; i++) ub[tl_rows + i] = rhs[i];
    QpResult res = qp_min_norm(Aq, ub.d);
    if (!res.status) return 0;
    cf_redis_qp_R.set_zero(r, 1);
    for (int i = 0; i < r; i++) cf_redis_qp_R[i] = res.x[i];
    torque_contact_R = NwJw_R * cf_redis_qp_R;
    for (int i = 0; i < contact_dof; i++) torque_contact[i] = torque_contact_R[i];
    return 1;
}

}  // namespace orc

// ------------------------------------------------------------------------------------------
// C API (ctypes) -- test/bench harness only.
// ------------------------------------------------------------------------------------------
using namespace orc;


// ================================================================================================
// a10: LQP / centroidal simplified-WBD variant.  PARITY UNPINNED: the reference solves every level
// with OSQP (ADMM, default eps 1e-3, sigma 1e-6) and holds no golden vector for this path; this
// restatement assembles the level QPs exactly as the reference does and solves them EXACTLY.
// ================================================================================================

// OSQP regularises its KKT system with sigma = 1e-6 (default setting, not overridden by the reference,
// src/dwbc_hqp.cpp:596-598).  The level Hessians are singular (internal contact forces carry no cost), so the
// exact solve needs a rule for those directions: the same sigma on the null-space coordinates u, i.e. among
// the level's minimisers the one closest to the previous level's y (Z is orthonormal).
static const double LQP_SIGMA = 1.0e-6;

// getNullSpace (reference src/math.cpp:349-360): orthonormal basis of null(A), A is me x nvar.
static Mat lqp_null_space(const Mat &A) {
    Cod f = cod_compute(A.t(), 1e-10);  // A^T P = Q R;  null(A) = last (nvar - rank) columns of Q
    return f.Q.right_cols(A.c - f.rank);
}

static void lqp_normalize(Mat &A, Mat &a) {  // HQP_Hierarch::normalizeConstraintMatrix, reference src/dwbc_hqp.cpp:555-581
    for (int i = 0; i < A.r; i++) {
        double s = 0;
        for (int j = 0; j < A.c; j++) s += A(i, j) * A(i, j);
        s = std::sqrt(s);
        if (s > 0) {
            for (int j = 0; j < A.c; j++) A(i, j) /= s;
            a[i] /= s;
        }
    }
}

// reference src/dwbc.cpp:4304-4430 (ConfigureLQP) and, with reduced = true, :4504-4632 (ConfigureLQP_R): the same hierarchy on
// A_R, A_R_inv, G_R (gravity only), J_CR, J_task J_R_INV_T^T, torque limit 200 except 600 on entry size-4 (:4529-4531), cost
// A_R / |A_| (the norm of the FULL mass matrix, :4519); only the COM / contact-chain task levels are added (:4609-4627).
void Robot::configure_lqp(bool reduced) {
    lqp_reduced = reduced;
    const Mat &A = reduced ? this->A_R : this->A;
    const Mat &A_inv = reduced ? this->A_R_inv : this->A_inv;
    const Mat &B = reduced ? this->G_R : this->B;
    const Mat &J_C = reduced ? this->J_CR : this->J_C;
    const int n = reduced ? rsys : this->n, m = reduced ? rmod : this->m;
    const int acc = n, tq = m, con = contact_dof, nvar = acc + con;
    lqp_acc = acc;
    lqp_con = con;
    lqp.clear();
    lqp_H.set_zero(nvar, nvar);
    const double an = this->A.norm();
    for (int j = 0; j < acc; j++)
        for (int i = 0; i < acc; i++) lqp_H(i, j) = A(i, j) / an * 5.0;  // :4341
    Mat JcT = J_C.t();
    // Priority 1: floating-base Newton-Euler rows (equality), torque limit +-200 (inequality); preset answer  :4343-4381
    {
        LqpLevel L;
        L.nineq = 2 * tq;
        L.neq = 6;
        L.A.set_zero(L.nineq, nvar);
        L.a.set_zero(L.nineq, 1);
        L.B.set_zero(6, nvar);
        L.b.set_zero(6, 1);
        for (int j = 0; j < acc; j++)
            for (int i = 0; i < 6; i++) L.B(i, j) = A(i, j);
        for (int j = 0; j < con; j++)
            for (int i = 0; i < 6; i++) L.B(i, acc + j) = JcT(i, j);
        for (int i = 0; i < 6; i++) L.b[i] = B[i];
        for (int i = 0; i < tq; i++) {
            for (int j = 0; j < acc; j++) L.A(i, j) = A(6 + i, j), L.A(tq + i, j) = -A(6 + i, j);
            for (int j = 0; j < con; j++) L.A(i, acc + j) = JcT(6 + i, j), L.A(tq + i, acc + j) = -JcT(6 + i, j);
            const double tl = (reduced && i == tq - 4) ? 600.0 : 200.0;
            L.a[i] = -tl + B[6 + i];
            L.a[tq + i] = -tl - B[6 + i];
        }
        lqp_normalize(L.A, L.a);
        lqp_normalize(L.B, L.b);
        L.v.set_zero(L.nineq, 1);
        L.w.set_zero(6, 1);
        L.y.set_zero(nvar, 1);
        Mat qdd = A_inv * B;
        for (int i = 0; i < acc; i++) L.y[i] = -qdd[i];
        L.status = 1;
        lqp.push_back(L);
    }
    // Priority 2: contact rows J_C qdd = 0 (equality); friction/ZMP rows on Fc and |qdd_joint| <= 5 (inequality)  :4383-4407
    {
        LqpLevel L;
        const int cs = contact_links * 10;
        L.nineq = cs + 2 * tq;
        L.neq = con;
        L.A.set_zero(L.nineq, nvar);
        L.a.set_zero(L.nineq, 1);
        Mat A_const, A_rot;
        int csz;
        contact_const(A_const, A_rot, csz);
        Mat C = -(A_const * A_rot);  // getContactConstraintMatrix, :480-513
        for (int j = 0; j < con; j++)
            for (int i = 0; i < cs; i++) L.A(i, acc + j) = C(i, j);
        for (int i = 0; i < tq; i++) {
            L.A(cs + i, 6 + i) = 1.0;
            L.A(cs + tq + i, 6 + i) = -1.0;
            L.a[cs + i] = -5.0;
            L.a[cs + tq + i] = -5.0;
        }
        L.B.set_zero(con, nvar);
        L.b.set_zero(con, 1);
        for (int j = 0; j < acc; j++)
            for (int i = 0; i < con; i++) L.B(i, j) = J_C(i, j);
        L.cost = true;
        lqp_normalize(L.A, L.a);
        lqp_normalize(L.B, L.b);
        lqp.push_back(L);
    }
    // Priority 3..: one level per task space, J_task qdd = f*  :4409-4425
    for (auto &t : ts) {
        if (reduced && t.noncont) continue;
        const Mat &TJ = reduced ? t.J_R : t.J;
        LqpLevel L;
        L.nineq = 0;
        L.neq = t.dof;
        L.B.set_zero(t.dof, nvar);
        L.b.set_zero(t.dof, 1);
        for (int j = 0; j < acc; j++)
            for (int i = 0; i < t.dof; i++) L.B(i, j) = TJ(i, j);
        for (int i = 0; i < t.dof; i++) L.b[i] = -t.fstar[i];
        L.cost = true;
        lqp_normalize(L.B, L.b);
        lqp.push_back(L);
    }
    lqp_prepare();
}

// HQP::prepare, reference src/dwbc_hqp.cpp:23-85
void Robot::lqp_prepare() {
    lqp[0].Z = lqp_null_space(lqp[0].B);
    lqp[0].nu = lqp[0].Z.c;
    for (size_t i = 1; i < lqp.size(); i++) {
        Mat Bn = lqp[i].B * lqp[i - 1].Z;
        Mat nb_ = lqp_null_space(Bn);
        lqp[i].Z = lqp[i - 1].Z * nb_;
        lqp[i].nu = lqp[i].Z.c;
    }
}

// HQP::solveSequentialSingle, reference src/dwbc_hqp.cpp:291-395: the QP over [u (null-space coordinates of the
// previous level); v (slack of this level's inequality rows)] is assembled as written there; HQP_Hierarch::solveOSQP
// (:583-631) passes H, g, A, ubA only (lbA = -inf, no variable bounds), so that is the problem solved here.
int Robot::lqp_solve_level(int i) {
    LqpLevel &L = lqp[i];
    const LqpLevel &P = lqp[i - 1];
    const int nu = P.nu, nv = L.nineq, nz = nu + nv;
    int rows = nv;
    for (int j = 0; j < i; j++) rows += lqp[j].nineq;
    L.qpH.set_zero(nz, nz);
    L.qpg.set_zero(nz, 1);
    L.qpA.set_zero(rows, nz);
    L.qpub.set_zero(rows, 1);
    if (nv > 0) {
        Mat AZ = L.A * P.Z, Ay = L.A * P.y;
        for (int r = 0; r < nv; r++) {
            for (int j = 0; j < nu; j++) L.qpA(r, j) = AZ(r, j);
            L.qpA(r, nu + r) = -1.0;
            L.qpub[r] = -Ay[r] - L.a[r];
            L.qpH(nu + r, nu + r) = 1.0;
        }
    }
    Mat E = L.B * P.Z, e = L.B * P.y + L.b;
    Mat EtE = E.t() * E, Ete = E.t() * e;
    for (int c = 0; c < nu; c++) {
        for (int r = 0; r < nu; r++) L.qpH(r, c) = EtE(r, c);
        L.qpg[c] = Ete[c];
    }
    if (L.cost) {
        Mat ZtH = P.Z.t() * lqp_H;
        Mat C = ZtH * P.Z, cg = ZtH * P.y;
        for (int c = 0; c < nu; c++) {
            for (int r = 0; r < nu; r++) L.qpH(r, c) += C(r, c);
            L.qpg[c] += cg[c];
        }
    }
    int ci = nv;
    for (int j = 0; j < i; j++) {
        if (lqp[j].nineq == 0) continue;
        Mat AZ = lqp[j].A * P.Z, Ay = lqp[j].A * P.y;
        for (int r = 0; r < lqp[j].nineq; r++) {
            for (int c = 0; c < nu; c++) L.qpA(ci + r, c) = AZ(r, c);
            L.qpub[ci + r] = -Ay[r] + lqp[j].v[r] - lqp[j].a[r];
        }
        ci += lqp[j].nineq;
    }
    // exact solve of  min 1/2 z^T (H + sigma I_u) z + g^T z  s.t.  qpA z <= qpub:
    // Cholesky of the u-block (the v-block is the identity) turns it into a least-distance problem.
    Mat M(nu, nu);
    for (int c = 0; c < nu; c++)
        for (int r = 0; r < nu; r++) M(r, c) = L.qpH(r, c) + (r == c ? LQP_SIGMA : 0.0);
    // lower Cholesky M = Lc Lc^T
    Mat Lc(nu, nu);
    for (int j = 0; j < nu; j++) {
        double s = M(j, j);
        for (int k = 0; k < j; k++) s -= Lc(j, k) * Lc(j, k);
        if (!(s > 0)) return L.status = 0;
        Lc(j, j) = std::sqrt(s);
        for (int r = j + 1; r < nu; r++) {
            double t = M(r, j);
            for (int k = 0; k < j; k++) t -= Lc(r, k) * Lc(j, k);
            Lc(r, j) = t / Lc(j, j);
        }
    }
    auto fwd = [&](std::vector<double> &x) {  // x <- Lc^{-1} x
        for (int r = 0; r < nu; r++) {
            double s = x[r];
            for (int k = 0; k < r; k++) s -= Lc(r, k) * x[k];
            x[r] = s / Lc(r, r);
        }
    };
    auto bwd = [&](std::vector<double> &x) {  // x <- Lc^{-T} x
        for (int r = nu - 1; r >= 0; r--) {
            double s = x[r];
            for (int k = r + 1; k < nu; k++) s -= Lc(k, r) * x[k];
            x[r] = s / Lc(r, r);
        }
    };
    // xhat = Lc^T u + Lc^{-1} g  =>  objective 1/2 |xhat|^2 + 1/2 |v|^2;  u = Lc^{-T} (xhat - gh)
    std::vector<double> gh(nu);
    for (int r = 0; r < nu; r++) gh[r] = L.qpg[r];
    fwd(gh);
    Mat Ah(rows, nz);
    std::vector<double> ubh(rows);
    for (int r = 0; r < rows; r++) {
        std::vector<double> row(nu);
        for (int c = 0; c < nu; c++) row[c] = L.qpA(r, c);
        fwd(row);  // row^T Lc^{-T} = (Lc^{-1} row)^T
        double s = 0;
        for (int c = 0; c < nu; c++) Ah(r, c) = row[c], s += row[c] * gh[c];
        for (int c = nu; c < nz; c++) Ah(r, c) = L.qpA(r, c);
        ubh[r] = L.qpub[r] + s;
    }
    QpResult res = qp_min_norm(Ah, ubh);
    if (!res.status) return L.status = 0;
    std::vector<double> uu(nu);
    for (int r = 0; r < nu; r++) uu[r] = res.x[r] - gh[r];
    bwd(uu);
    L.u.set_zero(nu, 1);
    for (int r = 0; r < nu; r++) L.u[r] = uu[r];
    L.y = P.y + P.Z * L.u;
    L.v.set_zero(nv, 1);
    for (int r = 0; r < nv; r++) L.v[r] = res.x[nu + r];
    L.w = L.B * L.y + L.b;
    return L.status = 1;
}

// CalcControlTorqueLQP -> HQP::solveSequential, reference src/dwbc.cpp:4432-4454, src/dwbc_hqp.cpp:397-403.
// The joint torque of the answer is the torque-limit row expression of priority 1 (src/dwbc.cpp:4355-4370):
// tau = A[6:,:] qdd + J_C^T[6:,:] Fc + B[6:].
int Robot::calc_lqp() {
    int ok = 1;
    for (size_t i = 1; i < lqp.size(); i++) {
        if (!lqp_solve_level((int)i)) {
            ok = 0;
            for (size_t k = i; k < lqp.size(); k++) lqp[k].y = lqp[i - 1].y, lqp[k].status = 0;
            break;
        }
    }
    const Mat &y = lqp.back().y;
    const Mat &A = lqp_reduced ? this->A_R : this->A;
    const Mat &B = lqp_reduced ? this->G_R : this->B;
    const Mat &J_C = lqp_reduced ? this->J_CR : this->J_C;
    const int n = lqp_reduced ? rsys : this->n, m = lqp_reduced ? rmod : this->m;
    lqp_torque.set_zero(m, 1);
    for (int i = 0; i < m; i++) {
        double s = B[6 + i];
        for (int j = 0; j < n; j++) s += A(6 + i, j) * y[j];
        for (int j = 0; j < lqp_con; j++) s += J_C(j, 6 + i) * y[n + j];
        lqp_torque[i] = s;
    }
    return ok;
}

// ConfigureLQP_R_NC, reference src/dwbc.cpp:4634-4760.  Decision vector: the nc_dof joint accelerations of the non-contact chain, no
// contact force.  q_acc = the acceleration part of the reduced LQP's final answer (hqp_.hqp_hs_.back().y_ans_.head(acc), reference
// tests/sp_test/jacc_compare.cpp:485-486): its first 6 entries are the base acceleration, its last 6 the centroidal acceleration of the
// condensed body.  Level 0 (solved, through HQP::solvefirst): J_I_nc qdd = fstar_gnc, torque rows +-200 on A_NC[nc,nc] qdd + G_NC;
// level 1: the 6D task of ts_[1] -- hard-coded in the reference (:4738-4745) -- in the pelvis-relative form Ja (f* - fstar_base),
// |qdd| <= 5.  Rows are NOT normalised here (updateConstraintMatrix is called on the raw matrices, :4677, 4755).
void Robot::configure_lqp_r_nc() {
    const int nc = nc_dof, acc = rsys;
    Mat q_acc(acc, 1);
    for (int i = 0; i < acc; i++) q_acc[i] = lqp.back().y[i];
    lqp_nc.clear();
    Mat Ann = A_NC.block(6, 6, nc, nc);
    lqp_nc_H = Ann;  // :4647
    {
        const double sc = 5.0 / Ann.norm();
        for (int i = 0; i < nc * nc; i++) lqp_nc_H[i] *= sc;
    }
    {   // virtual predecessor of level 0: the whole space, y = 0 (HQP::solvefirst works on the raw variables, src/dwbc_hqp.cpp:222-289)
        LqpLevel L;
        L.Z = Mat::identity(nc);
        L.nu = nc;
        L.y.set_zero(nc, 1);
        L.v.set_zero(0, 1);
        L.status = 1;
        lqp_nc.push_back(L);
    }
    {   // level 0  :4649-4678
        LqpLevel L;
        L.nineq = 2 * nc;
        L.neq = 6;
        L.A.set_zero(2 * nc, nc);
        L.a.set_zero(2 * nc, 1);
        L.B = J_I_nc;
        L.b.set_zero(6, 1);
        for (int i = 0; i < 6; i++) L.b[i] = -q_acc[acc - 6 + i];
        for (int i = 0; i < nc; i++) {
            for (int j = 0; j < nc; j++) L.A(i, j) = Ann(i, j), L.A(nc + i, j) = -Ann(i, j);
            L.a[i] = -200.0 + G[nc_joint[i]];
            L.a[nc + i] = -200.0 - G[nc_joint[i]];
        }
        L.cost = true;
        lqp_nc.push_back(L);
    }
    {   // level 1  :4735-4757
        LqpLevel L;
        const Task &t = ts[1];
        L.nineq = 2 * nc;
        L.neq = 6;
        M3 sk = skew3(pw[t.links[0].link] - pw[0]);
        Mat Ja = Mat::identity(6);
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) Ja(a, 3 + b) = sk.m[a][b];
        Mat df(6, 1);
        for (int i = 0; i < 6; i++) df[i] = t.fstar[i] - q_acc[i];
        Mat fl = Ja * df;
        L.B.set_zero(6, nc);
        L.b.set_zero(6, 1);
        for (int j = 0; j < nc; j++)
            for (int i = 0; i < 6; i++) L.B(i, j) = t.J(i, n - nc + j);  // J_task_.rightCols(nc_dof)
        for (int i = 0; i < 6; i++) L.b[i] = -fl[i];
        L.A.set_zero(2 * nc, nc);
        L.a.set_zero(2 * nc, 1);
        for (int i = 0; i < nc; i++) {
            L.A(i, i) = 1.0;
            L.A(nc + i, i) = -1.0;
            L.a[i] = -5.0;
            L.a[nc + i] = -5.0;
        }
        L.cost = true;
        lqp_nc.push_back(L);
    }
    // HQP::prepare: null spaces of the stacked equalities
    for (size_t i = 1; i < lqp_nc.size(); i++) {
        Mat Bn = lqp_nc[i].B * lqp_nc[i - 1].Z;
        lqp_nc[i].Z = lqp_nc[i - 1].Z * lqp_null_space(Bn);
        lqp_nc[i].nu = lqp_nc[i].Z.c;
    }
}

// CalcControlTorqueLQP_R_NC (reference src/dwbc.cpp:4479-4502): solvefirst + solveSequential, then the torque stitching of the only
// caller (reference tests/sp_test/jacc_compare.cpp:543-569): joints of the contact chain from the reduced LQP, the others from
// A_NC[nc,nc] qdd_nc + G_NC.
int Robot::calc_lqp_r_nc() {
    std::swap(lqp, lqp_nc);
    std::swap(lqp_H, lqp_nc_H);
    int ok = 1;
    for (size_t i = 1; i < lqp.size(); i++) {
        if (!lqp_solve_level((int)i)) {
            ok = 0;
            for (size_t k = i; k < lqp.size(); k++) lqp[k].y = lqp[i - 1].y, lqp[k].status = 0;
            break;
        }
    }
    std::swap(lqp, lqp_nc);
    std::swap(lqp_H, lqp_nc_H);
    const int nc = nc_dof;
    const Mat &y = lqp_nc.back().y;
    lqp_nc_torque.set_zero(nc, 1);
    for (int i = 0; i < nc; i++) {
        double s_ = G[nc_joint[i]];
        for (int j = 0; j < nc; j++) s_ += A_NC(6 + i, 6 + j) * y[j];
        lqp_nc_torque[i] = s_;
    }
    lqp_stitched.set_zero(m, 1);
    for (int i = 0; i < co_dof; i++) lqp_stitched[i] = lqp_torque[i];
    for (int i = 0; i < nc; i++) lqp_stitched[m - nc + i] = lqp_nc_torque[i];
    return ok;
}

extern "C" {

void *orc_create(int nb, const int *parent, const double *E_T, const double *r_T, const double *axis, const double *mass,
                 const double *com, const double *inertia) {
    Robot *R = new Robot();
    R->nb = nb; R->n = nb + 5; R->m = nb - 1;
    for (int i = 0; i < nb; i++) {
        R->parent.push_back(parent[i]);
        M3 E, I;
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) E.m[a][b] = E_T[i * 9 + a * 3 + b], I.m[a][b] = inertia[i * 9 + a * 3 + b];
        R->E_T.push_back(E);
        R->inertia.push_back(I);
        R->r_T.push_back({r_T[i * 3], r_T[i * 3 + 1], r_T[i * 3 + 2]});
        R->axis.push_back({axis[i * 3], axis[i * 3 + 1], axis[i * 3 + 2]});
        R->com.push_back({com[i * 3], com[i * 3 + 1], com[i * 3 + 2]});
        R->mass.push_back(mass[i]);
        R->total_mass += mass[i];
    }
    return R;
}
void orc_destroy(void *h) { delete (Robot *)h; }

int orc_add_contact(void *h, int link, const double *point, const double *normal, double lx, double ly) {
    Robot *R = (Robot *)h;
    for (auto &c : R->cc)
        if (c.link == link) return -1;  // "Contact Constraint Already Exist", reference src/dwbc.cpp:379-386
    Contact c;
    c.link = link;
    c.point = {point[0], point[1], point[2]};
    c.normal = {normal[0], normal[1], normal[2]};
    c.lx = lx; c.ly = ly;
    R->cc.push_back(c);
    return (int)R->cc.size() - 1;
}
// AddTaskSpace(heirarchy, mode, link, point): new level if heirarchy == size, else AddTaskLink
// (reference src/dwbc.cpp:553-600)
int orc_add_task(void *h, int level, int mode, int link, const double *point) {
    Robot *R = (Robot *)h;
    TaskLink tl;
    tl.mode = mode; tl.link = link;
    tl.point = {point[0], point[1], point[2]};
    tl.dof = (mode <= TASK_LINK_6D_CUSTOM_FRAME) ? 6 : 3;
    if ((int)R->ts.size() == level) {
        Task t;
        t.dof = tl.dof;
        t.links.push_back(tl);
        R->ts.push_back(t);
    } else if ((int)R->ts.size() > level) {
        R->ts[level].links.push_back(tl);
        R->ts[level].dof += tl.dof;
    } else
        return -1;
    R->ts[level].fstar.set_zero(R->ts[level].dof, 1);
    return 0;
}
int orc_add_custom_task(void *h, int level, int dof) {
    Robot *R = (Robot *)h;
    if ((int)R->ts.size() != level) return -1;
    Task t;
    t.custom = true;
    t.dof = dof;
    t.fstar.set_zero(dof, 1);
    R->ts.push_back(t);
    return 0;
}
void orc_set_torque_limit(void *h, const double *tlim) {
    Robot *R = (Robot *)h;
    R->tlim_set = true;
    R->tlim.set_zero(R->m, 1);
    for (int i = 0; i < R->m; i++) R->tlim[i] = tlim[i];
}
void orc_set_friction(void *h, int contact, double mu, double muz) {
    Robot *R = (Robot *)h;
    R->cc[contact].mu = mu; R->cc[contact].muz = muz;
}
void orc_update_kinematics(void *h, const double *q, const double *qd) { ((Robot *)h)->update_kinematics(q, qd); }
void orc_set_contact(void *h, const unsigned char *flags) { ((Robot *)h)->set_contact(flags); }
int orc_calc_contact_constraint(void *h) { return ((Robot *)h)->calc_contact_constraint(); }
void orc_set_task(void *h, int level, const double *fstar, const double *J) {
    Robot *R = (Robot *)h;
    Task &t = R->ts[level];
    t.fstar.set_zero(t.dof, 1);
    for (int i = 0; i < t.dof; i++) t.fstar[i] = fstar[i];
    if (J) {  // column-major t x n
        t.J.set_zero(t.dof, R->n);
        for (int i = 0; i < t.dof * R->n; i++) t.J[i] = J[i];
    }
}
void orc_calc_grav(void *h) { ((Robot *)h)->calc_grav(); }
void orc_update_task_space(void *h) { ((Robot *)h)->update_task_space(); }
int orc_calc_task_torque(void *h, int hqp) { return ((Robot *)h)->calc_task_torque(hqp != 0); }
int orc_calc_redistribute(void *h, int hqp) { return ((Robot *)h)->calc_redistribute(hqp != 0); }
// a9 reduced-dimension variant
void orc_reduced_dynamics(void *h) { ((Robot *)h)->reduced_dynamics(); }
int orc_reduced_contact_constraint(void *h) { return ((Robot *)h)->reduced_contact_constraint(); }
void orc_reduced_grav(void *h) { ((Robot *)h)->reduced_grav(); }
int orc_reduced_task_space(void *h) { return ((Robot *)h)->reduced_task_space(); }
int orc_reduced_task_torque(void *h, int hqp) { return ((Robot *)h)->reduced_task_torque(hqp != 0); }
int orc_reduced_redistribute(void *h, int hqp) { return ((Robot *)h)->reduced_redistribute(hqp != 0); }
void orc_set_reduced_torque_limit(void *h, const double *tlim, int count) {
    Robot *R = (Robot *)h;
    R->tlimR_set = true;
    R->tlimR.set_zero(count, 1);
    for (int i = 0; i < count; i++) R->tlimR[i] = tlim[i];
}
void orc_configure_lqp(void *h) { ((Robot *)h)->configure_lqp(false); }
void orc_configure_lqp_r(void *h) { ((Robot *)h)->configure_lqp(true); }
int orc_calc_lqp(void *h) { return ((Robot *)h)->calc_lqp(); }
int orc_lqp_levels(void *h) { return (int)((Robot *)h)->lqp.size(); }
void orc_configure_lqp_r_nc(void *h) { ((Robot *)h)->configure_lqp_r_nc(); }
int orc_calc_lqp_r_nc(void *h) { return ((Robot *)h)->calc_lqp_r_nc(); }
void orc_reduced_dims(void *h, int *out) {
    Robot *R = (Robot *)h;
    out[0] = R->co_dof; out[1] = R->nc_dof; out[2] = R->vc_dof; out[3] = R->rsys; out[4] = R->rmod;
}
// getContactForce(command_torque) -> out [contact_dof]; returns contact_dof
int orc_contact_force(void *h, const double *torque, double *out) {
    Robot *R = (Robot *)h;
    Mat t(R->m, 1);
    for (int i = 0; i < R->m; i++) t[i] = torque[i];
    Mat cf = R->get_contact_force(t);
    for (int i = 0; i < cf.r; i++) out[i] = cf[i];
    return cf.r;
}
// getZMP(contact_force [6 * ncontact, zero padded]) -> out [3]
void orc_zmp(void *h, const double *cf, double *out) {
    Robot *R = (Robot *)h;
    Mat c((int)(6 * R->cc.size()), 1);
    for (int i = 0; i < c.r; i++) c[i] = cf[i];
    V3 z = R->get_zmp(c);
    out[0] = z.x, out[1] = z.y, out[2] = z.z;
}
double orc_redis_eta(void *h) { return ((Robot *)h)->redis_eta; }

static const Mat *find_mat(Robot *R, const std::string &name, Mat &tmp) {
    auto lvl = [&](const char *prefix) -> int {
        size_t L = std::strlen(prefix);
        if (name.compare(0, L, prefix) == 0 && name.size() > L) return std::atoi(name.c_str() + L);
        return -1;
    };
    if (name == "A") return &R->A;
    if (name == "A_inv") return &R->A_inv;
    if (name == "B") return &R->B;
    if (name == "G") return &R->G;
    if (name == "CMM") return &R->CMM;
    if (name == "J_com") return &R->J_com;
    if (name == "J_C") return &R->J_C;
    if (name == "Lambda_contact") return &R->Lambda_c;
    if (name == "J_C_INV_T") return &R->J_C_INV_T;
    if (name == "N_C") return &R->N_C;
    if (name == "A_inv_N_C") return &R->A_inv_N_C;
    if (name == "W") return &R->W;
    if (name == "W_inv") return &R->W_inv;
    if (name == "V2") return &R->V2;
    if (name == "NwJw") return &R->NwJw;
    if (name == "P_C") return &R->P_C;
    if (name == "torque_grav") return &R->torque_grav;
    if (name == "torque_task") return &R->torque_task;
    if (name == "torque_contact") return &R->torque_contact;
    if (name == "cf_redis_qp") return &R->cf_redis_qp;
    if (name == "redisA") return &R->redisA;
    if (name == "redisub") return &R->redisub;
    {
        const std::pair<const char *, const Mat *> red[] = {
            {"A_NC", &R->A_NC}, {"SI_nc_b", &R->SI_nc_b}, {"SI_nc_l", &R->SI_nc_l}, {"cmm_nc", &R->cmm_nc}, {"J_I_nc", &R->J_I_nc}, {"J_R", &R->J_R},
            {"A_R_inv", &R->A_R_inv}, {"A_R", &R->A_R}, {"J_I_nc_inv_T", &R->J_I_nc_inv_T}, {"N_I_nc", &R->N_I_nc}, {"G_R", &R->G_R}, {"J_CR", &R->J_CR},
            {"Lambda_CR", &R->Lambda_CR}, {"J_CR_INV_T", &R->J_CR_INV_T}, {"N_CR", &R->N_CR}, {"A_R_inv_N_CR", &R->A_R_inv_N_CR}, {"W_R", &R->W_R},
            {"W_R_inv", &R->W_R_inv}, {"V2_R", &R->V2_R}, {"NwJw_R", &R->NwJw_R}, {"P_CR", &R->P_CR}, {"torque_grav_R", &R->torque_grav_R},
            {"J_base_R_kt", &R->J_base_R_kt}, {"torque_task_R", &R->torque_task_R}, {"torque_task_NC", &R->torque_task_NC},
            {"torque_contact_R", &R->torque_contact_R}, {"cf_redis_qp_R", &R->cf_redis_qp_R}, {"force_on_nc_r", &R->force_on_nc_r},
            {"force_on_nc_R_qp", &R->force_on_nc_R_qp}, {"nc_qp_contact", &R->nc_qp_contact}, {"nc_qpA", &R->nc_qpA}, {"nc_qpub", &R->nc_qpub}};
        for (auto &e : red)
            if (name == e.first) return e.second;
    }
    if (name == "com_pos") {
        tmp.set_zero(3, 1);
        tmp[0] = R->com_pos.x; tmp[1] = R->com_pos.y; tmp[2] = R->com_pos.z;
        return &tmp;
    }
    if (name == "link_pos") {  // 3 x nb
        tmp.set_zero(3, R->nb);
        for (int i = 0; i < R->nb; i++) tmp(0, i) = R->pw[i].x, tmp(1, i) = R->pw[i].y, tmp(2, i) = R->pw[i].z;
        return &tmp;
    }
    if (name == "link_rot") {  // 9 x nb, each column a column-major 3x3 (world <- link)
        tmp.set_zero(9, R->nb);
        for (int i = 0; i < R->nb; i++)
            for (int a = 0; a < 3; a++)
                for (int b = 0; b < 3; b++) tmp(b * 3 + a, i) = R->Rw[i].m[a][b];
        return &tmp;
    }
    if (name == "link_vel") {  // 6 x nb: [v_world; w_world] of each link origin (reference src/link.cpp:85-88)
        tmp.set_zero(6, R->nb);
        for (int i = 0; i < R->nb; i++) {
            V3 w = mul(R->Rw[i], V3{R->vel[i].v[0], R->vel[i].v[1], R->vel[i].v[2]});
            V3 v = mul(R->Rw[i], V3{R->vel[i].v[3], R->vel[i].v[4], R->vel[i].v[5]});
            tmp(0, i) = v.x; tmp(1, i) = v.y; tmp(2, i) = v.z; tmp(3, i) = w.x; tmp(4, i) = w.y; tmp(5, i) = w.z;
        }
        return &tmp;
    }
    int l;
    if (name == "lqp_H") return &R->lqp_H;
    if (name == "lqp_torque") return &R->lqp_torque;
    if (name == "lqp_nc_H") return &R->lqp_nc_H;
    if (name == "lqp_nc_torque") return &R->lqp_nc_torque;
    if (name == "lqp_stitched_torque") return &R->lqp_stitched;
    {
        const int nl = (int)R->lqp_nc.size();
        if ((l = lvl("lqpnc_y")) >= 0 && l < nl) return &R->lqp_nc[l].y;
        if ((l = lvl("lqpnc_v")) >= 0 && l < nl) return &R->lqp_nc[l].v;
        if ((l = lvl("lqpnc_w")) >= 0 && l < nl) return &R->lqp_nc[l].w;
        if ((l = lvl("lqpnc_Z")) >= 0 && l < nl) return &R->lqp_nc[l].Z;
        if ((l = lvl("lqpnc_A")) >= 0 && l < nl) return &R->lqp_nc[l].A;
        if ((l = lvl("lqpnc_a")) >= 0 && l < nl) return &R->lqp_nc[l].a;
        if ((l = lvl("lqpnc_B")) >= 0 && l < nl) return &R->lqp_nc[l].B;
        if ((l = lvl("lqpnc_b")) >= 0 && l < nl) return &R->lqp_nc[l].b;
        if ((l = lvl("lqpnc_qpH")) >= 0 && l < nl) return &R->lqp_nc[l].qpH;
        if ((l = lvl("lqpnc_qpg")) >= 0 && l < nl) return &R->lqp_nc[l].qpg;
        if ((l = lvl("lqpnc_qpA")) >= 0 && l < nl) return &R->lqp_nc[l].qpA;
        if ((l = lvl("lqpnc_qpub")) >= 0 && l < nl) return &R->lqp_nc[l].qpub;
    }
    {
        const int nl = (int)R->lqp.size();
        if ((l = lvl("lqp_y")) >= 0 && l < nl) return &R->lqp[l].y;
        if ((l = lvl("lqp_v")) >= 0 && l < nl) return &R->lqp[l].v;
        if ((l = lvl("lqp_w")) >= 0 && l < nl) return &R->lqp[l].w;
        if ((l = lvl("lqp_u")) >= 0 && l < nl) return &R->lqp[l].u;
        if ((l = lvl("lqp_Z")) >= 0 && l < nl) return &R->lqp[l].Z;
        if ((l = lvl("lqp_A")) >= 0 && l < nl) return &R->lqp[l].A;
        if ((l = lvl("lqp_a")) >= 0 && l < nl) return &R->lqp[l].a;
        if ((l = lvl("lqp_B")) >= 0 && l < nl) return &R->lqp[l].B;
        if ((l = lvl("lqp_b")) >= 0 && l < nl) return &R->lqp[l].b;
        if ((l = lvl("lqp_qpH")) >= 0 && l < nl) return &R->lqp[l].qpH;
        if ((l = lvl("lqp_qpg")) >= 0 && l < nl) return &R->lqp[l].qpg;
        if ((l = lvl("lqp_qpA")) >= 0 && l < nl) return &R->lqp[l].qpA;
        if ((l = lvl("lqp_qpub")) >= 0 && l < nl) return &R->lqp[l].qpub;
    }
    if ((l = lvl("J_task")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].J;
    if ((l = lvl("Lambda_task")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].Lambda;
    if ((l = lvl("J_kt")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].Jkt;
    if ((l = lvl("Null_task")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].Null;
    if ((l = lvl("f_star_qp")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].fstar_qp;
    if ((l = lvl("contact_qp")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].contact_qp;
    if ((l = lvl("torque_h")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].torque_h;
    if ((l = lvl("qpA")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].qpA;
    if ((l = lvl("R_J_task")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].J_R;
    if ((l = lvl("R_Lambda_task")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].Lambda_R;
    if ((l = lvl("R_J_kt")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].Jkt_R;
    if ((l = lvl("R_Null_task")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].Null_R;
    if ((l = lvl("R_torque_h")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].torque_h_R;
    if ((l = lvl("qpub")) >= 0 && l < (int)R->ts.size()) return &R->ts[l].qpub;
    return nullptr;
}

// Copies the named matrix (column-major) into out (may be NULL to query the shape). Returns
// rows*cols, or -1 for an unknown name.
int orc_get(void *h, const char *name, double *out, int *rows, int *cols) {
    Robot *R = (Robot *)h;
    Mat tmp;
    const Mat *M = find_mat(R, name, tmp);
    if (!M) return -1;
    if (rows) *rows = M->r;
    if (cols) *cols = M->c;
    if (out)
        for (int i = 0; i < M->r * M->c; i++) out[i] = M->d[i];
    return M->r * M->c;
}
// active rows of the level's task QP (level >= 0) or of the redistribution QP (level = -1)
int orc_get_active(void *h, int level, int *out, int cap) {
    Robot *R = (Robot *)h;
    const std::vector<int> &a = level < 0 ? R->redis_active : R->ts[level].active;
    for (int i = 0; i < (int)a.size() && i < cap; i++) out[i] = a[i];
    return (int)a.size();
}
int orc_get_qp_info(void *h, int level, int *status, int *prox_iters) {
    Robot *R = (Robot *)h;
    if (level < 0) {
        *status = R->redis_status;
        *prox_iters = 0;
    } else {
        *status = R->ts[level].qp_status;
        *prox_iters = R->ts[level].prox_iters;
    }
    return 0;
}

// Stand-alone access to the QP layer (tests): lexicographic QP on user data, column-major A.
int orc_lexqp(int rows, int nv, int t, const double *A, const double *ub, double *x, int *prox_iters) {
    Mat Am(rows, nv);
    for (int i = 0; i < rows * nv; i++) Am[i] = A[i];
    std::vector<double> u(ub, ub + rows);
    QpResult r = lexqp(Am, u, t);
    if (prox_iters) *prox_iters = r.prox_iters;
    if (r.status)
        for (int i = 0; i < nv; i++) x[i] = r.x[i];
    return r.status;
}

// Dense helpers exposed for unit tests of the linear algebra restatement.
void orc_cod_pinv(int rows, int cols, const double *A, double thr, double *pinv, int *rank) {
    Mat Am(rows, cols);
    for (int i = 0; i < rows * cols; i++) Am[i] = A[i];
    Cod f = cod_compute(Am, thr);
    *rank = f.rank;
    for (int i = 0; i < rows * cols; i++) pinv[i] = f.pinv[i];
}

// Whole cycle over a batch (CPU baseline + bulk parity): one state per call in reference call
// order, cold-start QP per state.  `stages`: bit0 grav, bit1 task torque, bit2 HQP, bit3 redistribute.
// fstar: concatenation over levels of [B x t_l] row-major blocks; flags: [B x ncontact].
// Outputs [B x m] row-major torques, status [B] (bit0 contact ok, bit1 task ok, bit2 redistribute ok).
// bit4: CalcContactRedistribute(hqp=false) (analytic two-foot branch) instead of the QP.  Optional outputs: contact_force
// [B x 6*ncontact] = getContactForce(tau_grav + tau_task + tau_contact) (zero padded), zmp [B x 3] = getZMP(contact_force).
// active / gray (may be null): [B][levels + 1][4] bit masks of the rows each QP reports active / weakly active (last = redistribution)
static int run_batch_impl(void *h, int B, const double *q, const double *qd, const double *fstar, const unsigned char *flags,
                          int stages, double *tau_grav, double *tau_task, double *tau_contact, int *status, int nthreads,
                          double *contact_force, double *zmp, unsigned *active, unsigned *gray) {
    Robot *R0 = (Robot *)h;
    int n = R0->n, m = R0->m, nc = (int)R0->cc.size();
    std::vector<size_t> foff;
    size_t off = 0;
    for (auto &t : R0->ts) {
        foff.push_back(off);
        off += (size_t)B * t.dof;
    }
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads <= 0) nthreads = 1;
    std::atomic<int> next(0);
    const int chunk = 16;
    auto worker = [&]() {
        Robot R = *R0;
        while (true) {
            int b0 = next.fetch_add(chunk);
            if (b0 >= B) break;
            int b1 = std::min(B, b0 + chunk);
            for (int b = b0; b < b1; b++) {
                R.update_kinematics(q + (size_t)b * (n + 1), qd + (size_t)b * n);
                R.set_contact(flags + (size_t)b * nc);
                int st = R.calc_contact_constraint() ? 1 : 0;
                for (size_t l = 0; l < R.ts.size(); l++) {
                    Task &t = R.ts[l];
                    for (int i = 0; i < t.dof; i++) t.fstar[i] = fstar[foff[l] + (size_t)b * t.dof + i];
                }
                R.calc_grav();
                if (stages & 2) st |= R.calc_task_torque((stages & 4) != 0) ? 2 : 0;
                if (stages & 8) st |= R.calc_redistribute(true) ? 4 : 0;
                else if (stages & 16) st |= R.calc_redistribute(false) ? 4 : 0;
                if (contact_force || zmp) {
                    Mat cf = R.get_contact_force(R.torque_grav + R.torque_task + R.torque_contact);
                    if (contact_force)
                        for (int i = 0; i < 6 * nc; i++) contact_force[(size_t)b * 6 * nc + i] = i < cf.r ? cf[i] : 0.0;
                    if (zmp) {
                        Mat cfp(6 * nc, 1);
                        for (int i = 0; i < cf.r && i < 6 * nc; i++) cfp[i] = cf[i];
                        V3 z = R.get_zmp(cfp);
                        zmp[(size_t)b * 3] = z.x, zmp[(size_t)b * 3 + 1] = z.y, zmp[(size_t)b * 3 + 2] = z.z;
                    }
                }
                for (int i = 0; i < m; i++) {
                    tau_grav[(size_t)b * m + i] = R.torque_grav[i];
                    tau_task[(size_t)b * m + i] = R.torque_task[i];
                    tau_contact[(size_t)b * m + i] = R.torque_contact[i];
                }
                status[b] = st;
                if (active || gray) {
                    const size_t nq = R.ts.size() + 1;
                    for (size_t l = 0; l < nq; l++) {
                        const std::vector<int> &a = l < R.ts.size() ? R.ts[l].active : R.redis_active;
                        const std::vector<int> &g = l < R.ts.size() ? R.ts[l].gray : R.redis_gray;
                        unsigned wa[4] = {0, 0, 0, 0}, wg[4] = {0, 0, 0, 0};
                        const bool ran = l < R.ts.size() ? ((stages & 4) != 0) : ((stages & 8) != 0);
                        if (ran) {
                            for (int i : a)
                                if (i < 128) wa[i >> 5] |= 1u << (i & 31);
                            for (int i : g)
                                if (i < 128) wg[i >> 5] |= 1u << (i & 31);
                        }
                        for (int k = 0; k < 4; k++) {
                            if (active) active[((size_t)b * nq + l) * 4 + k] = wa[k];
                            if (gray) gray[((size_t)b * nq + l) * 4 + k] = wg[k];
                        }
                    }
                }
            }
        }
    };
    if (nthreads == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; i++) th.emplace_back(worker);
        for (auto &t : th) t.join();
    }
    return 0;
}
int orc_run_batch(void *h, int B, const double *q, const double *qd, const double *fstar, const unsigned char *flags,
                  int stages, double *tau_grav, double *tau_task, double *tau_contact, int *status, int nthreads,
                  double *contact_force, double *zmp) {
    return run_batch_impl(h, B, q, qd, fstar, flags, stages, tau_grav, tau_task, tau_contact, status, nthreads, contact_force, zmp, nullptr, nullptr);
}
int orc_run_batch_sets(void *h, int B, const double *q, const double *qd, const double *fstar, const unsigned char *flags,
                       int stages, double *tau_grav, double *tau_task, double *tau_contact, int *status, int nthreads,
                       unsigned *active, unsigned *gray) {
    return run_batch_impl(h, B, q, qd, fstar, flags, stages, tau_grav, tau_task, tau_contact, status, nthreads, nullptr, nullptr, active, gray);
}

int orc_num_threads() {
    int n = (int)std::thread::hardware_concurrency();
    return n > 0 ? n : 1;
}
}

extern "C" void orc_nnls(int m, int n, const double *E, const double *f, double *x) {
    orc::Mat Em(m, n);
    for (int i = 0; i < m * n; i++) Em[i] = E[i];
    std::vector<double> fv(f, f + m), xv;
    orc::nnls(Em, fv, xv);
    for (int i = 0; i < n; i++) x[i] = xv[i];
}
