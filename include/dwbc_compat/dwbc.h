// dwbc.h (compat) -- a single-state DWBC::RobotData with the reference's method names, argument lists and public fields
// (reference include/dwbc.h:59-430), implemented as a batch of ONE over the C ABI of libdwbc_b200 (include/dwbc_b200.h).
// A controller written against the reference,
//     rd_.LoadModelData(urdf, true, false); rd_.UpdateKinematics(q, qdot, qddot); rd_.AddContactConstraint(...); rd_.AddTaskSpace(...);
//     rd_.SetContact(true, true); rd_.CalcContactConstraint(); rd_.SetTaskSpace(0, fstar); rd_.CalcGravCompensation();
//     rd_.CalcTaskControlTorque(true); rd_.CalcContactRedistribute(true);          (reference tests/dwbc_test.cpp:29-131)
// compiles against this header unchanged and reads the same public members (J_C, A_inv_, Lambda_contact, J_C_INV_T, N_C, W, W_inv,
// NwJw, torque_grav_, torque_task_, torque_contact_, ts_[i].f_star_qp_ ...).  Every method runs the CUDA kernels (there is no CPU
// path); the public members are refreshed from the device after the call that produces them.  This class is the drop-in for ONE
// robot; the batched classes (DWBC::RobotDataBatch in dwbc_b200.hpp, dwbc_multi_*) are where the throughput is.
//
// Matrix types: Eigen's when <Eigen/Dense> is on the include path, otherwise the minimal column-major stand-in below (the build
// image of this repository has no Eigen).  No CUDA header is needed to compile against this file.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../dwbc_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Dense>) && !defined(DWBC_COMPAT_NO_EIGEN)
#include <Eigen/Dense>
#define DWBC_COMPAT_EIGEN 1
#endif
#endif

#ifndef DWBC_COMPAT_EIGEN
namespace dwbc_mini {
// Just enough of Eigen's dense interface for the reference's call sequences: column-major, dynamic size, comma initialiser.
class MatrixXd {
public:
    typedef long Index;
    typedef double Scalar;
    MatrixXd() {}
    MatrixXd(Index r, Index c) : r_(r), c_(c), d_((size_t)(r * c), 0.0) {}
    static MatrixXd Zero(Index r, Index c) { return MatrixXd(r, c); }
    static MatrixXd Identity(Index r, Index c) {
        MatrixXd m(r, c);
        for (Index i = 0; i < r && i < c; i++) m(i, i) = 1.0;
        return m;
    }
    void resize(Index r, Index c) { r_ = r, c_ = c, d_.assign((size_t)(r * c), 0.0); }
    void setZero() { std::fill(d_.begin(), d_.end(), 0.0); }
    void setZero(Index r, Index c) { resize(r, c); }
    void setConstant(double v) { std::fill(d_.begin(), d_.end(), v); }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Index size() const { return r_ * c_; }
    double *data() { return d_.data(); }
    const double *data() const { return d_.data(); }
    double &operator()(Index i, Index j) { return d_[(size_t)(i + j * r_)]; }
    double operator()(Index i, Index j) const { return d_[(size_t)(i + j * r_)]; }
    double norm() const {
        double s = 0;
        for (double v : d_) s += v * v;
        return std::sqrt(s);
    }
    bool hasNaN() const {
        for (double v : d_)
            if (v != v) return true;
        return false;
    }
    MatrixXd operator-(const MatrixXd &o) const {
        MatrixXd m(r_, c_);
        for (size_t i = 0; i < d_.size(); i++) m.d_[i] = d_[i] - o.d_[i];
        return m;
    }
    MatrixXd operator+(const MatrixXd &o) const {
        MatrixXd m(r_, c_);
        for (size_t i = 0; i < d_.size(); i++) m.d_[i] = d_[i] + o.d_[i];
        return m;
    }

protected:
    Index r_ = 0, c_ = 0;
    std::vector<double> d_;
};
class VectorXd : public MatrixXd {
public:
    VectorXd() {}
    explicit VectorXd(Index n) : MatrixXd(n, 1) {}
    VectorXd(const MatrixXd &m) : MatrixXd(m) {}
    static VectorXd Zero(Index n) { return VectorXd(n); }
    static VectorXd Constant(Index n, double v) {
        VectorXd x(n);
        x.MatrixXd::setConstant(v);
        return x;
    }
    void resize(Index n) { MatrixXd::resize(n, 1); }
    void setZero(Index n) { MatrixXd::resize(n, 1); }
    void setZero() { MatrixXd::setZero(); }
    void setConstant(Index n, double v) {
        MatrixXd::resize(n, 1);
        MatrixXd::setConstant(v);
    }
    double &operator()(Index i) { return d_[(size_t)i]; }
    double operator()(Index i) const { return d_[(size_t)i]; }
    double &operator[](Index i) { return d_[(size_t)i]; }
    double operator[](Index i) const { return d_[(size_t)i]; }
    // q << a, b, c, ...;
    struct Comma {
        VectorXd &v;
        Index k;
        Comma &operator,(double x) {
            v(k++) = x;
            return *this;
        }
    };
    Comma operator<<(double x) {
        (*this)(0) = x;
        return Comma{*this, 1};
    }
};
class Vector3d : public VectorXd {
public:
    Vector3d() : VectorXd(3) {}
    Vector3d(double x, double y, double z) : VectorXd(3) { d_[0] = x, d_[1] = y, d_[2] = z; }
    static Vector3d Zero() { return Vector3d(); }
};
class Matrix3d : public MatrixXd {
public:
    Matrix3d() : MatrixXd(3, 3) {}
};
}  // namespace dwbc_mini
namespace Eigen {
using dwbc_mini::Matrix3d;
using dwbc_mini::MatrixXd;
using dwbc_mini::Vector3d;
using dwbc_mini::VectorXd;
}  // namespace Eigen
#endif

using namespace Eigen;

// reference include/dwbc_contact_constraint.h:19-25, include/dwbc_task.h:23-33 (global enums, as in the reference)
enum CONTACT_TYPE { CONTACT_6D, CONTACT_POINT, CONTACT_LINK, CONTACT_LINE };
enum LINK_MODE {
    TASK_LINK_6D,
    TASK_LINK_6D_COM_FRAME,
    TASK_LINK_6D_CUSTOM_FRAME,
    TASK_LINK_POSITION,
    TASK_LINK_POSITION_COM_FRAME,
    TASK_LINK_POSITION_CUSTOM_FRAME,
    TASK_LINK_ROTATION,
    TASK_LINK_ROTATION_CUSTOM_FRAME
};
enum TASK_MODE { TASK_LINK, TASK_CUSTOM = 100 };

namespace DWBC {

// per-link public fields the path fills (reference include/dwbc_link.h: xpos, rotm)
struct Link {
    Vector3d xpos;
    Matrix3d rotm;
};
// per-level public fields (reference include/dwbc_task.h)
struct TaskSpace {
    int task_dof_ = 0, heirarchy_ = 0;
    VectorXd f_star_, f_star_qp_, contact_qp_, torque_h_;
    MatrixXd J_task_, Lambda_task_, J_kt_, Null_task_;
};

class RobotData {
public:
    // ---- public fields of the reference (include/dwbc.h:67-232), the ones on the hot path ----
    int system_dof_ = 0, model_dof_ = 0, link_num_ = 0, contact_dof_ = 0, contact_link_num_ = 0;
    double total_mass_ = 0.0;
    bool is_floating_ = true;
    VectorXd q_system_, q_dot_system_, q_ddot_system_, torque_limit_;
    MatrixXd A_, A_inv_, J_C, Lambda_contact, J_C_INV_T, N_C, A_inv_N_C, W, W_inv, V2, NwJw;
    VectorXd G_, B_, P_C, torque_grav_, torque_task_, torque_contact_, cf_redis_qp_;
    Vector3d com_pos;
    std::vector<Link> link_;
    std::vector<TaskSpace> ts_;
    bool torque_limit_set_ = false;

    RobotData() = default;
    RobotData(const RobotData &) = delete;  // (the reference copies RobotData by value; a device handle is not copied silently -- use CopyFrom)
    RobotData &operator=(const RobotData &) = delete;
    ~RobotData() { release(); }

    // LoadModelData(urdf_path, floating, verbose)   reference include/dwbc.h:237, src/dwbc.cpp:102-123
    void LoadModelData(std::string urdf_path, bool floating, int verbose = 0) {
        (void)verbose;
        release();
        model_ = dwbc_model_load_urdf(urdf_path.c_str(), floating ? 1 : 0);
        if (!model_) fail("LoadModelData");
        is_floating_ = floating;
        dwbc_model_info(model_, &system_dof_, &model_dof_, &link_num_, &total_mass_);
        batch_ = dwbc_batch_create(model_, 1, device_);
        if (!batch_) fail("LoadModelData");
        ck(dwbc_batch_set_outputs(batch_, DWBC_OUT_ALL), "LoadModelData");
        dwbc_batch_limits(batch_, &cmax_, &rmax_, nullptr, nullptr);
        const size_t n = system_dof_, m = model_dof_;
        size_t cap = (n + 1) * 2 + n * n + 9 * link_num_ + 64;
        d_q_ = dalloc((n + 1) * sizeof(double));
        d_qd_ = dalloc(n * sizeof(double));
        d_flags_ = dalloc(16);
        d_out_ = dalloc(cap * sizeof(double));
        d_tau_ = dalloc((3 * m + 8) * sizeof(double));
        host_.assign(cap, 0.0);
        link_.assign(link_num_ + 1, Link());
        (void)m;
    }
    void SetDevice(int device) { device_ = device; }  // before LoadModelData

    // UpdateKinematics(q, qdot, qddot, update_kinematics)   :251, src/dwbc.cpp:279-371 (the flag's false branch is broken in the reference)
    void UpdateKinematics(const VectorXd q_virtual, const VectorXd q_dot_virtual, const VectorXd q_ddot_virtual, bool update_kinematics = true) {
        (void)update_kinematics;
        if ((int)q_virtual.size() != system_dof_ + 1 || (int)q_dot_virtual.size() != system_dof_) {
            std::fprintf(stderr, "ERROR : q size is not matched with system_dof_\n");  // reference src/dwbc.cpp:282-296 prints and returns
            return;
        }
        q_system_ = q_virtual, q_dot_system_ = q_dot_virtual, q_ddot_system_ = q_ddot_virtual;
        up(d_q_, q_virtual.data(), (system_dof_ + 1) * sizeof(double));
        up(d_qd_, q_dot_virtual.data(), system_dof_ * sizeof(double));
        ck(dwbc_batch_update_kinematics(batch_, (const double *)d_q_, (const double *)d_qd_, nullptr), "UpdateKinematics");
        have_kin_ = true;
        get(DWBC_MAT_A, 0, A_);
        get(DWBC_MAT_A_INV, 0, A_inv_);
        getv(DWBC_MAT_G, 0, G_);
        getv(DWBC_MAT_B, 0, B_);
        VectorXd c;
        getv(DWBC_MAT_COM, 0, c);
        for (int k = 0; k < 3; k++) com_pos(k) = c(k);
        MatrixXd P, R;
        get(DWBC_MAT_LINK_POS, 0, P);
        get(DWBC_MAT_LINK_ROT, 0, R);
        for (int i = 0; i < link_num_; i++) {
            for (int k = 0; k < 3; k++) link_[i].xpos(k) = P(k, i);
            for (int k = 0; k < 9; k++) link_[i].rotm(k % 3, k / 3) = R(k, i);
        }
        for (int k = 0; k < 3; k++) link_[link_num_].xpos(k) = c(k);  // the "COM" pseudo link
        if (ncontact_ > 0 && have_flags_) push_contact();  // the reference refreshes the contact Jacobians of the active contacts here (:309-312)
    }
    // AddContactConstraint   :259,265, src/dwbc.cpp:377-411
    void AddContactConstraint(int link_number, int contact_type, Vector3d contact_point, Vector3d contact_vector, double contact_x = 0, double contact_y = 0,
                              bool verbose = false) {
        (void)verbose;
        const double p[3] = {contact_point(0), contact_point(1), contact_point(2)}, nrm[3] = {contact_vector(0), contact_vector(1), contact_vector(2)};
        if (dwbc_batch_add_contact(batch_, link_number, contact_type, p, nrm, contact_x, contact_y) < 0) {
            std::fprintf(stderr, "%s\n", dwbc_last_error());  // "Contact Constraint Already Exist for Link" etc.: printed, not thrown (src/dwbc.cpp:381-386)
            return;
        }
        ncontact_++;
    }
    void AddContactConstraint(const char *link_name, int contact_type, Vector3d contact_point, Vector3d contact_vector, double contact_x = 0,
                              double contact_y = 0, bool verbose = false) {
        const int id = dwbc_model_link_id(model_, link_name);
        if (id < 0) {
            std::fprintf(stderr, "Link Name is Wrong : %s\n", link_name);  // src/dwbc.cpp:407-411
            return;
        }
        AddContactConstraint(id, contact_type, contact_point, contact_vector, contact_x, contact_y, verbose);
    }
    // SetContact(flags...)   :291, include/dwbc.h:432-474
    template <typename... Types>
    void SetContact(Types... flags) {
        const bool f[] = {static_cast<bool>(flags)...};
        const int nf = (int)sizeof...(flags);
        std::memset(flags_, 0, sizeof(flags_));
        contact_link_num_ = 0;
        for (int i = 0; i < nf && i < ncontact_ && i < 16; i++) flags_[i] = f[i] ? 1 : 0, contact_link_num_ += f[i] ? 1 : 0;
        contact_dof_ = 6 * contact_link_num_;
        have_flags_ = true;
        push_contact();
    }
    // CalcContactConstraint()   :280, src/dwbc.cpp:456-478
    int CalcContactConstraint() {
        ck(dwbc_batch_calc_contact_constraint(batch_, nullptr), "CalcContactConstraint");
        const int c = contact_dof_, n = system_dof_, m = model_dof_, r = c > 6 ? c - 6 : 0;
        MatrixXd t;
        get(DWBC_MAT_LAMBDA_C, 0, t), crop(t, c, c, Lambda_contact);
        get(DWBC_MAT_J_C_INV_T, 0, t), crop(t, c, n, J_C_INV_T);
        get(DWBC_MAT_N_C, 0, N_C);
        get(DWBC_MAT_A_INV_N_C, 0, A_inv_N_C);
        get(DWBC_MAT_W, 0, W);
        get(DWBC_MAT_W_INV, 0, W_inv);
        get(DWBC_MAT_V2, 0, t), crop(t, r, m, V2);
        get(DWBC_MAT_NWJW, 0, t), crop(t, m, r, NwJw);
        VectorXd pc;
        getv(DWBC_MAT_P_C, 0, pc);
        P_C.resize(c);
        for (int i = 0; i < c; i++) P_C(i) = pc(i);
        return status() & DWBC_ST_CONTACT_OK ? 1 : 0;
    }
    // AddTaskSpace   :318-320, src/dwbc.cpp:522-600
    void AddTaskSpace(int heirarchy, int task_mode, int task_dof) {  // TASK_CUSTOM
        if (task_mode != TASK_CUSTOM) {
            std::fprintf(stderr, "AddTaskSpace(heirarchy, TASK_CUSTOM, task_dof) expected\n");
            return;
        }
        if (dwbc_batch_add_custom_task(batch_, heirarchy, task_dof) < 0) {
            std::fprintf(stderr, "%s\n", dwbc_last_error());
            return;
        }
        sync_levels();
    }
    void AddTaskSpace(int heirarchy, int task_mode, int link_number, Vector3d task_point, bool verbose = false) {
        (void)verbose;
        const double p[3] = {task_point(0), task_point(1), task_point(2)};
        if (dwbc_batch_add_task(batch_, heirarchy, task_mode, link_number, p) < 0) {
            std::fprintf(stderr, "%s\n", dwbc_last_error());
            return;
        }
        const int dof = task_mode <= TASK_LINK_6D_CUSTOM_FRAME ? 6 : 3;
        if (heirarchy == (int)level_dof_.size()) level_dof_.push_back(dof);
        else level_dof_[heirarchy] += dof;
        sync_levels();
    }
    void AddTaskSpace(int heirarchy, int task_mode, const char *link_name, Vector3d task_point, bool verbose = false) {
        const int id = dwbc_model_link_id(model_, link_name);
        if (id < 0) {
            std::fprintf(stderr, "Link Name is Wrong : %s\n", link_name);
            return;
        }
        AddTaskSpace(heirarchy, task_mode, id, task_point, verbose);
    }
    void ClearTaskSpace() {
        dwbc_batch_clear_tasks(batch_);
        level_dof_.clear();
        sync_levels();
    }
    // SetTaskSpace(heirarchy, f_star, J_task)   :333, src/dwbc.cpp:662-678
    void SetTaskSpace(int heirarchy, const VectorXd &f_star, const MatrixXd &J_task = MatrixXd::Zero(1, 1)) {
        if (heirarchy < 0 || heirarchy >= (int)ts_.size()) {
            std::fprintf(stderr, "ERROR : task space size overflow\n");  // src/dwbc.cpp:664-671
            return;
        }
        TaskSpace &t = ts_[heirarchy];
        t.f_star_ = f_star;
        if (!d_f_[heirarchy]) d_f_[heirarchy] = dalloc(sizeof(double) * 64);
        up(d_f_[heirarchy], f_star.data(), sizeof(double) * (size_t)f_star.size());
        const double *jd = nullptr;
        if (J_task.rows() == t.task_dof_ && J_task.cols() == system_dof_) {
            if (!d_J_[heirarchy]) d_J_[heirarchy] = dalloc(sizeof(double) * 64 * (size_t)system_dof_);
            up(d_J_[heirarchy], J_task.data(), sizeof(double) * (size_t)J_task.size());
            jd = (const double *)d_J_[heirarchy];
        }
        ck(dwbc_batch_set_task(batch_, heirarchy, (const double *)d_f_[heirarchy], jd, 0, nullptr), "SetTaskSpace");
    }
    // SetTorqueLimit   :249, src/dwbc.cpp:254-258
    void SetTorqueLimit(const VectorXd &torque_limit) {
        if ((int)torque_limit.size() != model_dof_) {
            std::fprintf(stderr, "ERROR : torque limit size is not matched with model_dof_\n");
            return;
        }
        torque_limit_ = torque_limit;
        torque_limit_set_ = true;
        ck(dwbc_batch_set_torque_limit(batch_, torque_limit.data()), "SetTorqueLimit");
    }
    // CalcGravCompensation()   :246, src/dwbc.cpp:875-889
    VectorXd CalcGravCompensation() {
        solve(DWBC_SOLVE_GRAV);
        return torque_grav_;
    }
    // CalcTaskSpace()   src/dwbc.cpp:795-816 (done inside CalcTaskControlTorque here; kept so that call sequences compile)
    void CalcTaskSpace(bool update = true) { (void)update; }
    // CalcTaskControlTorque(hqp, init, update_task_space)   :349, src/dwbc.cpp:818-873.  A call is always a cold start (init = true).
    int CalcTaskControlTorque(bool hqp = true, bool init = true, bool update_task_space = true) {
        (void)init, (void)update_task_space;
        const int st = solve(DWBC_SOLVE_TASK | (hqp ? DWBC_SOLVE_HQP : 0));
        for (size_t l = 0; l < ts_.size(); l++) {
            TaskSpace &t = ts_[l];
            MatrixXd tmp;
            get(DWBC_MAT_J_TASK, (int)l, t.J_task_);
            get(DWBC_MAT_LAMBDA_TASK, (int)l, t.Lambda_task_);
            get(DWBC_MAT_J_KT, (int)l, t.J_kt_);
            if (l + 1 < ts_.size()) get(DWBC_MAT_NULL_TASK, (int)l, t.Null_task_);
            getv(DWBC_MAT_TORQUE_H, (int)l, t.torque_h_);
            if (hqp) {
                getv(DWBC_MAT_FSTAR_QP, (int)l, t.f_star_qp_);
                VectorXd cq;
                getv(DWBC_MAT_CONTACT_QP, (int)l, cq);
                const int r = contact_dof_ > 6 ? contact_dof_ - 6 : 0;
                t.contact_qp_.resize(r);
                for (int i = 0; i < r; i++) t.contact_qp_(i) = cq(i);
            }
        }
        return st & DWBC_ST_TASK_OK ? 1 : 0;
    }
    // CalcContactRedistribute(hqp, init)   :298, src/dwbc.cpp:1372-1620
    int CalcContactRedistribute(bool hqp = true, bool init = true) {
        (void)init;
        const int st = solve(hqp ? DWBC_SOLVE_REDIST : DWBC_SOLVE_REDIST_ANALYTIC);
        VectorXd cf;
        getv(DWBC_MAT_CF_REDIS, 0, cf);
        const int r = contact_dof_ > 6 ? contact_dof_ - 6 : 0;
        cf_redis_qp_.resize(r);
        for (int i = 0; i < r; i++) cf_redis_qp_(i) = cf(i);
        return st & DWBC_ST_REDIST_OK ? 1 : 0;
    }
    // getContactForce(command_torque)   :303, src/dwbc.cpp:891-896
    VectorXd getContactForce(const VectorXd &command_torque) {
        up(d_tau_, command_torque.data(), sizeof(double) * (size_t)model_dof_);
        ck(dwbc_batch_get_contact_force(batch_, (const double *)d_tau_, (double *)d_out_, nullptr), "getContactForce");
        dn(host_.data(), d_out_, sizeof(double) * (size_t)cmax_);
        VectorXd cf(contact_dof_);
        for (int i = 0; i < contact_dof_; i++) cf(i) = host_[i];
        return cf;
    }
    int getLinkID(std::string name) { return dwbc_model_link_id(model_, name.c_str()); }
    dwbc_batch_t *handle() { return batch_; }

private:
    dwbc_model_t *model_ = nullptr;
    dwbc_batch_t *batch_ = nullptr;
    int device_ = 0, cmax_ = 12, rmax_ = 6, ncontact_ = 0;
    bool have_kin_ = false, have_flags_ = false;
    uint8_t flags_[16] = {0};
    void *d_q_ = nullptr, *d_qd_ = nullptr, *d_flags_ = nullptr, *d_out_ = nullptr, *d_tau_ = nullptr;
    void *d_f_[8] = {nullptr}, *d_J_[8] = {nullptr};
    std::vector<void *> dev_;
    std::vector<double> host_;
    std::vector<int> level_dof_;

    [[noreturn]] void fail(const char *who) const { throw std::runtime_error(std::string(who) + ": " + dwbc_last_error()); }
    int ck(int rc, const char *who) const {
        if (rc < 0) fail(who);
        return rc;
    }
    void *dalloc(size_t bytes) {
        void *p = dwbc_device_alloc(device_, bytes);
        if (!p) fail("device allocation");
        dev_.push_back(p);
        return p;
    }
    void up(void *d, const void *h, size_t bytes) { ck(dwbc_memcpy_h2d(device_, d, h, bytes, nullptr), "copy to device"); }
    void dn(void *h, const void *d, size_t bytes) { ck(dwbc_memcpy_d2h(device_, h, d, bytes, nullptr), "copy from device"); }
    void release() {
        for (void *p : dev_) dwbc_device_free(device_, p);
        dev_.clear();
        for (int i = 0; i < 8; i++) d_f_[i] = d_J_[i] = nullptr;
        if (batch_) dwbc_batch_destroy(batch_);
        if (model_) dwbc_model_destroy(model_);
        batch_ = nullptr, model_ = nullptr;
        ncontact_ = 0, have_kin_ = have_flags_ = false;
        level_dof_.clear(), ts_.clear();
    }
    void sync_levels() {
        // custom levels report their dof through the handle; keep ts_ in step with the hierarchy
        int r = 0, c = 0;
        std::vector<TaskSpace> old;
        old.swap(ts_);
        for (int l = 0; l < 8; l++) {
            if (dwbc_batch_get_matrix(batch_, DWBC_MAT_FSTAR_QP, l, nullptr, &r, &c, nullptr) != DWBC_OK) {
                // levels whose buffers do not exist yet still answer the size query as long as the level exists
                if (std::string(dwbc_last_error()).find("level out of range") != std::string::npos) break;
                r = l < (int)level_dof_.size() ? level_dof_[l] : 0;
                if (r == 0) break;
            }
            TaskSpace t = l < (int)old.size() ? old[l] : TaskSpace();
            t.task_dof_ = r;
            t.heirarchy_ = l;
            ts_.push_back(t);
        }
    }
    void push_contact() {
        if (!have_kin_) return;
        up(d_flags_, flags_, 16);
        ck(dwbc_batch_set_contact(batch_, (const uint8_t *)d_flags_, nullptr), "SetContact");
        MatrixXd t;
        get(DWBC_MAT_J_C, 0, t);
        crop(t, contact_dof_, system_dof_, J_C);
    }
    int solve(int stages) {
        double *tau = (double *)d_tau_;
        const size_t m = model_dof_;
        ck(dwbc_batch_solve(batch_, stages, tau, tau + m, tau + 2 * m, (int32_t *)(tau + 3 * m), nullptr), "solve");
        std::vector<double> h(3 * m + 1);
        dn(h.data(), d_tau_, sizeof(double) * (3 * m + 1));
        torque_grav_.resize(m), torque_task_.resize(m), torque_contact_.resize(m);
        for (size_t i = 0; i < m; i++) torque_grav_(i) = h[i], torque_task_(i) = h[m + i], torque_contact_(i) = h[2 * m + i];
        int32_t st;
        std::memcpy(&st, &h[3 * m], sizeof(st));
        status_ = st;
        return st;
    }
    int status_ = 0;
    int status() {
        int32_t st = 0;
        ck(dwbc_batch_solve(batch_, 0, nullptr, nullptr, nullptr, (int32_t *)d_tau_, nullptr), "status");
        dn(&st, d_tau_, sizeof(st));
        return st;
    }
    void get(int which, int level, MatrixXd &out) {
        int r = 0, c = 0;
        ck(dwbc_batch_get_matrix(batch_, which, level, (double *)d_out_, &r, &c, nullptr), "get_matrix");
        if ((size_t)r * c > host_.size()) host_.resize((size_t)r * c);
        dn(host_.data(), d_out_, sizeof(double) * (size_t)r * c);
        out.resize(r, c);
        std::memcpy(out.data(), host_.data(), sizeof(double) * (size_t)r * c);
    }
    void getv(int which, int level, VectorXd &out) {
        MatrixXd t;
        get(which, level, t);
        out.resize(t.rows() * t.cols());
        std::memcpy(out.data(), t.data(), sizeof(double) * (size_t)t.size());
    }
    static void crop(const MatrixXd &src, int rows, int cols, MatrixXd &dst) {  // the workspace pads to cmax / rmax; the reference's fields have the state's sizes
        dst.resize(rows, cols);
        for (int j = 0; j < cols; j++)
            for (int i = 0; i < rows; i++) dst(i, j) = src(i, j);
    }
};

}  // namespace DWBC
