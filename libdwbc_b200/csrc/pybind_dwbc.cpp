// Python module `dwbc`: the reference's pybind bridge (reference src/pybind/rl_wbc_bridge.cpp:148-158) over the C ABI of libdwbc_b200.
//
//   dwbc.RlWBCBridge(env_id)       the reference class, signature for signature: reflectAction(action), UpdateKinematics(qpos, qvel, qacc),
//                                  SetContact(left, right), SetTaskSpace(heirarchy, f_star), CalcTorque(), getTorqueCommand() -> list[float],
//                                  Reset().  One MuJoCo environment; every call runs the CUDA kernels on a batch of one through the
//                                  header-compatible DWBC::RobotData (include/dwbc_compat/dwbc.h).
//   dwbc.RlWBCBridgeBatch(num_envs, device = 0)   the same calls for num_envs environments at once: numpy [B, ...] in, float32 [B, 33] out.
//
// Host code only: g++ + pybind11, linked against libdwbc_b200.so (no CUDA header needed).  The URDF is $DWBC_URDF, else the path compiled in
// as DWBC_URDF_PATH (the reference compiles URDF_DIR in the same way, src/pybind/rl_wbc_bridge.cpp:10-13).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/dwbc_compat/dwbc.h"

namespace py = pybind11;

#ifndef DWBC_URDF_PATH
#define DWBC_URDF_PATH "dyros_tocabi.urdf"
#endif

static std::string urdf_path() {
    const char *e = std::getenv("DWBC_URDF");
    return e && *e ? std::string(e) : std::string(DWBC_URDF_PATH);
}
// initial posture of the bridge, reference src/pybind/rl_wbc_bridge.cpp:24-30
static const double Q_INIT[40] = {0, 0, 0.92983, 0, 0, 0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0, 0, 0,
                                  0.3, 0.3, 1.5, -1.27, -1, 0, -1, 0, 0, 0, -0.3, -0.3, -1.5, 1.27, 1, 0, 1, 0, 1};

class RlWBCBridge {
public:
    explicit RlWBCBridge(int env_id) : env_id_(env_id) {
        rd_.LoadModelData(urdf_path(), true, false);
        q.setZero(rd_.system_dof_ + 1);
        qdot.setZero(rd_.system_dof_);
        qddot.setZero(rd_.system_dof_);
        ctrl_command_.resize(rd_.model_dof_);
        for (int i = 0; i < 40 && i < rd_.system_dof_ + 1; i++) q(i) = Q_INIT[i];
        rd_.UpdateKinematics(q, qdot, qddot);
        rd_.AddContactConstraint(6, CONTACT_TYPE::CONTACT_6D, Vector3d(0.03, 0, -0.1585), Vector3d(0, 0, 1), 0.15, 0.075);
        rd_.AddContactConstraint(12, CONTACT_TYPE::CONTACT_6D, Vector3d(0.03, 0, -0.1585), Vector3d(0, 0, 1), 0.15, 0.075);
        rd_.AddTaskSpace(0, TASK_LINK_6D, 0, Vector3d::Zero());
        rd_.AddTaskSpace(1, TASK_LINK_ROTATION, 15, Vector3d::Zero());
        VectorXd tlim;
        tlim.setConstant(rd_.model_dof_, 300);
        rd_.SetTorqueLimit(tlim);
    }
    void reflectAction(py::array_t<double>) {}  // the reference's body is commented out (src/pybind/rl_wbc_bridge.cpp:52-80)
    void UpdateKinematics(py::array_t<double, py::array::c_style | py::array::forcecast> qpos, py::array_t<double, py::array::c_style | py::array::forcecast> qvel,
                          py::array_t<double, py::array::c_style | py::array::forcecast> qacc) {
        const double *qp = qpos.data(), *qv = qvel.data(), *qa = qacc.data();
        if (qpos.size() < rd_.system_dof_ + 1 || qvel.size() < rd_.system_dof_) throw std::runtime_error("UpdateKinematics: qpos / qvel too short");
        q(0) = qp[0], q(1) = qp[1], q(2) = qp[2];
        q(3) = qp[4], q(4) = qp[5], q(5) = qp[6];
        q(rd_.system_dof_) = qp[3];  // MuJoCo: qw first
        for (int j = 0; j < rd_.model_dof_; j++) q(j + 6) = qp[j + 7];
        for (int j = 0; j < rd_.system_dof_; j++) qdot(j) = qv[j], qddot(j) = qacc.size() >= rd_.system_dof_ ? qa[j] : 0.0;
        rd_.UpdateKinematics(q, qdot, qddot);
    }
    void SetContact(bool left, bool right) {
        rd_.SetContact(left, right);
        rd_.CalcContactConstraint();
        rd_.CalcTaskSpace();
    }
    void SetTaskSpace(int heirarchy, py::array_t<double, py::array::c_style | py::array::forcecast> f_star) {
        VectorXd f((long)f_star.size());
        for (long i = 0; i < (long)f_star.size(); i++) f(i) = f_star.data()[i];
        rd_.SetTaskSpace(heirarchy, f);
    }
    void CalcTorque() {
        task_init = false;  // the reference passes this flag as `hqp` (src/pybind/rl_wbc_bridge.cpp:123-129): closed form + analytic redistribution
        rd_.CalcGravCompensation();
        rd_.CalcTaskControlTorque(task_init);
        rd_.CalcContactRedistribute(task_init);
    }
    std::vector<float> getTorqueCommand() {
        for (int i = 0; i < rd_.model_dof_; i++) ctrl_command_[i] = (float)(rd_.torque_grav_(i) + rd_.torque_task_(i) + rd_.torque_contact_(i));
        return ctrl_command_;
    }
    void Reset() { task_init = true; }

private:
    int env_id_;
    DWBC::RobotData rd_;
    VectorXd q, qdot, qddot;
    std::vector<float> ctrl_command_;
    bool task_init = true;
};

// num_envs environments at once (the batched entry point north_star adds), same call names; host numpy arrays, staged through the C ABI.
class RlWBCBridgeBatch {
public:
    RlWBCBridgeBatch(int num_envs, int device) : B_(num_envs), dev_(device) {
        model_ = dwbc_model_load_urdf(urdf_path().c_str(), 1);
        if (!model_) fail("RlWBCBridgeBatch");
        dwbc_model_info(model_, &n_, &m_, nullptr, nullptr);
        b_ = dwbc_batch_create(model_, B_, dev_);
        if (!b_) fail("RlWBCBridgeBatch");
        ck(dwbc_batch_set_outputs(b_, DWBC_OUT_ALL));
        const double pt[3] = {0.03, 0, -0.1585}, nz[3] = {0, 0, 1}, zero[3] = {0, 0, 0};
        ck(dwbc_batch_add_contact(b_, 6, DWBC_CONTACT_6D, pt, nz, 0.15, 0.075));
        ck(dwbc_batch_add_contact(b_, 12, DWBC_CONTACT_6D, pt, nz, 0.15, 0.075));
        ck(dwbc_batch_add_task(b_, 0, DWBC_TASK_LINK_6D, 0, zero));
        ck(dwbc_batch_add_task(b_, 1, DWBC_TASK_LINK_ROTATION, 15, zero));
        std::vector<double> tl(m_, 300.0);
        ck(dwbc_batch_set_torque_limit(b_, tl.data()));
        const size_t B = B_;
        d_qpos_ = dalloc(B * (n_ + 1) * 8), d_qvel_ = dalloc(B * n_ * 8), d_flags_ = dalloc(B * 2), d_f0_ = dalloc(B * 6 * 8), d_f1_ = dalloc(B * 3 * 8);
        d_cmd_ = dalloc(B * m_ * 4), d_st_ = dalloc(B * 4);
    }
    ~RlWBCBridgeBatch() {
        for (void *p : allocs_) dwbc_device_free(dev_, p);
        if (b_) dwbc_batch_destroy(b_);
        if (model_) dwbc_model_destroy(model_);
    }
    void reflectAction(py::array_t<double>) {}
    void UpdateKinematics(py::array_t<double, py::array::c_style | py::array::forcecast> qpos, py::array_t<double, py::array::c_style | py::array::forcecast> qvel, py::object) {
        if (qpos.size() != (long)B_ * (n_ + 1) || qvel.size() != (long)B_ * n_) throw std::runtime_error("UpdateKinematics: qpos [B, 40] / qvel [B, 39] expected");
        ck(dwbc_memcpy_h2d(dev_, d_qpos_, qpos.data(), sizeof(double) * B_ * (n_ + 1), nullptr));
        ck(dwbc_memcpy_h2d(dev_, d_qvel_, qvel.data(), sizeof(double) * B_ * n_, nullptr));
        ck(dwbc_batch_update_kinematics_mujoco(b_, (const double *)d_qpos_, (const double *)d_qvel_, nullptr));
    }
    void SetContact(py::array_t<uint8_t, py::array::c_style | py::array::forcecast> left, py::array_t<uint8_t, py::array::c_style | py::array::forcecast> right) {
        std::vector<uint8_t> fl((size_t)B_ * 2);
        for (int b = 0; b < B_; b++) {
            fl[2 * b] = left.data()[left.size() == 1 ? 0 : b] ? 1 : 0;
            fl[2 * b + 1] = right.data()[right.size() == 1 ? 0 : b] ? 1 : 0;
        }
        ck(dwbc_memcpy_h2d(dev_, d_flags_, fl.data(), fl.size(), nullptr));
        ck(dwbc_batch_set_contact(b_, (const uint8_t *)d_flags_, nullptr));
        ck(dwbc_batch_calc_contact_constraint(b_, nullptr));
    }
    void SetTaskSpace(int heirarchy, py::array_t<double, py::array::c_style | py::array::forcecast> f_star) {
        const int t = heirarchy == 0 ? 6 : 3;
        if (heirarchy < 0 || heirarchy > 1 || f_star.size() != (long)B_ * t) throw std::runtime_error("SetTaskSpace: f_star [B, 6] (level 0) or [B, 3] (level 1) expected");
        void *d = heirarchy == 0 ? d_f0_ : d_f1_;
        ck(dwbc_memcpy_h2d(dev_, d, f_star.data(), sizeof(double) * B_ * t, nullptr));
        ck(dwbc_batch_set_task(b_, heirarchy, (const double *)d, nullptr, 0, nullptr));
    }
    void CalcTorque() { ck(dwbc_batch_solve(b_, DWBC_SOLVE_GRAV | DWBC_SOLVE_TASK | DWBC_SOLVE_REDIST_ANALYTIC, nullptr, nullptr, nullptr, (int32_t *)d_st_, nullptr)); }
    py::array_t<float> getTorqueCommand() {
        ck(dwbc_batch_get_torque_command_f32(b_, (float *)d_cmd_, nullptr));
        py::array_t<float> out({(py::ssize_t)B_, (py::ssize_t)m_});
        ck(dwbc_memcpy_d2h(dev_, out.mutable_data(), d_cmd_, sizeof(float) * B_ * m_, nullptr));
        return out;
    }
    void Reset() {}

private:
    int B_, dev_, n_ = 0, m_ = 0;
    dwbc_model_t *model_ = nullptr;
    dwbc_batch_t *b_ = nullptr;
    void *d_qpos_, *d_qvel_, *d_flags_, *d_f0_, *d_f1_, *d_cmd_, *d_st_;
    std::vector<void *> allocs_;
    [[noreturn]] void fail(const char *who) { throw std::runtime_error(std::string(who) + ": " + dwbc_last_error()); }
    void ck(int rc) {
        if (rc < 0) throw std::runtime_error(dwbc_last_error());
    }
    void *dalloc(size_t bytes) {
        void *p = dwbc_device_alloc(dev_, bytes);
        if (!p) fail("device allocation");
        allocs_.push_back(p);
        return p;
    }
};

PYBIND11_MODULE(dwbc, m) {
    m.doc() = "libdwbc_b200: the reference's RlWBCBridge over the batched CUDA whole-body-control path";
    py::class_<RlWBCBridge>(m, "RlWBCBridge")
        .def(py::init<int>())
        .def("reflectAction", &RlWBCBridge::reflectAction)
        .def("UpdateKinematics", &RlWBCBridge::UpdateKinematics)
        .def("SetContact", &RlWBCBridge::SetContact)
        .def("SetTaskSpace", &RlWBCBridge::SetTaskSpace)
        .def("CalcTorque", &RlWBCBridge::CalcTorque)
        .def("getTorqueCommand", &RlWBCBridge::getTorqueCommand)
        .def("Reset", &RlWBCBridge::Reset);
    py::class_<RlWBCBridgeBatch>(m, "RlWBCBridgeBatch")
        .def(py::init<int, int>(), py::arg("num_envs"), py::arg("device") = 0)
        .def("reflectAction", &RlWBCBridgeBatch::reflectAction)
        .def("UpdateKinematics", &RlWBCBridgeBatch::UpdateKinematics)
        .def("SetContact", &RlWBCBridgeBatch::SetContact)
        .def("SetTaskSpace", &RlWBCBridgeBatch::SetTaskSpace)
        .def("CalcTorque", &RlWBCBridgeBatch::CalcTorque)
        .def("getTorqueCommand", &RlWBCBridgeBatch::getTorqueCommand)
        .def("Reset", &RlWBCBridgeBatch::Reset);
}
