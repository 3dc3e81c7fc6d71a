// dwbc_b200.hpp -- header-only C++ host shim over the C ABI (dwbc_b200.h): DWBC::RobotDataBatch keeps the method names, argument
// meaning, call order and 1/0 return convention of the reference's DWBC::RobotData (reference include/dwbc.h:59-430) for B
// states at once.  Where RobotData takes / exposes Eigen objects of one state, this class takes / returns DEVICE pointers to
// state-major arrays (layout in dwbc_b200.h); Eigen is not required.  A controller written against RobotData,
//     rd.UpdateKinematics(q, qdot, qddot); rd.SetContact(true, true); rd.CalcContactConstraint();
//     rd.SetTaskSpace(0, fstar); rd.CalcGravCompensation(); rd.CalcTaskControlTorque(true); rd.CalcContactRedistribute(true);
// (reference tests/dwbc_test.cpp:61-130) reads the same here with batched arguments.  Errors of the C ABI become
// std::runtime_error; per-state outcomes stay in the status array (DWBC_ST_* bits), like the reference's int returns.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "dwbc_b200.h"

namespace DWBC {

class RobotDataBatch {
public:
    int system_dof_ = 0, model_dof_ = 0, link_num_ = 0;
    double total_mass_ = 0.0;

    RobotDataBatch() = default;
    RobotDataBatch(const RobotDataBatch &) = delete;
    RobotDataBatch &operator=(const RobotDataBatch &) = delete;
    ~RobotDataBatch() {
        if (batch_) dwbc_batch_destroy(batch_);
        if (model_) dwbc_model_destroy(model_);
    }

    // LoadModelData(urdf_path, floating, verbose)   reference include/dwbc.h:237, + the batch size and device of this handle
    void LoadModelData(const std::string &urdf_path, bool floating, int batch, int device = 0, void *stream = nullptr) {
        model_ = dwbc_model_load_urdf(urdf_path.c_str(), floating ? 1 : 0);
        if (!model_) fail("LoadModelData");
        dwbc_model_info(model_, &system_dof_, &model_dof_, &link_num_, &total_mass_);
        batch_ = dwbc_batch_create(model_, batch, device);
        if (!batch_) fail("LoadModelData");
        stream_ = stream;
        B_ = batch;
    }
    int batch() const { return B_; }
    void SetStream(void *stream) { stream_ = stream; }
    int getLinkID(const std::string &name) const { return dwbc_model_link_id(model_, name.c_str()); }

    // UpdateKinematics(q, qdot, qddot)   :251   q_dev [B x (n+1)], qdot_dev [B x n]; qddot is unused by the reference's path
    void UpdateKinematics(const double *q_dev, const double *qdot_dev, const double * /*qddot_dev*/ = nullptr) {
        ck(dwbc_batch_update_kinematics(batch_, q_dev, qdot_dev, stream_), "UpdateKinematics");
    }
    // AddContactConstraint(link, contact_type, point, normal, lx, ly)   :259,265
    void AddContactConstraint(int link, int contact_type, const double point[3], const double normal[3], double lx = 0, double ly = 0) {
        ck(dwbc_batch_add_contact(batch_, link, contact_type, point, normal, lx, ly), "AddContactConstraint");
        ncontact_++;
    }
    void AddContactConstraint(const char *link_name, int contact_type, const double point[3], const double normal[3], double lx = 0, double ly = 0) {
        AddContactConstraint(link_checked(link_name), contact_type, point, normal, lx, ly);
    }
    // SetContact(flags...)   :291   flags_dev [B x number of registered contacts] (uint8)
    void SetContact(const uint8_t *flags_dev) { ck(dwbc_batch_set_contact(batch_, flags_dev, stream_), "SetContact"); }
    // CalcContactConstraint()   :280   (per-state 1/0 in the status array, bit DWBC_ST_CONTACT_OK)
    int CalcContactConstraint() { return ck(dwbc_batch_calc_contact_constraint(batch_, stream_), "CalcContactConstraint") == 0; }
    // AddTaskSpace(heirarchy, task_mode, link, point)   :319-320   and (heirarchy, TASK_CUSTOM, task_dof)   :318
    void AddTaskSpace(int heirarchy, int task_mode, int link, const double point[3]) {
        ck(dwbc_batch_add_task(batch_, heirarchy, task_mode, link, point), "AddTaskSpace");
    }
    void AddTaskSpace(int heirarchy, int task_mode, const char *link_name, const double point[3]) {
        AddTaskSpace(heirarchy, task_mode, link_checked(link_name), point);
    }
    void AddTaskSpaceCustom(int heirarchy, int task_dof) { ck(dwbc_batch_add_custom_task(batch_, heirarchy, task_dof), "AddTaskSpace"); }
    void ClearTaskSpace() { ck(dwbc_batch_clear_tasks(batch_), "ClearTaskSpace"); }
    // SetTaskSpace(heirarchy, f_star, J_task)   :333   f_star_dev [B x t]; J_task_dev column-major t x n (shared) or [B x t x n]
    void SetTaskSpace(int heirarchy, const double *f_star_dev, const double *J_task_dev = nullptr, bool J_per_state = false) {
        ck(dwbc_batch_set_task(batch_, heirarchy, f_star_dev, J_task_dev, J_per_state ? 1 : 0, stream_), "SetTaskSpace");
    }
    // SetTorqueLimit(torque_limit)   :249   (host vector, model_dof_ entries)
    void SetTorqueLimit(const std::vector<double> &tlim) {
        if ((int)tlim.size() != model_dof_) throw std::runtime_error("SetTorqueLimit: size is not matched");
        ck(dwbc_batch_set_torque_limit(batch_, tlim.data()), "SetTorqueLimit");
    }
    // CalcGravCompensation()   :246   -> torque_grav_dev [B x m]
    void CalcGravCompensation(double *torque_grav_dev) { ck(dwbc_batch_solve(batch_, DWBC_SOLVE_GRAV, torque_grav_dev, nullptr, nullptr, nullptr, stream_), "CalcGravCompensation"); }
    // CalcTaskControlTorque(hqp, init, update_task_space)   :349   (a batch is always a cold start: init = true)
    int CalcTaskControlTorque(bool hqp, double *torque_task_dev, int32_t *status_dev = nullptr) {
        const int stages = DWBC_SOLVE_GRAV | DWBC_SOLVE_TASK | (hqp ? DWBC_SOLVE_HQP : 0);
        return ck(dwbc_batch_solve(batch_, stages, nullptr, torque_task_dev, nullptr, status_dev, stream_), "CalcTaskControlTorque") == 0;
    }
    // CalcContactRedistribute(hqp, init)   :298   hqp = false: analytic two-foot branch
    int CalcContactRedistribute(bool hqp, double *torque_contact_dev, int32_t *status_dev = nullptr) {
        const int stages = hqp ? DWBC_SOLVE_REDIST : DWBC_SOLVE_REDIST_ANALYTIC;
        return ck(dwbc_batch_solve(batch_, stages, nullptr, nullptr, torque_contact_dev, status_dev, stream_), "CalcContactRedistribute") == 0;
    }
    // the whole cycle in one call (stages = OR of DWBC_SOLVE_*)
    void Solve(int stages, double *torque_grav_dev, double *torque_task_dev, double *torque_contact_dev, int32_t *status_dev) {
        ck(dwbc_batch_solve(batch_, stages, torque_grav_dev, torque_task_dev, torque_contact_dev, status_dev, stream_), "Solve");
    }
    // getContactForce(command_torque)   :303,  getZMP   src/dwbc.cpp:898-939
    void getContactForce(const double *command_torque_dev, double *contact_force_dev) {
        ck(dwbc_batch_get_contact_force(batch_, command_torque_dev, contact_force_dev, stream_), "getContactForce");
    }
    void getZMP(const double *contact_force_dev, double *zmp_dev) { ck(dwbc_batch_get_zmp(batch_, contact_force_dev, zmp_dev, stream_), "getZMP"); }
    // public RobotData fields (A_, A_inv_, J_C, Lambda_contact, ..., ts_[i].J_kt_): which = DWBC_MAT_*; returns rows, cols
    void getMatrix(int which, int level, double *out_dev, int *rows = nullptr, int *cols = nullptr) {
        ck(dwbc_batch_get_matrix(batch_, which, level, out_dev, rows, cols, stream_), "getMatrix");
    }
    // Reduced* (Ahn 2024)   :411-416
    void ReducedDynamicsCalculate(const std::vector<uint8_t> &pattern) { ck(dwbc_batch_reduced_dynamics(batch_, pattern.data(), stream_), "ReducedDynamicsCalculate"); }
    int ReducedCalcContactConstraint() { return ck(dwbc_batch_reduced_calc_contact_constraint(batch_, stream_), "ReducedCalcContactConstraint") == 0; }
    void ReducedSolve(int stages, double *torque_grav_dev, double *torque_task_dev, double *torque_contact_dev, int32_t *status_dev) {
        ck(dwbc_batch_reduced_solve(batch_, stages, torque_grav_dev, torque_task_dev, torque_contact_dev, status_dev, stream_), "ReducedSolve");
    }
    // ConfigureLQP + CalcControlTorqueLQP (Ahn 2025)   :365-371
    void CalcControlTorqueLQP(double *y_dev, double *torque_dev, int32_t *status_dev) {
        ck(dwbc_batch_calc_control_torque_lqp(batch_, y_dev, torque_dev, status_dev, stream_), "CalcControlTorqueLQP");
    }
    // whole cycle with HOST buffers (H2D + stages + D2H), what a host-side controller loop calls
    void CycleHost(int stages, const double *q, const double *qdot, const double *fstar, const uint8_t *flags, double *torque_grav, double *torque_task,
                   double *torque_contact, int32_t *status) {
        ck(dwbc_batch_cycle_host(batch_, stages, q, qdot, fstar, flags, torque_grav, torque_task, torque_contact, status, stream_), "CycleHost");
    }
    void SetHostSlices(int slices) { ck(dwbc_batch_set_host_slices(batch_, slices), "SetHostSlices"); }  // 0 = automatic, 1 = unsliced
    dwbc_batch_t *handle() { return batch_; }

private:
    dwbc_model_t *model_ = nullptr;
    dwbc_batch_t *batch_ = nullptr;
    void *stream_ = nullptr;
    int B_ = 0, ncontact_ = 0;
    [[noreturn]] void fail(const char *who) const { throw std::runtime_error(std::string(who) + ": " + dwbc_last_error()); }
    int ck(int rc, const char *who) const {
        if (rc < 0) fail(who);
        return rc;
    }
    int link_checked(const char *name) const {
        const int id = dwbc_model_link_id(model_, name);
        if (id < 0) throw std::runtime_error(std::string("Link Name is Wrong : ") + name);  // reference src/dwbc.cpp:407-411
        return id;
    }
};

}  // namespace DWBC
