// Test program for include/dwbc_compat/dwbc.h: the reference's verification sequence (reference tests/dwbc_test.cpp:29-131 CASE 1 and
// :133-235 CASE 2) written against DWBC::RobotData exactly as the reference writes it -- same calls, same public members -- with the
// golden binaries (tests/golden/cases/{1,2}, byte-identical copies of reference tests/cases) checked the way the reference's
// check_binary does (tests/dwbc_test_util.h:29-62), only with a tighter bound.  Built by tests/test_abi.py against the emulation
// build (CPU) and by tests/test_gpu_parity.py against libdwbc_b200.so (GPU).  Usage: robotdata_check <urdf> <golden dir>
#include <cstdio>
#include <fstream>
#include <string>

#include "dwbc.h"

using namespace DWBC;

static std::string g_dir;
static int g_fail = 0;

template <class Matrix>
static double check_binary(const char *filename, const Matrix &matrix) {
    std::ifstream in(g_dir + filename, std::ios::in | std::ios::binary);
    long rows = 0, cols = 0;
    in.read((char *)(&rows), sizeof(long));
    in.read((char *)(&cols), sizeof(long));
    MatrixXd mat_get;
    mat_get.resize(rows, cols);
    in.read((char *)mat_get.data(), rows * cols * sizeof(double));
    if (!in || rows != (long)matrix.rows() || cols != (long)matrix.cols()) {
        std::printf("%-24s SHAPE MISMATCH golden %ld x %ld, got %ld x %ld\n", filename, rows, cols, (long)matrix.rows(), (long)matrix.cols());
        g_fail++;
        return 1e300;
    }
    MatrixXd mine(rows, cols);
    for (long j = 0; j < cols; j++)
        for (long i = 0; i < rows; i++) mine(i, j) = matrix.data()[i + j * rows];
    const double rel = (mat_get - mine).norm() / (mat_get.norm() > 0 ? mat_get.norm() : 1.0);
    std::printf("%-24s rel err %.3e\n", filename, rel);
    return rel;
}
#define REQUIRE_REL(name, m, tol)                           \
    do {                                                    \
        if (!(check_binary(name, m) < (tol))) g_fail++;     \
    } while (0)

static void run_case(const std::string &urdf_path, int cs) {
    RobotData rd_;
    rd_.LoadModelData(urdf_path, true, false);
    VectorXd q = VectorXd::Zero(rd_.system_dof_ + 1);
    VectorXd qdot = VectorXd::Zero(rd_.system_dof_);
    VectorXd qddot = VectorXd::Zero(rd_.system_dof_);
    VectorXd fstar_1;
    fstar_1.setZero(6);
    VectorXd fstar_2;
    fstar_2.setZero(3);
    if (cs == 1) {
        q << 0, 0, 0.92983, 0, 0, 0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0, 0, 0, 0.3, 0.3, 1.5, -1.27, -1, 0, -1, 0, 0, 0,
            -0.3, -0.3, -1.5, 1.27, 1, 0, 1, 0, 1;
        fstar_1 << 0.1, 4.0, 0.1, 0.1, -0.1, 0.1;
        fstar_2 << 0.1, -0.1, 0.1;
    } else {
        q << 0, 0, 0.92983, 0, 0, 0, 0.1, 0.0, -0.24, 0.5, -0.6, 0.0, 0.05, 0.0, -0.21, 0.7, -0.31, 0.0, 0, 0, 0, 0.2, 0.5, 1.5, -1.27, -1.2, 0, -1, 0, 0,
            0, -0.3, -0.3, -1.5, 1.27, 1.3, 0.1, 1.3, 0, 1;
        fstar_1 << 0.4, 2.0, 0.1, 0.3, -0.1, 0.1;
        fstar_2 << 0.1, 0.1, 0.1;
    }
    std::string case_set = cs == 1 ? "/1" : "/2";
    rd_.UpdateKinematics(q, qdot, qddot);
    int left_foot_id = 6;
    int right_foot_id = 12;
    rd_.AddContactConstraint(left_foot_id, CONTACT_TYPE::CONTACT_6D, Vector3d(0.03, 0, -0.1585), Vector3d(0, 0, 1), 0.15, 0.075);
    rd_.AddContactConstraint(right_foot_id, CONTACT_TYPE::CONTACT_6D, Vector3d(0.03, 0, -0.1585), Vector3d(0, 0, 1), 0.15, 0.075);
    rd_.AddContactConstraint(23, CONTACT_TYPE::CONTACT_6D, Vector3d(0.03, 0, -0.1585), Vector3d(0, 0, 1), 0.04, 0.04);
    rd_.AddContactConstraint(31, CONTACT_TYPE::CONTACT_6D, Vector3d(0.03, 0, -0.1585), Vector3d(0, 0, 1), 0.04, 0.04);
    rd_.AddTaskSpace(0, TASK_LINK_6D, 0, Vector3d::Zero());
    rd_.AddTaskSpace(1, TASK_LINK_ROTATION, 15, Vector3d::Zero());
    VectorXd tlim;
    tlim.setConstant(rd_.model_dof_, 300);
    rd_.SetTorqueLimit(tlim);
    rd_.UpdateKinematics(q, qdot, qddot);
    rd_.SetContact(true, true);
    rd_.CalcContactConstraint();
    REQUIRE_REL((case_set + "/J_C").c_str(), rd_.J_C, 1e-9);
    REQUIRE_REL((case_set + "/A_inv_").c_str(), rd_.A_inv_, 1e-9);
    REQUIRE_REL((case_set + "/Lambda_contact").c_str(), rd_.Lambda_contact, 1e-9);
    REQUIRE_REL((case_set + "/J_C_INV_T").c_str(), rd_.J_C_INV_T, 1e-9);
    REQUIRE_REL((case_set + "/N_C").c_str(), rd_.N_C, 1e-9);
    REQUIRE_REL((case_set + "/W").c_str(), rd_.W, 1e-9);
    REQUIRE_REL((case_set + "/NwJw").c_str(), rd_.NwJw, 1e-9);
    REQUIRE_REL((case_set + "/W_inv").c_str(), rd_.W_inv, 1e-9);
    rd_.SetTaskSpace(0, fstar_1);
    rd_.SetTaskSpace(1, fstar_2);
    REQUIRE_REL((case_set + "/torque_grav_").c_str(), rd_.CalcGravCompensation(), 1e-9);
    if (!rd_.CalcTaskControlTorque(true)) g_fail++, std::printf("CalcTaskControlTorque returned 0\n");
    REQUIRE_REL((case_set + "/torque_task_").c_str(), rd_.torque_task_, 1e-9);
    if (!rd_.CalcContactRedistribute(true)) g_fail++, std::printf("CalcContactRedistribute returned 0\n");
    // case 2: qpOASES' point on a 4-dimensional optimal face is determined to ~3e-5 only
    REQUIRE_REL((case_set + "/torque_contact_").c_str(), rd_.torque_contact_, cs == 1 ? 1e-9 : 1e-4);
    std::printf("case %d: f_star_qp_[0] size %ld, contact_dof_ %d, link_[0].xpos z %.5f\n", cs, (long)rd_.ts_[0].f_star_qp_.size(), rd_.contact_dof_, rd_.link_[0].xpos(2));
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    g_dir = argv[2];
    try {
        run_case(argv[1], 1);
        run_case(argv[1], 2);
    } catch (const std::exception &e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 3;
    }
    std::printf(g_fail ? "FAILED %d checks\n" : "ALL OK\n", g_fail);
    return g_fail ? 1 : 0;
}
