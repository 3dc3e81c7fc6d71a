// Test program for include/dwbc_b200.hpp (built by tests/test_abi.py against the EMULATION build of the C ABI, where "device"
// pointers are host pointers): the reference's gravity-compensation cycle (reference tests/dwbc_test.cpp:48-120) for one state.
// Prints torque_grav_ (33 values) and the status word.
#include <cstdio>
#include <vector>

#include "dwbc_b200.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    DWBC::RobotDataBatch rd;
    rd.LoadModelData(argv[1], true, 1);
    const double q[40] = {0, 0, 0.92983, 0, 0, 0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0.0, 0.0, -0.24, 0.6, -0.36, 0.0, 0, 0, 0,
                          0.3, 0.3, 1.5, -1.27, -1, 0, -1, 0, 0, 0, -0.3, -0.3, -1.5, 1.27, 1, 0, 1, 0, 1.0};
    std::vector<double> qd(39, 0.0), tg(33), tt(33), tc(33);
    const double pt[3] = {0.03, 0, -0.1585}, nz[3] = {0, 0, 1}, zero[3] = {0, 0, 0};
    rd.AddContactConstraint("l_ankleroll_link", DWBC_CONTACT_6D, pt, nz, 0.15, 0.075);
    rd.AddContactConstraint("r_ankleroll_link", DWBC_CONTACT_6D, pt, nz, 0.15, 0.075);
    rd.AddTaskSpace(0, DWBC_TASK_LINK_6D, "pelvis_link", zero);
    rd.AddTaskSpace(1, DWBC_TASK_LINK_ROTATION, "upperbody_link", zero);
    rd.SetTorqueLimit(std::vector<double>(33, 300.0));
    rd.UpdateKinematics(q, qd.data());
    const uint8_t flags[2] = {1, 1};
    rd.SetContact(flags);
    rd.CalcContactConstraint();
    const double f0[6] = {0.1, 4.0, 0.1, 0.1, -0.1, 0.1}, f1[3] = {0.1, -0.1, 0.1};
    rd.SetTaskSpace(0, f0);
    rd.SetTaskSpace(1, f1);
    int32_t st = 0;
    rd.Solve(DWBC_SOLVE_GRAV | DWBC_SOLVE_HQP | DWBC_SOLVE_REDIST, tg.data(), tt.data(), tc.data(), &st);
    for (int i = 0; i < 33; i++) std::printf("%.17g %.17g %.17g\n", tg[i], tt[i], tc[i]);
    std::printf("status %d\n", st);
    try {
        rd.AddContactConstraint("no_such_link", DWBC_CONTACT_6D, pt, nz);
        return 3;
    } catch (const std::runtime_error &e) {
        std::printf("error ok: %s\n", e.what());
    }
    return 0;
}
