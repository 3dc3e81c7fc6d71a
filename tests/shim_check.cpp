// Test program for include/dwbc_b200.hpp: the reference's cycle (reference tests/dwbc_test.cpp:48-131) for one state through
// DWBC::RobotDataBatch.  Built by tests/test_abi.py against the EMULATION build of the C ABI (CPU) and by tests/test_gpu_parity.py
// against libdwbc_b200.so (GPU); device buffers come from the staging helpers of the C ABI, so the same source runs on both.
// Prints the three torque vectors (33 rows) and the status word.
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
    // device buffers: q 40 | qd 39 | f0 6 | f1 3 | tg 33 | tt 33 | tc 33 | status | flags
    double *dev = (double *)dwbc_device_alloc(0, sizeof(double) * 256);
    if (!dev) return 4;
    double *d_q = dev, *d_qd = dev + 40, *d_f0 = dev + 80, *d_f1 = dev + 88, *d_tg = dev + 96, *d_tt = dev + 130, *d_tc = dev + 164;
    int32_t *d_st = (int32_t *)(dev + 200);
    uint8_t *d_fl = (uint8_t *)(dev + 204);
    auto up = [&](void *d, const void *h, size_t n) { dwbc_memcpy_h2d(0, d, h, n, nullptr); };
    auto dn = [&](void *h, const void *d, size_t n) { dwbc_memcpy_d2h(0, h, d, n, nullptr); };
    const double pt[3] = {0.03, 0, -0.1585}, nz[3] = {0, 0, 1}, zero[3] = {0, 0, 0};
    rd.AddContactConstraint("l_ankleroll_link", DWBC_CONTACT_6D, pt, nz, 0.15, 0.075);
    rd.AddContactConstraint("r_ankleroll_link", DWBC_CONTACT_6D, pt, nz, 0.15, 0.075);
    rd.AddTaskSpace(0, DWBC_TASK_LINK_6D, "pelvis_link", zero);
    rd.AddTaskSpace(1, DWBC_TASK_LINK_ROTATION, "upperbody_link", zero);
    rd.SetTorqueLimit(std::vector<double>(33, 300.0));
    up(d_q, q, sizeof(q));
    up(d_qd, qd.data(), sizeof(double) * 39);
    rd.UpdateKinematics(d_q, d_qd);
    const uint8_t flags[2] = {1, 1};
    up(d_fl, flags, 2);
    rd.SetContact(d_fl);
    rd.CalcContactConstraint();
    const double f0[6] = {0.1, 4.0, 0.1, 0.1, -0.1, 0.1}, f1[3] = {0.1, -0.1, 0.1};
    up(d_f0, f0, sizeof(f0));
    up(d_f1, f1, sizeof(f1));
    rd.SetTaskSpace(0, d_f0);
    rd.SetTaskSpace(1, d_f1);
    int32_t st = 0;
    rd.Solve(DWBC_SOLVE_GRAV | DWBC_SOLVE_HQP | DWBC_SOLVE_REDIST, d_tg, d_tt, d_tc, d_st);
    dn(tg.data(), d_tg, sizeof(double) * 33);
    dn(tt.data(), d_tt, sizeof(double) * 33);
    dn(tc.data(), d_tc, sizeof(double) * 33);
    dn(&st, d_st, sizeof(st));
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
