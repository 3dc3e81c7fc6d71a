// Robot model: host-side URDF reader and the flat, device-resident description the kernels use.
//
// Replaces, for the hot path, what the reference obtains from RBDL's URDF reader plus
// RobotData::InitModelData (reference src/dwbc.cpp:102-236): a floating-base tree whose joints are
// numbered depth first with the children of every link sorted by joint name, fixed-joint children merged
// into their parent, q = [p | quat xyz | theta | quat w], qdot = [v_world | omega_body | theta_dot]
// (reference src/pybind/rl_wbc_bridge.cpp:88-98).  Link id i = i-th moving body (0 = base); the pseudo
// link "COM" is id nb (reference src/dwbc.cpp:230-231).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace dwbc {

constexpr int MAXB = 48;      // bodies (base + revolute joints); TOCABI has 34, the reference's DoF study goes to 45+1
constexpr int MAX_DEPTH = 32;

// Plain-old-data, copied to the device once per model.  All 3x3 matrices are row-major.
struct DevModel {
    int nb;  // bodies, body 0 = floating base
    int n;   // system dof = nb + 5
    int m;   // actuated dof = nb - 1
    int max_depth;
    double total_mass;
    int parent[MAXB];
    int depth[MAXB];
    int level_start[MAX_DEPTH + 2];  // bodies sorted by depth: order[level_start[d] .. level_start[d+1])
    int order[MAXB];
    int child_start[MAXB + 1];  // children of body b: child_list[child_start[b] .. child_start[b+1])
    int child_list[MAXB];
    uint64_t anc_mask[MAXB];  // bit j set <=> body j is b itself or an ancestor of b (j >= 1)
    double E_T[MAXB][9];      // rotation joint-frame -> parent-frame at q = 0
    double r_T[MAXB][3];      // joint origin in the parent frame
    double axis[MAXB][3];     // unit joint axis in the joint (= body) frame
    double mass[MAXB];
    double com[MAXB][3];      // body frame
    double inertia[MAXB][9];  // about the COM, body frame
};

struct HostModel {
    DevModel dev;
    std::vector<std::string> names;  // body (URDF link) names
    // returns link id, nb for "COM", -1 if unknown (case-insensitive, reference src/dwbc.cpp:393-404)
    int link_id(const char *name) const;
};

// Throws std::runtime_error with a message on malformed input / unsupported joint types.
HostModel load_urdf_model(const std::string &path, bool floating);

}  // namespace dwbc
