// Device-side views shared by the kernels and the C ABI: configuration (contacts, task hierarchy,
// torque limit) and the per-batch workspace in HBM.
//
// HBM layout: every per-state quantity is stored STATE-MAJOR -- state b owns one contiguous record per
// array -- and every matrix inside a record is column-major with its natural leading dimension, i.e. the
// reference's Eigen layout.  A warp therefore streams its state's operands with fully coalesced 8-byte
// (or 16-byte) loads, and dwbc_batch_get_matrix is a plain device-to-device copy.
#pragma once
#include <cstdint>

#include "model.h"

namespace dwbc {

constexpr int MAX_CONTACTS = 4;        // registered contacts (reference tests register 4, enable 2)
constexpr int MAX_ACTIVE_CONTACTS = 2; // simultaneously active 6D contacts
constexpr int CMAX = 6 * MAX_ACTIVE_CONTACTS;  // contact dof capacity
constexpr int RMAX = CMAX - 6;         // internal-force dimension capacity
constexpr int MAX_LEVELS = 4;
constexpr int MAX_TASK_LINKS = 8;      // over all levels
constexpr int MAX_TASK_DOF = 24;       // per level
constexpr int QP_ROWS_MAX = 2 * (MAXB - 1) + 10 * MAX_ACTIVE_CONTACTS;
#ifdef DWBC_TEAM
constexpr int TEAM = DWBC_TEAM;        // build-time override (the emulation build of the tests uses a smaller team)
#else
constexpr int TEAM = 128;              // threads cooperating on one state in the dense stages (one CTA per state)
#endif
constexpr int ACTIVE_WORDS = 4;        // bit mask words per QP (>= QP_ROWS_MAX/32)

struct ContactCfg {
    int link;
    double point[3];
    double lx, ly, mu, muz;
};
struct TaskLinkCfg {
    int mode, link, dof;
    double point[3];
};
struct TaskCfg {
    int custom;       // 1: Jacobian supplied by the caller
    int dof;
    int link_start, link_count;  // into BatchCfg::tlinks
};
struct BatchCfg {
    int ncontact;
    ContactCfg contacts[MAX_CONTACTS];
    int nlevels;
    TaskCfg tasks[MAX_LEVELS];
    TaskLinkCfg tlinks[MAX_TASK_LINKS];
    int tlim_set;
    double tlim[MAXB];
};

struct Work {
    int B, n, m, nb;
    // a1
    double *A, *Ainv, *G, *Bnl, *pose, *com, *linkvel;
    // a2
    double *Jc;        // [B][CMAX*n], ld CMAX
    int *cdof;         // [B] contact dof of the state (0, 6, 12)
    int *cidx;         // [B][MAX_ACTIVE_CONTACTS] registered index of each active contact (-1 = none)
    // a3 / a4
    double *Lam;       // [B][CMAX*CMAX]
    double *Jbar;      // [B][CMAX*n]     J_C_INV_T
    double *AinvNc;    // [B][n*n]
    double *W, *Winv;  // [B][m*m]
    double *V2;        // [B][RMAX*m], ld RMAX (rows = basis vectors)
    double *NwJw;      // [B][m*RMAX]
    double *PC;        // [B][CMAX]
    double *tau_grav;  // [B][m]
    int *status;       // [B] DWBC_ST_* bits
    // a5
    const double *fstar[MAX_LEVELS];   // caller-provided (device)
    const double *jcustom[MAX_LEVELS]; // caller-provided custom Jacobians
    int jcustom_per_state[MAX_LEVELS];
    double *Jt[MAX_LEVELS];      // [B][t*n]
    double *Lamt[MAX_LEVELS];    // [B][t*t]
    double *Jkt[MAX_LEVELS];     // [B][m*t]
    double *Null[MAX_LEVELS];    // [B][m*m]
    double *Ntq[MAX_LEVELS];     // [B][m*t]   Null_{l-1} J_kt Lambda_t
    // a6 / a7
    double *fstar_qp[MAX_LEVELS];   // [B][t]
    double *contact_qp[MAX_LEVELS]; // [B][RMAX]
    double *torque_h[MAX_LEVELS];   // [B][m]
    uint32_t *active[MAX_LEVELS + 1];  // [B][ACTIVE_WORDS], last = redistribution
    double *tau_task, *tau_contact, *cf_redis;  // [B][m], [B][m], [B][RMAX]
    int *qp_iters;  // [B] total active-set iterations (diagnostics)
    long long *hqp_timing;  // diagnostics (null unless dwbc_batch_set_lean_timing): [B][32], hqp_kernel accumulates cycles per phase in slots 21..29
};

}  // namespace dwbc
