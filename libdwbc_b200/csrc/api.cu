// C ABI of the batched whole-body-control hot path (include/dwbc_b200.h).  Host logic only: argument
// checking, workspace ownership, kernel launches.  There is deliberately no CPU code path for the math.
#include "../../include/dwbc_b200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <functional>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "batch.h"
#include "cuda_compat.h"
#include "kernels_contact.cuh"
#include "kernels_hqp.cuh"
#include "kernels_io.cuh"
#include "kernels_lean.cuh"
#include "kernels_lqp.cuh"
#include "kernels_post.cuh"
#include "kernels_reduced.cuh"
#include "kernels_task.cuh"
#include "model.h"

using namespace dwbc;

struct dwbc_model {
    HostModel hm;
};

struct dwbc_batch {
    HostModel hm;
    int B = 0, device = 0;
    BatchCfg cfg;
    DevModel *d_model = nullptr;
    BatchCfg *d_cfg = nullptr;
    bool cfg_dirty = true;
    size_t alloc_bytes = 0;  // device memory this handle has allocated so far (dwbc_batch_workspace_bytes)
    Work w;
    std::vector<void *> allocs;
    std::vector<void *> task_allocs;
    bool tasks_allocated = false;  // tasks_level > 0
    // staging for the *_host entry point
    double *d_q = nullptr, *d_qd = nullptr, *d_fstar = nullptr;
    uint8_t *d_flags = nullptr;
    size_t d_fstar_cap = 0;
    // Output groups (dwbc_batch_set_outputs).  -1 = automatic: the staged calls materialise every public RobotData field
    // (DWBC_OUT_ALL), dwbc_batch_cycle_host produces torques + status only (DWBC_OUT_TORQUES).
    int outputs = -1;
    bool in_cycle_host = false;
    bool full_allocated = false;  // the per-state matrices of DWBC_OUT_ALL (A_, A_inv_, J_C, Lambda_contact, A_inv_N_C, W, W_inv, V2, ...)
    // DWBC_OUT_TORQUES: the staged calls only record their arguments; dwbc_batch_solve runs the whole cycle (kernels_lean.cuh)
    const double *lean_q = nullptr, *lean_qd = nullptr;
    const uint8_t *lean_flags = nullptr;
    bool lean_kin_pending = false, lean_contact_pending = false;
    int *d_list = nullptr, *d_count = nullptr;
    int *d_list2 = nullptr;  // QP-active states of the second contact class (d_count[2..3] are its counters)
    int *d_plist = nullptr, *d_pcount = nullptr;  // contact-pattern classes of the one-kernel cycle: lists [2][B], counts [2 streams][2]
    size_t list_off = 0;  // dwbc_batch_cycle_host: the slice's part of d_list and its stream's counter
    int count_slot = 0;
    long long *lean_timing = nullptr;  // diagnostics: device buffer [B][32] of phase stamps (dwbc_batch_set_lean_timing)
    int last_path = 0;  // 1: the last solve ran lean_cycle_kernel, 0: the generic kernels
    int tasks_level = 0;  // 0: no per-level buffers, 1: those of DWBC_OUT_TORQUES, 2: all
    // dwbc_batch_cycle_host pipelines slices of the batch over these streams (copies of one slice under the kernels of another)
    int host_slices = 0;  // 0 = automatic
    bool pipe_ready = false;
    cudaStream_t pipe_stream[2] = {};
    cudaEvent_t pipe_fork = {}, pipe_join[2] = {};
    int sm_count = 148;
    // call-order tracking
    bool have_kin = false, have_contact = false, have_cc = false;
    // optional per-kernel timing (dwbc_batch_set_profiling): event pairs recorded on the launch stream
    bool profiling = false;
    struct Pending {
        int k;
        cudaEvent_t e0, e1;
    };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> free_events;
    double k_ms[DWBC_K_COUNT] = {0};
    long long k_launches[DWBC_K_COUNT] = {0};
    // a9 reduced-dimension variant: second workspace (n = rn, m = rm) + extras
    Work wr, wq;  // wq: scratch outputs of the base-wrench QP (CalcSingleTaskTorqueWithQP_R_NC)
    BatchCfg *d_cfgNC = nullptr;
    ReducedX rx;
    ReducedCfg rc;
    BatchCfg cfgR;
    BatchCfg *d_cfgR = nullptr;
    std::vector<void *> red_allocs, red_task_allocs;
    int red_rn = 0;
    std::string red_task_sig;
    bool have_red_dyn = false, have_red_cc = false;
    bool red_tlim_set = false;
    std::vector<double> red_tlim;
    // section 8(f): f* generation and the MuJoCo / float32 wire format
    struct TrajHost {
        double gains[18] = {0};
        const double *pos_time = nullptr, *pos_init = nullptr, *vel_init = nullptr, *pos_des = nullptr, *vel_des = nullptr;
        const double *rot_time = nullptr, *rot_init = nullptr, *rot_des = nullptr;
    } traj[MAX_TASK_LINKS];
    const double *last_qd = nullptr;  // qdot of the last update_kinematics (caller-owned or mj_qd)
    double *mj_q = nullptr, *mj_qd = nullptr;
    // a10 LQP variant
    LqpOut lq{}, lqr{};
    double *lqnc_y = nullptr, *lqnc_tau_nc = nullptr, *lqnc_tau = nullptr;  // ConfigureLQP_R_NC + CalcControlTorqueLQP_R_NC
    int *lqnc_status = nullptr, *lqnc_iters = nullptr;
    bool have_lqp_r_nc = false;
    bool lqp_allocated = false, have_lqp = false, lqr_allocated = false, have_lqp_r = false;
};

namespace {
thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
bool cuda_ok(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return true;
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define CK(call)                                               \
    do {                                                       \
        if (!cuda_ok((call), #call)) return DWBC_ERR_CUDA;     \
    } while (0)

template <class T>
bool dev_alloc(dwbc_batch *b, T **p, size_t count, std::vector<void *> *list = nullptr) {
    void *raw = nullptr;
    if (!cuda_ok(cudaMalloc(&raw, count * sizeof(T) + 16), "cudaMalloc")) return false;
    cudaMemset(raw, 0, count * sizeof(T) + 16);
    b->alloc_bytes += count * sizeof(T) + 16;
    (list ? *list : b->allocs).push_back(raw);
    *p = (T *)raw;
    return true;
}

cudaEvent_t prof_event(dwbc_batch *b) {
    cudaEvent_t e;
    if (!b->free_events.empty()) {
        e = b->free_events.back();
        b->free_events.pop_back();
        return e;
    }
    cudaEventCreate(&e);
    return e;
}
// Brackets one kernel launch with events when profiling is on (no host synchronisation here).
struct KernelScope {
    dwbc_batch *b;
    cudaStream_t st;
    int k;
    cudaEvent_t e0{}, e1{};
    KernelScope(dwbc_batch *b_, int k_, cudaStream_t st_) : b(b_), st(st_), k(k_) {
        if (b->profiling) {
            e0 = prof_event(b);
            e1 = prof_event(b);
            cudaEventRecord(e0, st);
        }
    }
    ~KernelScope() {
        g_launches++;
        b->k_launches[k]++;
        if (b->profiling) {
            cudaEventRecord(e1, st);
            b->pending.push_back({k, e0, e1});
        }
    }
};

constexpr int WARPS_PER_BLOCK = 4;
inline int grid_for(int B, int wpb) { return (B + wpb - 1) / wpb; }

int upload_cfg(dwbc_batch *b, cudaStream_t st) {
    if (!b->cfg_dirty) return DWBC_OK;
    CK(cudaMemcpyAsync(b->d_cfg, &b->cfg, sizeof(BatchCfg), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // cfg lives on the host stack of the handle: keep it simple and safe
    b->cfg_dirty = false;
    return DWBC_OK;
}

// The records of states [s0, s0 + cnt) as a batch of their own: every per-state array is state-major (batch.h), so a slice is the
// same Work with each pointer advanced by s0 records.
Work slice_work(const Work &f, const BatchCfg &cfg, size_t s0, int cnt) {
    Work v = f;
    const size_t n = f.n, m = f.m, nb = f.nb;
    v.B = cnt;
    auto adv = [&](auto *&p, size_t rec) {
        if (p) p += s0 * rec;
    };
    adv(v.A, n * n), adv(v.Ainv, n * n), adv(v.G, n), adv(v.Bnl, n), adv(v.pose, nb * POSE_STRIDE), adv(v.com, 16), adv(v.linkvel, nb * 6);
    adv(v.Jc, CMAX * n), adv(v.cdof, 1), adv(v.cidx, MAX_ACTIVE_CONTACTS);
    adv(v.Lam, CMAX * CMAX), adv(v.Jbar, CMAX * n), adv(v.AinvNc, n * n), adv(v.W, m * m), adv(v.Winv, m * m), adv(v.V2, RMAX * m);
    adv(v.NwJw, m * RMAX), adv(v.PC, CMAX), adv(v.tau_grav, m), adv(v.status, 1);
    for (int l = 0; l < MAX_LEVELS; l++) {
        const size_t t = l < cfg.nlevels ? cfg.tasks[l].dof : 0;
        adv(v.fstar[l], t);
        if (v.jcustom_per_state[l]) adv(v.jcustom[l], t * n);
        adv(v.Jt[l], t * n), adv(v.Lamt[l], t * t), adv(v.Jkt[l], m * t), adv(v.Null[l], m * m), adv(v.Ntq[l], m * t);
        adv(v.fstar_qp[l], t), adv(v.contact_qp[l], RMAX), adv(v.torque_h[l], m), adv(v.active[l], ACTIVE_WORDS);
    }
    adv(v.active[MAX_LEVELS], ACTIVE_WORDS);
    adv(v.tau_task, m), adv(v.tau_contact, m), adv(v.cf_redis, RMAX), adv(v.qp_iters, 1);
    return v;
}

// Per-level buffers.  full = false: what the torque path needs (Ntorque, f_star_qp_, contact_qp_, active sets);
// full = true: every public per-level field (J_task_, Lambda_task_, J_kt_, Null_task_, torque_h_) as well.
int alloc_tasks(dwbc_batch *b, bool full = true) {
    const int want = full ? 2 : 1;
    if (b->tasks_level >= want) return DWBC_OK;
    const size_t B = b->B, n = b->w.n, m = b->w.m;
    for (int l = 0; l < b->cfg.nlevels; l++) {
        const size_t t = b->cfg.tasks[l].dof;
        if (b->tasks_level < 1 &&
            (!dev_alloc(b, &b->w.Ntq[l], B * m * t, &b->task_allocs) || !dev_alloc(b, &b->w.fstar_qp[l], B * t, &b->task_allocs) ||
             !dev_alloc(b, &b->w.contact_qp[l], B * RMAX, &b->task_allocs) || !dev_alloc(b, &b->w.active[l], B * ACTIVE_WORDS, &b->task_allocs)))
            return DWBC_ERR_CUDA;
        if (full && (!dev_alloc(b, &b->w.Jt[l], B * t * n, &b->task_allocs) || !dev_alloc(b, &b->w.Lamt[l], B * t * t, &b->task_allocs) ||
                     !dev_alloc(b, &b->w.Jkt[l], B * m * t, &b->task_allocs) || !dev_alloc(b, &b->w.Null[l], B * m * m, &b->task_allocs) ||
                     !dev_alloc(b, &b->w.torque_h[l], B * m, &b->task_allocs)))
            return DWBC_ERR_CUDA;
    }
    b->tasks_level = want;
    b->tasks_allocated = true;
    return DWBC_OK;
}
// The per-state matrices of DWBC_OUT_ALL, allocated when the generic kernels first run (a handle that only ever produces torques
// keeps ~19 KB per state instead of ~110 KB).
int ensure_full(dwbc_batch *b) {
    if (b->full_allocated) return DWBC_OK;
    Work &w = b->w;
    const size_t B = w.B, n = w.n, mm = w.m, nb = w.nb;
    const bool ok = dev_alloc(b, &w.A, B * n * n) && dev_alloc(b, &w.Ainv, B * n * n) && dev_alloc(b, &w.G, B * n) && dev_alloc(b, &w.Bnl, B * n) &&
                    dev_alloc(b, &w.com, B * 16) && dev_alloc(b, &w.linkvel, B * nb * 6) && dev_alloc(b, &w.Jc, B * CMAX * n) &&
                    dev_alloc(b, &w.Lam, B * CMAX * CMAX) && dev_alloc(b, &w.AinvNc, B * n * n) && dev_alloc(b, &w.W, B * mm * mm) &&
                    dev_alloc(b, &w.Winv, B * mm * mm) && dev_alloc(b, &w.V2, B * RMAX * mm);
    if (!ok) return DWBC_ERR_CUDA;
    b->full_allocated = true;
    return DWBC_OK;
}
void free_tasks(dwbc_batch *b) {
    for (void *p : b->task_allocs) cudaFree(p);
    b->task_allocs.clear();
    b->tasks_allocated = false;
    b->tasks_level = 0;
    for (int l = 0; l < MAX_LEVELS; l++) {
        b->w.Jt[l] = b->w.Lamt[l] = b->w.Jkt[l] = b->w.Null[l] = b->w.Ntq[l] = nullptr;
        b->w.fstar_qp[l] = b->w.contact_qp[l] = b->w.torque_h[l] = nullptr;
        b->w.active[l] = nullptr;
        b->w.fstar[l] = b->w.jcustom[l] = nullptr;
        b->w.jcustom_per_state[l] = 0;
    }
}

// ---- DWBC_OUT_TORQUES: the one-kernel cycle (kernels_lean.cuh) ----
inline bool lean_staged(const dwbc_batch *b) { return b->outputs == DWBC_OUT_TORQUES || (b->outputs < 0 && b->in_cycle_host); }

#ifndef DWBC_LEAN_LDN
#define DWBC_LEAN_LDN 44
#endif

template <class S>
int launch_lean(dwbc_batch *b, int stages, cudaStream_t st) {
    using L = LeanSmem<S>;
    const size_t smem = sizeof(double) * L::total * LEAN_GANG;
    const int lean_grid = (b->B + LEAN_GANG - 1) / LEAN_GANG;
    CK(cudaFuncSetAttribute(lean_cycle_kernel<S, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(lean_cycle_kernel<S, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LeanArgs a{};
    a.q = b->lean_q;
    a.qd = b->lean_qd;
    a.flags = b->lean_flags;
    a.hqp = (stages & DWBC_SOLVE_HQP) ? 1 : 0;
    a.redist = (stages & DWBC_SOLVE_REDIST) ? 1 : 0;
    a.timing = b->lean_timing ? b->lean_timing + 32 * b->list_off : nullptr;
    a.list = b->d_list + b->list_off;
    a.count = b->d_count + b->count_slot;
    // the slice's own region of the class lists: [2][slice size] starting at 2 * (first state of the slice)
    int *plist0 = b->d_plist + 2 * b->list_off, *pcount = b->d_pcount + 2 * b->count_slot;
    // QP-active states: one list per contact class (hqp_kernel is launched per class, the two-contact class with its dimensions as constants)
    int *const list1 = a.list, *const count1 = a.count, *const list2 = b->d_list2 + b->list_off, *const count2 = b->d_count + 2 + b->count_slot;
    CK(cudaMemsetAsync(count1, 0, sizeof(int), st));
    CK(cudaMemsetAsync(count2, 0, sizeof(int), st));
    CK(cudaMemsetAsync(pcount, 0, 2 * sizeof(int), st));
    {
        KernelScope ks(b, DWBC_K_LEAN, st);
        DWBC_LAUNCH(lean_classify_kernel, (b->B + 255) / 256, 256, 0, st, b->d_cfg, a.flags, b->B, plist0, pcount);
    }
    {
        KernelScope ks(b, DWBC_K_LEAN, st);
        a.plist = plist0;
        a.pcount = pcount;
        DWBC_LAUNCH((lean_cycle_kernel<S, 12>), lean_grid, TEAM * LEAN_GANG, smem, st, b->d_model, b->d_cfg, b->cfg, b->w, a);
    }
    {
        KernelScope ks(b, DWBC_K_LEAN, st);
        a.plist = plist0 + b->B;
        a.pcount = pcount + 1;
        a.list = list2;
        a.count = count2;
        DWBC_LAUNCH((lean_cycle_kernel<S, 6>), lean_grid, TEAM * LEAN_GANG, smem, st, b->d_model, b->d_cfg, b->cfg, b->w, a);
    }
    CK(cudaGetLastError());
    if (stages & (DWBC_SOLVE_HQP | DWBC_SOLVE_REDIST)) {
        // the states whose QPs are active (compact list written by the kernel above): one CTA each, CTAs beyond the count exit
        int nvmax = RMAX > 0 ? RMAX : 1;
        if (stages & DWBC_SOLVE_HQP)
            for (int l = 0; l < b->cfg.nlevels; l++) nvmax = b->cfg.tasks[l].dof + RMAX > nvmax ? b->cfg.tasks[l].dof + RMAX : nvmax;
        HqpSmem H(b->w.n, b->w.m, nvmax);
        const size_t hs = sizeof(double) * H.total;
        CK(cudaFuncSetAttribute(hqp_kernel_t<0, 0, 0, -1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
        Work wq = b->w;
        wq.hqp_timing = a.timing;
        const bool spec = b->w.n == 39 && b->w.m == 33 && nvmax == 27 && b->cfg.tlim_set;  // the headline model and hierarchy
        {
            KernelScope ks(b, DWBC_K_HQP, st);
            if (spec) {  // dimensions, row counts and the contact window as compile-time constants
                CK(cudaFuncSetAttribute(hqp_kernel_t<39, 33, 27, 66, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
                DWBC_LAUNCH((hqp_kernel_t<39, 33, 27, 66, 12>), b->B, TEAM, hs, st, b->d_model, b->d_cfg, wq, a.hqp, a.redist, nvmax, (const int *)list1, (const int *)count1);
            } else {
                DWBC_LAUNCH((hqp_kernel_t<0, 0, 0, -1, -1>), b->B, TEAM, hs, st, b->d_model, b->d_cfg, wq, a.hqp, a.redist, nvmax, (const int *)list1, (const int *)count1);
            }
        }
        {
            KernelScope ks(b, DWBC_K_HQP, st);
            if (spec) {
                CK(cudaFuncSetAttribute(hqp_kernel_t<39, 33, 27, 66, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
                DWBC_LAUNCH((hqp_kernel_t<39, 33, 27, 66, 6>), b->B, TEAM, hs, st, b->d_model, b->d_cfg, wq, a.hqp, a.redist, nvmax, (const int *)list2, (const int *)count2);
            } else {
                DWBC_LAUNCH((hqp_kernel_t<0, 0, 0, -1, -1>), b->B, TEAM, hs, st, b->d_model, b->d_cfg, wq, a.hqp, a.redist, nvmax, (const int *)list2, (const int *)count2);
            }
        }
    }
    CK(cudaGetLastError());
    return DWBC_OK;
}

// Runs the whole cycle in lean_cycle_kernel when a specialisation exists for the handle's model and hierarchy and the stage set;
// *done = false (and nothing launched) otherwise.
int lean_which(const dwbc_batch *b, int stages) {
    if (stages & DWBC_SOLVE_HQP) stages |= DWBC_SOLVE_TASK;
    if (TEAM != 128) return -1;
    if (stages & DWBC_SOLVE_REDIST_ANALYTIC) return -1;
    if (b->w.nb != 34) return -1;  // TOCABI (33 joints); other models run the generic kernels
    const bool tasks = (stages & DWBC_SOLVE_TASK) != 0;
    const BatchCfg &c = b->cfg;
    if ((stages & DWBC_SOLVE_REDIST) && !tasks && c.nlevels > 0) return -1;  // a lone redistribution reads torque_task_ of an earlier call
    if (c.ncontact <= 0) return -1;
    int d[MAX_LEVELS] = {0, 0, 0, 0};
    const int nl = tasks ? c.nlevels : 0;
    for (int l = 0; l < nl; l++) d[l] = c.tasks[l].dof;
    auto is = [&](int n, int a0, int a1, int a2) { return nl == n && d[0] == a0 && d[1] == a1 && d[2] == a2; };
    int which = -1;
    if (is(0, 0, 0, 0)) which = 0;
    else if (is(3, 3, 6, 21)) which = 1;
    else if (is(2, 3, 3, 0)) which = 2;
    else if (is(2, 6, 3, 0)) which = 3;
    return which;
}
int try_lean_cycle(dwbc_batch *b, int stages, cudaStream_t st, bool *done) {
    *done = false;
    const int which = lean_which(b, stages);
    if (which < 0) return DWBC_OK;
    const bool tasks = (stages & DWBC_SOLVE_TASK) != 0;
    int rc = upload_cfg(b, st);
    if (rc) return rc;
    if (tasks || (stages & DWBC_SOLVE_REDIST)) {
        rc = alloc_tasks(b, false);
        if (rc) return rc;
    }
    constexpr int LDN = DWBC_LEAN_LDN;
    switch (which) {
    case 0: rc = launch_lean<LeanSpec<34, LevelPack<>, LDN>>(b, stages, st); break;
    case 1: rc = launch_lean<LeanSpec<34, LevelPack<3, 6, 21>, LDN>>(b, stages, st); break;
    case 2: rc = launch_lean<LeanSpec<34, LevelPack<3, 3>, LDN>>(b, stages, st); break;
    case 3: rc = launch_lean<LeanSpec<34, LevelPack<6, 3>, LDN>>(b, stages, st); break;
    }
    if (rc) return rc;
    *done = true;
    return DWBC_OK;
}
}  // namespace

extern "C" {

const char *dwbc_last_error(void) { return g_err.c_str(); }
long long dwbc_launch_count(void) { return g_launches.load(); }

dwbc_model_t *dwbc_model_load_urdf(const char *urdf_path, int floating) {
    if (!urdf_path) {
        g_err = "urdf_path is null";
        return nullptr;
    }
    try {
        dwbc_model *m = new dwbc_model();
        m->hm = load_urdf_model(urdf_path, floating != 0);
        return m;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}
void dwbc_model_destroy(dwbc_model_t *m) { delete m; }
int dwbc_model_info(const dwbc_model_t *m, int *system_dof, int *model_dof, int *n_links, double *total_mass) {
    if (!m) return fail(DWBC_ERR_ARG, "model is null");
    if (system_dof) *system_dof = m->hm.dev.n;
    if (model_dof) *model_dof = m->hm.dev.m;
    if (n_links) *n_links = m->hm.dev.nb;
    if (total_mass) *total_mass = m->hm.dev.total_mass;
    return DWBC_OK;
}
int dwbc_model_link_id(const dwbc_model_t *m, const char *name) { return m ? m->hm.link_id(name) : -1; }
const char *dwbc_model_link_name(const dwbc_model_t *m, int link) {
    if (!m || link < 0 || link > m->hm.dev.nb) return nullptr;
    return link == m->hm.dev.nb ? "COM" : m->hm.names[link].c_str();
}

dwbc_batch_t *dwbc_batch_create(const dwbc_model_t *m, int batch, int device) {
    if (!m || batch <= 0) {
        g_err = "dwbc_batch_create: bad arguments";
        return nullptr;
    }
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
    dwbc_batch *b = new (std::nothrow) dwbc_batch();
    if (!b) return nullptr;
    b->hm = m->hm;
    b->B = batch;
    b->device = device;
    std::memset(&b->cfg, 0, sizeof(b->cfg));
    std::memset(&b->w, 0, sizeof(b->w));
    std::memset(&b->wr, 0, sizeof(b->wr));
    std::memset(&b->rx, 0, sizeof(b->rx));
    std::memset(&b->rc, 0, sizeof(b->rc));
    cudaDeviceProp prop;
    if (cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) b->sm_count = prop.multiProcessorCount;
    Work &w = b->w;
    w.B = batch;
    w.n = b->hm.dev.n;
    w.m = b->hm.dev.m;
    w.nb = b->hm.dev.nb;
    const size_t B = batch, n = w.n, mm = w.m, nb = w.nb;
    (void)n;
    bool ok = dev_alloc(b, &b->d_model, 1) && dev_alloc(b, &b->d_cfg, 1) && dev_alloc(b, &w.pose, B * nb * POSE_STRIDE) &&
              dev_alloc(b, &w.cdof, B) && dev_alloc(b, &w.cidx, B * MAX_ACTIVE_CONTACTS) && dev_alloc(b, &w.Jbar, B * CMAX * n) &&
              dev_alloc(b, &w.NwJw, B * mm * RMAX) && dev_alloc(b, &w.PC, B * CMAX) && dev_alloc(b, &w.tau_grav, B * mm) &&
              dev_alloc(b, &w.status, B) && dev_alloc(b, &w.tau_task, B * mm) && dev_alloc(b, &w.tau_contact, B * mm) &&
              dev_alloc(b, &w.cf_redis, B * RMAX) && dev_alloc(b, &w.active[MAX_LEVELS], B * ACTIVE_WORDS) && dev_alloc(b, &w.qp_iters, B) &&
              dev_alloc(b, &b->d_q, B * (n + 1)) && dev_alloc(b, &b->d_qd, B * n) && dev_alloc(b, &b->d_flags, B * MAX_CONTACTS) &&
              dev_alloc(b, &b->d_list, B) && dev_alloc(b, &b->d_list2, B) && dev_alloc(b, &b->d_count, 4) && dev_alloc(b, &b->d_plist, 2 * B) && dev_alloc(b, &b->d_pcount, 8);
    if (ok) ok = cuda_ok(cudaMemcpy(b->d_model, &b->hm.dev, sizeof(DevModel), cudaMemcpyHostToDevice), "cudaMemcpy(model)");
    if (!ok) {
        dwbc_batch_destroy(b);
        return nullptr;
    }
    return b;
}
void dwbc_batch_destroy(dwbc_batch_t *b) {
    if (!b) return;
    cudaSetDevice(b->device);
    for (void *p : b->allocs) cudaFree(p);
    for (void *p : b->task_allocs) cudaFree(p);
    for (void *p : b->red_allocs) cudaFree(p);
    for (void *p : b->red_task_allocs) cudaFree(p);
    if (b->d_fstar) cudaFree(b->d_fstar);
    if (b->pipe_ready) {
        for (int i = 0; i < 2; i++) {
            cudaStreamDestroy(b->pipe_stream[i]);
            cudaEventDestroy(b->pipe_join[i]);
        }
        cudaEventDestroy(b->pipe_fork);
    }
    for (auto &p : b->pending) {
        cudaEventDestroy(p.e0);
        cudaEventDestroy(p.e1);
    }
    for (cudaEvent_t e : b->free_events) cudaEventDestroy(e);
    delete b;
}
int dwbc_batch_size(const dwbc_batch_t *b) { return b ? b->B : 0; }
int dwbc_batch_limits(const dwbc_batch_t *, int *cmax, int *rmax, int *max_levels, int *max_task_dof) {
    if (cmax) *cmax = CMAX;
    if (rmax) *rmax = RMAX;
    if (max_levels) *max_levels = MAX_LEVELS;
    if (max_task_dof) *max_task_dof = MAX_TASK_DOF;
    return DWBC_OK;
}

int dwbc_batch_add_contact(dwbc_batch_t *b, int link, int contact_type, const double point[3], const double normal[3], double lx, double ly) {
    if (!b || !point) return fail(DWBC_ERR_ARG, "add_contact: null argument");
    (void)normal;  // kept for signature parity; the reference stores it but the path never reads it
    if (contact_type != DWBC_CONTACT_6D) return fail(DWBC_ERR_ARG, "add_contact: only CONTACT_6D is supported on this path");
    if (link < 0 || link >= b->hm.dev.nb) return fail(DWBC_ERR_ARG, "add_contact: link out of range");
    for (int i = 0; i < b->cfg.ncontact; i++)
        if (b->cfg.contacts[i].link == link) return fail(DWBC_ERR_ARG, "Contact Constraint Already Exist for Link");
    if (b->cfg.ncontact >= MAX_CONTACTS) return fail(DWBC_ERR_LIMIT, "add_contact: too many contacts");
    ContactCfg &c = b->cfg.contacts[b->cfg.ncontact];
    c.link = link;
    for (int k = 0; k < 3; k++) c.point[k] = point[k];
    c.lx = lx;
    c.ly = ly;
    c.mu = 0.2;  // reference src/contact_constraint.cpp:42
    c.muz = 0.2;
    b->cfg_dirty = true;
    return b->cfg.ncontact++;
}
int dwbc_batch_set_friction(dwbc_batch_t *b, int contact, double mu, double muz) {
    if (!b || contact < 0 || contact >= b->cfg.ncontact) return fail(DWBC_ERR_ARG, "set_friction: bad contact index");
    b->cfg.contacts[contact].mu = mu;
    b->cfg.contacts[contact].muz = muz;
    b->cfg_dirty = true;
    return DWBC_OK;
}
int dwbc_batch_add_task(dwbc_batch_t *b, int level, int task_mode, int link, const double point[3]) {
    if (!b || !point) return fail(DWBC_ERR_ARG, "add_task: null argument");
    if (task_mode < DWBC_TASK_LINK_6D || task_mode > DWBC_TASK_LINK_ROTATION_CUSTOM_FRAME) return fail(DWBC_ERR_ARG, "add_task: unknown task mode");
    if (link < 0 || link > b->hm.dev.nb) return fail(DWBC_ERR_ARG, "Link Name is not Correct");
    BatchCfg &c = b->cfg;
    if (level < 0 || level > c.nlevels) return fail(DWBC_ERR_ARG, "add_task: hierarchy levels must be added in order");
    int ntl = 0;
    for (int l = 0; l < c.nlevels; l++) {
        for (int k = 0; k < c.tasks[l].link_count; k++)
            if (c.tlinks[c.tasks[l].link_start + k].link == link) return fail(DWBC_ERR_ARG, "Task Space Already Exist for Link");
        ntl += c.tasks[l].link_count;
    }
    if (ntl >= MAX_TASK_LINKS) return fail(DWBC_ERR_LIMIT, "add_task: too many task links");
    const int dof = task_mode <= DWBC_TASK_LINK_6D_CUSTOM_FRAME ? 6 : 3;
    if (level == c.nlevels) {
        if (c.nlevels >= MAX_LEVELS) return fail(DWBC_ERR_LIMIT, "add_task: too many hierarchy levels");
        TaskCfg &t = c.tasks[c.nlevels++];
        t.custom = 0;
        t.dof = 0;
        t.link_start = ntl;
        t.link_count = 0;
    } else if (level != c.nlevels - 1) {
        // links of a level are stored contiguously: only the newest level can still grow
        return fail(DWBC_ERR_ARG, "add_task: links can only be appended to the most recently added level");
    }
    TaskCfg &t = c.tasks[level];
    if (t.custom) return fail(DWBC_ERR_ARG, "add_task: level holds a custom task");
    if (t.dof + dof > MAX_TASK_DOF) return fail(DWBC_ERR_LIMIT, "add_task: task dof exceeds MAX_TASK_DOF");
    b->traj[t.link_start + t.link_count] = dwbc_batch::TrajHost();  // a (re)assigned slot starts without gains / trajectories
    TaskLinkCfg &tl = c.tlinks[t.link_start + t.link_count++];
    tl.mode = task_mode;
    tl.link = link;
    tl.dof = dof;
    for (int k = 0; k < 3; k++) tl.point[k] = point[k];
    t.dof += dof;
    b->cfg_dirty = true;
    free_tasks(b);
    return DWBC_OK;
}
int dwbc_batch_add_custom_task(dwbc_batch_t *b, int level, int task_dof) {
    if (!b) return fail(DWBC_ERR_ARG, "add_custom_task: null handle");
    BatchCfg &c = b->cfg;
    if (level != c.nlevels) return fail(DWBC_ERR_ARG, "add_custom_task: hierarchy levels must be added in order");
    if (c.nlevels >= MAX_LEVELS) return fail(DWBC_ERR_LIMIT, "add_custom_task: too many hierarchy levels");
    if (task_dof <= 0 || task_dof > MAX_TASK_DOF) return fail(DWBC_ERR_LIMIT, "add_custom_task: task dof out of range");
    int ntl = 0;
    for (int l = 0; l < c.nlevels; l++) ntl += c.tasks[l].link_count;
    TaskCfg &t = c.tasks[c.nlevels++];
    t.custom = 1;
    t.dof = task_dof;
    t.link_start = ntl;
    t.link_count = 0;
    b->cfg_dirty = true;
    free_tasks(b);
    return DWBC_OK;
}
int dwbc_batch_clear_tasks(dwbc_batch_t *b) {
    if (!b) return fail(DWBC_ERR_ARG, "clear_tasks: null handle");
    b->cfg.nlevels = 0;
    for (auto &t : b->traj) t = dwbc_batch::TrajHost();  // gains / trajectory pointers belong to the task links that are gone
    b->cfg_dirty = true;
    free_tasks(b);
    return DWBC_OK;
}
int dwbc_batch_set_torque_limit(dwbc_batch_t *b, const double *tlim_host) {
    if (!b || !tlim_host) return fail(DWBC_ERR_ARG, "set_torque_limit: null argument");
    b->cfg.tlim_set = 1;
    for (int i = 0; i < b->w.m; i++) b->cfg.tlim[i] = tlim_host[i];
    b->cfg_dirty = true;
    return DWBC_OK;
}

int dwbc_batch_update_kinematics(dwbc_batch_t *b, const double *q_dev, const double *qdot_dev, void *stream) {
    if (!b || !q_dev || !qdot_dev) return fail(DWBC_ERR_ARG, "update_kinematics: null argument");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (lean_staged(b)) {
        // DWBC_OUT_TORQUES: nothing is computed before dwbc_batch_solve knows the whole cycle; q / qdot must stay valid until then
        b->lean_q = q_dev;
        b->lean_qd = qdot_dev;
        b->lean_kin_pending = true;
        b->lean_contact_pending = false;
        b->have_kin = true;
        b->last_qd = qdot_dev;
        b->have_contact = b->have_cc = false;
        b->have_red_dyn = b->have_red_cc = false;
        return DWBC_OK;
    }
    {
        const int rc = ensure_full(b);
        if (rc) return rc;
    }
    b->lean_kin_pending = false;
    const int wpb = WARPS_PER_BLOCK, grid = grid_for(b->B, wpb);
    const int n = b->w.n;
    CK(cudaMemsetAsync(b->w.status, 0xff, sizeof(int) * b->B, st));  // "valid so far"; stages clear it on failure
    {
        size_t smem = sizeof(double) * wpb * rbd_smem_doubles(b->w.nb);
        CK(cudaFuncSetAttribute(rbd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  // > 48 KB for the 40+ joint models
        KernelScope ks(b, DWBC_K_RBD, st);
        DWBC_LAUNCH(rbd_kernel, grid, wpb * 32, smem, st, b->d_model, q_dev, qdot_dev, b->w);
        CK(cudaGetLastError());
    }
    {
        size_t smem = sizeof(double) * ainv_smem_doubles(n);
        CK(cudaFuncSetAttribute(ainv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelScope ks(b, DWBC_K_AINV, st);
        DWBC_LAUNCH(ainv_kernel, b->B, TEAM, smem, st, b->w);
    }
    CK(cudaGetLastError());
    b->have_kin = true;
    b->last_qd = qdot_dev;
    b->have_contact = b->have_cc = false;
    b->have_red_dyn = b->have_red_cc = false;
    return DWBC_OK;
}

int dwbc_batch_set_contact(dwbc_batch_t *b, const uint8_t *flags_dev, void *stream) {
    if (!b || !flags_dev) return fail(DWBC_ERR_ARG, "set_contact: null argument");
    if (!b->have_kin) return fail(DWBC_ERR_ORDER, "set_contact: call update_kinematics first");
    if (b->cfg.ncontact == 0) return fail(DWBC_ERR_ORDER, "set_contact: no contact registered");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (b->lean_kin_pending) {  // DWBC_OUT_TORQUES: recorded, evaluated by dwbc_batch_solve
        b->lean_flags = flags_dev;
        b->lean_contact_pending = true;
        b->have_contact = true;
        b->have_cc = false;
        return DWBC_OK;
    }
    int rc = upload_cfg(b, st);
    if (rc) return rc;
    const int wpb = WARPS_PER_BLOCK, grid = grid_for(b->B, wpb);
    {
        KernelScope ks(b, DWBC_K_CONTACT_JAC, st);
        DWBC_LAUNCH(contact_jacobian_kernel, grid, wpb * 32, 0, st, b->d_model, b->d_cfg, flags_dev, b->w);
    }
    CK(cudaGetLastError());
    b->have_contact = true;
    b->have_cc = false;
    b->have_red_dyn = b->have_red_cc = false;
    return DWBC_OK;
}

int dwbc_batch_calc_contact_constraint(dwbc_batch_t *b, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "calc_contact_constraint: null handle");
    if (!b->have_contact) return fail(DWBC_ERR_ORDER, "calc_contact_constraint: call set_contact first");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (b->lean_kin_pending) {
        b->have_cc = true;
        return DWBC_OK;
    }
    ContactSmem L(b->w.n, b->w.m);
    size_t smem = sizeof(double) * L.total;
    CK(cudaFuncSetAttribute(contact_dynamics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelScope ks(b, DWBC_K_CONTACT_DYN, st);
        DWBC_LAUNCH(contact_dynamics_kernel, b->B, TEAM, smem, st, b->d_model, b->w);
    }
    CK(cudaGetLastError());
    b->have_cc = true;
    return DWBC_OK;
}

int dwbc_batch_set_task(dwbc_batch_t *b, int level, const double *fstar_dev, const double *jtask_dev, int jtask_per_state, void *) {
    if (!b || !fstar_dev) return fail(DWBC_ERR_ARG, "set_task: null argument");
    if (level < 0 || level >= b->cfg.nlevels) return fail(DWBC_ERR_ARG, "ERROR : task space size overflow");
    if (b->cfg.tasks[level].custom && !jtask_dev && !b->w.jcustom[level]) return fail(DWBC_ERR_ARG, "set_task: a custom task needs its Jacobian");
    b->w.fstar[level] = fstar_dev;
    if (jtask_dev) {
        b->w.jcustom[level] = jtask_dev;
        b->w.jcustom_per_state[level] = jtask_per_state ? 1 : 0;
    }
    return DWBC_OK;
}

int dwbc_batch_solve(dwbc_batch_t *b, int stages, double *tau_grav_dev, double *tau_task_dev, double *tau_contact_dev, int32_t *status_dev,
                     void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "solve: null handle");
    if (!b->have_cc) return fail(DWBC_ERR_ORDER, "solve: call calc_contact_constraint first");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_cfg(b, st);
    if (rc) return rc;
    const size_t B = b->B, m = b->w.m;
    if (stages & DWBC_SOLVE_HQP) stages |= DWBC_SOLVE_TASK;
    if (stages & DWBC_SOLVE_TASK) {
        if (b->cfg.nlevels == 0) return fail(DWBC_ERR_ORDER, "solve: no task registered");
        for (int l = 0; l < b->cfg.nlevels; l++)
            if (!b->w.fstar[l]) return fail(DWBC_ERR_ORDER, "solve: set_task has not been called for every level");
    }
    bool lean_done = false;
    if (b->lean_kin_pending) {
        if (!b->lean_contact_pending) return fail(DWBC_ERR_ORDER, "solve: call set_contact first");
        rc = try_lean_cycle(b, stages, st, &lean_done);
        if (rc) return rc;
        if (!lean_done) {
            // no specialisation of the one-kernel cycle for this model / hierarchy / stage set: run the generic kernels on the recorded inputs
            b->lean_kin_pending = false;
            const double *q = b->lean_q, *qd = b->lean_qd;
            const uint8_t *fl = b->lean_flags;
            const int keep = b->outputs;
            const bool keep_host = b->in_cycle_host;
            b->outputs = DWBC_OUT_ALL;
            b->in_cycle_host = false;
            rc = dwbc_batch_update_kinematics(b, q, qd, stream);
            if (!rc) rc = dwbc_batch_set_contact(b, fl, stream);
            if (!rc) rc = dwbc_batch_calc_contact_constraint(b, stream);
            b->outputs = keep;
            b->in_cycle_host = keep_host;
            if (rc) return rc;
        }
    }
    b->last_path = lean_done ? 1 : 0;
    if (!lean_done && (stages & (DWBC_SOLVE_TASK | DWBC_SOLVE_REDIST | DWBC_SOLVE_REDIST_ANALYTIC))) {
        rc = alloc_tasks(b);
        if (rc) return rc;
    }
    if (lean_done) stages = 0;  // everything ran above; only the copies of the results are left
    if (stages & DWBC_SOLVE_TASK) {
        TaskSmem L(b->w.n, b->w.m, b->cfg.tasks, b->cfg.nlevels);
        size_t smem = sizeof(double) * L.total;
        if (smem > 227 * 1024) return fail(DWBC_ERR_LIMIT, "solve: the task hierarchy does not fit the shared memory of one SM");
        CK(cudaFuncSetAttribute(task_space_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelScope ks(b, DWBC_K_TASK, st);
        DWBC_LAUNCH(task_space_kernel, b->B, TEAM, smem, st, b->d_model, b->d_cfg, b->w, (stages & DWBC_SOLVE_HQP) ? 1 : 0, 0);
    }
    // Without DWBC_SOLVE_TASK torque_task_ / torque_contact_ keep the values of the previous call: the reference's split order
    // CalcTaskControlTorque(true) then CalcContactRedistribute(true) feeds torque_grav_ + torque_task_ + torque_contact_ into the
    // redistribution (reference src/dwbc.cpp:1372-1375).  The workspace is zero-initialised, task_space_kernel clears both.
    if (stages & (DWBC_SOLVE_HQP | DWBC_SOLVE_REDIST)) {
        int nvmax = RMAX > 0 ? RMAX : 1;
        if (stages & DWBC_SOLVE_HQP)
            for (int l = 0; l < b->cfg.nlevels; l++) nvmax = b->cfg.tasks[l].dof + RMAX > nvmax ? b->cfg.tasks[l].dof + RMAX : nvmax;
        HqpSmem L(b->w.n, b->w.m, nvmax);
        size_t smem = sizeof(double) * L.total;
        CK(cudaFuncSetAttribute(hqp_kernel_t<0, 0, 0, -1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelScope ks(b, DWBC_K_HQP, st);
        DWBC_LAUNCH((hqp_kernel_t<0, 0, 0, -1, -1>), b->B, TEAM, smem, st, b->d_model, b->d_cfg, b->w, (stages & DWBC_SOLVE_HQP) ? 1 : 0,
                    (stages & DWBC_SOLVE_REDIST) ? 1 : 0, nvmax, (const int *)nullptr, (const int *)nullptr);
    }
    if ((stages & DWBC_SOLVE_REDIST_ANALYTIC) && !(stages & DWBC_SOLVE_REDIST)) {
        const int wpb = WARPS_PER_BLOCK, grid = grid_for(b->B, wpb);
        const size_t smem = sizeof(double) * wpb * (MAXB + CMAX + 6 * 7 + 8);
        KernelScope ks(b, DWBC_K_POST, st);
        DWBC_LAUNCH(redistribute_analytic_kernel, grid, wpb * 32, smem, st, b->d_model, b->d_cfg, b->w);
    }
    CK(cudaGetLastError());
    if (tau_grav_dev) CK(cudaMemcpyAsync(tau_grav_dev, b->w.tau_grav, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (tau_task_dev) CK(cudaMemcpyAsync(tau_task_dev, b->w.tau_task, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (tau_contact_dev) CK(cudaMemcpyAsync(tau_contact_dev, b->w.tau_contact, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (status_dev) CK(cudaMemcpyAsync(status_dev, b->w.status, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    return DWBC_OK;
}

int dwbc_batch_get_matrix(dwbc_batch_t *b, int which, int level, double *out_dev, int *rows, int *cols, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "get_matrix: null handle");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Work &w = b->w;
    const int n = w.n, m = w.m, nb = w.nb;
    const double *src = nullptr;
    int r = 0, c = 0;
    int t = 0;
    if (which >= DWBC_MAT_J_TASK) {
        if (level < 0 || level >= b->cfg.nlevels) return fail(DWBC_ERR_ARG, "get_matrix: level out of range");
        if (!b->tasks_allocated) return fail(DWBC_ERR_ORDER, "get_matrix: task quantities have not been computed yet");
        t = b->cfg.tasks[level].dof;
    }
    switch (which) {
    case DWBC_MAT_A: src = w.A; r = n; c = n; break;
    case DWBC_MAT_A_INV: src = w.Ainv; r = n; c = n; break;
    case DWBC_MAT_G: src = w.G; r = n; c = 1; break;
    case DWBC_MAT_B: src = w.Bnl; r = n; c = 1; break;
    case DWBC_MAT_J_C: src = w.Jc; r = CMAX; c = n; break;
    case DWBC_MAT_LAMBDA_C: src = w.Lam; r = CMAX; c = CMAX; break;
    case DWBC_MAT_J_C_INV_T: src = w.Jbar; r = CMAX; c = n; break;
    case DWBC_MAT_A_INV_N_C: src = w.AinvNc; r = n; c = n; break;
    case DWBC_MAT_W: src = w.W; r = m; c = m; break;
    case DWBC_MAT_W_INV: src = w.Winv; r = m; c = m; break;
    case DWBC_MAT_V2: src = w.V2; r = RMAX; c = m; break;
    case DWBC_MAT_NWJW: src = w.NwJw; r = m; c = RMAX; break;
    case DWBC_MAT_P_C: src = w.PC; r = CMAX; c = 1; break;
    case DWBC_MAT_TORQUE_GRAV: src = w.tau_grav; r = m; c = 1; break;
    case DWBC_MAT_TORQUE_TASK: src = w.tau_task; r = m; c = 1; break;
    case DWBC_MAT_TORQUE_CONTACT: src = w.tau_contact; r = m; c = 1; break;
    case DWBC_MAT_CF_REDIS: src = w.cf_redis; r = RMAX; c = 1; break;
    case DWBC_MAT_J_TASK: src = w.Jt[level]; r = t; c = n; break;
    case DWBC_MAT_LAMBDA_TASK: src = w.Lamt[level]; r = t; c = t; break;
    case DWBC_MAT_J_KT: src = w.Jkt[level]; r = m; c = t; break;
    case DWBC_MAT_NULL_TASK: src = w.Null[level]; r = m; c = m; break;
    case DWBC_MAT_FSTAR_QP: src = w.fstar_qp[level]; r = t; c = 1; break;
    case DWBC_MAT_CONTACT_QP: src = w.contact_qp[level]; r = RMAX; c = 1; break;
    case DWBC_MAT_TORQUE_H: src = w.torque_h[level]; r = m; c = 1; break;
    case DWBC_MAT_N_C: r = n; c = n; break;
    case DWBC_MAT_COM: r = 3; c = 1; break;
    case DWBC_MAT_LINK_POS: r = 3; c = nb; break;
    case DWBC_MAT_LINK_ROT: r = 9; c = nb; break;
    case DWBC_MAT_LINK_VEL: src = w.linkvel; r = 6; c = nb; break;
    default: return fail(DWBC_ERR_ARG, "get_matrix: unknown matrix id");
    }
    if (rows) *rows = r;
    if (cols) *cols = c;
    if (!out_dev) return DWBC_OK;
    const size_t bytes = sizeof(double) * (size_t)b->B * r * c;
    const bool computed_on_request = which == DWBC_MAT_N_C || which == DWBC_MAT_COM || which == DWBC_MAT_LINK_POS || which == DWBC_MAT_LINK_ROT;
    const bool torque_group = which == DWBC_MAT_TORQUE_GRAV || which == DWBC_MAT_TORQUE_TASK || which == DWBC_MAT_TORQUE_CONTACT ||
                              which == DWBC_MAT_CF_REDIS || which == DWBC_MAT_FSTAR_QP || which == DWBC_MAT_CONTACT_QP;
    if ((b->last_path == 1 || b->lean_kin_pending) && !torque_group)
        return fail(DWBC_ERR_ORDER, "get_matrix: this field is not materialised with DWBC_OUT_TORQUES; call dwbc_batch_set_outputs(b, DWBC_OUT_ALL) before the cycle");
    if (!computed_on_request && !src) return fail(DWBC_ERR_ORDER, "get_matrix: the field has not been computed yet");
    if (computed_on_request && !b->full_allocated) return fail(DWBC_ERR_ORDER, "get_matrix: call update_kinematics first");
    if (src) {
        CK(cudaMemcpyAsync(out_dev, src, bytes, cudaMemcpyDeviceToDevice, st));
        return DWBC_OK;
    }
    const int wpb = WARPS_PER_BLOCK, grid = grid_for(b->B, wpb);
    if (which == DWBC_MAT_N_C) {
        if (!b->have_cc) return fail(DWBC_ERR_ORDER, "get_matrix(N_C): call calc_contact_constraint first");
        KernelScope ks(b, DWBC_K_EXPORT, st);
        DWBC_LAUNCH(nc_kernel, grid, wpb * 32, 0, st, b->w, out_dev);
    } else {
        KernelScope ks(b, DWBC_K_EXPORT, st);
        DWBC_LAUNCH(pose_export_kernel, grid, wpb * 32, 0, st, b->w, which == DWBC_MAT_COM ? 0 : (which == DWBC_MAT_LINK_POS ? 1 : 2), out_dev);
    }
    CK(cudaGetLastError());
    return DWBC_OK;
}

#define NEED_ALL(b, who)                                                                                                          \
    do {                                                                                                                          \
        if ((b)->lean_kin_pending || (b)->last_path == 1)                                                                         \
            return fail(DWBC_ERR_ORDER, who ": needs the fields of DWBC_OUT_ALL (call dwbc_batch_set_outputs(b, DWBC_OUT_ALL) before the cycle)"); \
    } while (0)

int dwbc_batch_get_contact_force(dwbc_batch_t *b, const double *torque_dev, double *cf_dev, void *stream) {
    if (!b || !cf_dev) return fail(DWBC_ERR_ARG, "get_contact_force: null argument");
    NEED_ALL(b, "get_contact_force");
    if (!b->have_cc) return fail(DWBC_ERR_ORDER, "get_contact_force: call calc_contact_constraint first");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_cfg(b, st);
    if (rc) return rc;
    const int wpb = WARPS_PER_BLOCK, grid = grid_for(b->B, wpb);
    {
        KernelScope ks(b, DWBC_K_POST, st);
        DWBC_LAUNCH(contact_force_kernel, grid, wpb * 32, sizeof(double) * wpb * (MAXB + CMAX), st, b->d_model, b->d_cfg, b->w, 1, torque_dev,
                    (const double *)nullptr, cf_dev, (double *)nullptr);
    }
    CK(cudaGetLastError());
    return DWBC_OK;
}
int dwbc_batch_get_zmp(dwbc_batch_t *b, const double *cf_dev, double *zmp_dev, void *stream) {
    if (!b || !cf_dev || !zmp_dev) return fail(DWBC_ERR_ARG, "get_zmp: null argument");
    NEED_ALL(b, "get_zmp");
    if (!b->have_cc) return fail(DWBC_ERR_ORDER, "get_zmp: call calc_contact_constraint first");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_cfg(b, st);
    if (rc) return rc;
    const int wpb = WARPS_PER_BLOCK, grid = grid_for(b->B, wpb);
    {
        KernelScope ks(b, DWBC_K_POST, st);
        DWBC_LAUNCH(contact_force_kernel, grid, wpb * 32, sizeof(double) * wpb * (MAXB + CMAX), st, b->d_model, b->d_cfg, b->w, 2,
                    (const double *)nullptr, cf_dev, (double *)nullptr, zmp_dev);
    }
    CK(cudaGetLastError());
    return DWBC_OK;
}

int dwbc_batch_get_active_set(dwbc_batch_t *b, int level, uint32_t *out_dev, void *stream) {
    if (!b || !out_dev) return fail(DWBC_ERR_ARG, "get_active_set: null argument");
    if (level >= b->cfg.nlevels || level < -1) return fail(DWBC_ERR_ARG, "get_active_set: level out of range");
    if (level >= 0 && !b->tasks_allocated) return fail(DWBC_ERR_ORDER, "get_active_set: no QP has been solved yet");
    CK(cudaSetDevice(b->device));
    const uint32_t *src = level < 0 ? b->w.active[MAX_LEVELS] : b->w.active[level];
    CK(cudaMemcpyAsync(out_dev, src, sizeof(uint32_t) * (size_t)b->B * ACTIVE_WORDS, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DWBC_OK;
}

int dwbc_batch_get_qp_iterations(dwbc_batch_t *b, int32_t *out_dev, void *stream) {
    if (!b || !out_dev) return fail(DWBC_ERR_ARG, "get_qp_iterations: null argument");
    CK(cudaSetDevice(b->device));
    CK(cudaMemcpyAsync(out_dev, b->w.qp_iters, sizeof(int32_t) * (size_t)b->B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DWBC_OK;
}

int dwbc_batch_set_outputs(dwbc_batch_t *b, int outputs) {
    if (!b) return fail(DWBC_ERR_ARG, "set_outputs: null handle");
    if (outputs != DWBC_OUT_AUTO && outputs != DWBC_OUT_TORQUES && outputs != DWBC_OUT_ALL) return fail(DWBC_ERR_ARG, "set_outputs: unknown output group");
    b->outputs = outputs;
    b->lean_kin_pending = b->lean_contact_pending = false;
    b->have_kin = b->have_contact = b->have_cc = false;  // the next cycle starts from update_kinematics
    return DWBC_OK;
}
int dwbc_batch_last_path(const dwbc_batch_t *b) { return b ? b->last_path : -1; }
long long dwbc_batch_workspace_bytes(const dwbc_batch_t *b) { return b ? (long long)b->alloc_bytes : -1; }
int dwbc_batch_set_lean_timing(dwbc_batch_t *b, long long *stamps_dev) {
    if (!b) return fail(DWBC_ERR_ARG, "set_lean_timing: null handle");
    b->lean_timing = stamps_dev;
    return DWBC_OK;
}

int dwbc_batch_set_profiling(dwbc_batch_t *b, int enable) {
    if (!b) return fail(DWBC_ERR_ARG, "set_profiling: null handle");
    b->profiling = enable != 0;
    return DWBC_OK;
}

int dwbc_batch_get_kernel_times(dwbc_batch_t *b, double *ms_sum, long long *launches, int reset) {
    if (!b) return fail(DWBC_ERR_ARG, "get_kernel_times: null handle");
    CK(cudaSetDevice(b->device));
    for (auto &p : b->pending) {
        CK(cudaEventSynchronize(p.e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, p.e0, p.e1));
        b->k_ms[p.k] += ms;
        b->free_events.push_back(p.e0);
        b->free_events.push_back(p.e1);
    }
    b->pending.clear();
    for (int k = 0; k < DWBC_K_COUNT; k++) {
        if (ms_sum) ms_sum[k] = b->k_ms[k];
        if (launches) launches[k] = b->k_launches[k];
        if (reset) b->k_ms[k] = 0.0, b->k_launches[k] = 0;
    }
    return DWBC_OK;
}

const char *dwbc_kernel_name(int k) {
    static const char *names[DWBC_K_COUNT] = {"rbd_kernel", "ainv_kernel", "contact_jacobian_kernel", "contact_dynamics_kernel",
                                              "task_space_kernel", "hqp_kernel", "export_kernels", "post_kernels", "reduced_kernels", "lqp_kernel", "lean_cycle_kernel"};
    return (k >= 0 && k < DWBC_K_COUNT) ? names[k] : nullptr;
}

int dwbc_batch_set_host_slices(dwbc_batch_t *b, int slices) {
    if (!b || slices < 0) return fail(DWBC_ERR_ARG, "set_host_slices: bad argument");
    b->host_slices = slices;
    return DWBC_OK;
}

int dwbc_batch_cycle_host(dwbc_batch_t *b, int stages, const double *q_host, const double *qdot_host, const double *fstar_host,
                          const uint8_t *flags_host, double *tau_grav_host, double *tau_task_host, double *tau_contact_host,
                          int32_t *status_host, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "cycle_host: null argument");
    const double *lv[MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr};
    if (fstar_host) {
        size_t off = 0;
        for (int l = 0; l < b->cfg.nlevels; l++) {
            lv[l] = fstar_host + off;
            off += (size_t)b->cfg.tasks[l].dof * b->B;
        }
    }
    return dwbc_batch_cycle_host_levels(b, stages, q_host, qdot_host, fstar_host ? lv : nullptr, flags_host, tau_grav_host, tau_task_host,
                                        tau_contact_host, status_host, stream);
}

int dwbc_batch_cycle_host_levels(dwbc_batch_t *b, int stages, const double *q_host, const double *qdot_host, const double *const *fstar_levels_host,
                                 const uint8_t *flags_host, double *tau_grav_host, double *tau_task_host, double *tau_contact_host,
                                 int32_t *status_host, void *stream) {
    if (!b || !q_host || !qdot_host || !flags_host) return fail(DWBC_ERR_ARG, "cycle_host: null argument");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = b->B, n = b->w.n, m = b->w.m, nc = b->cfg.ncontact;
    const bool need_tasks = (stages & (DWBC_SOLVE_TASK | DWBC_SOLVE_HQP)) != 0;
    size_t ftot = 0;
    for (int l = 0; l < b->cfg.nlevels; l++) ftot += (size_t)b->cfg.tasks[l].dof * B;
    if (need_tasks) {
        if (!fstar_levels_host) return fail(DWBC_ERR_ARG, "cycle_host: fstar_host is null");
        for (int l = 0; l < b->cfg.nlevels; l++)
            if (!fstar_levels_host[l]) return fail(DWBC_ERR_ARG, "cycle_host: f* of a hierarchy level is null");
        if (ftot > b->d_fstar_cap) {
            if (b->d_fstar) cudaFree(b->d_fstar);
            void *raw = nullptr;
            CK(cudaMalloc(&raw, ftot * sizeof(double)));
            b->d_fstar = (double *)raw;
            b->d_fstar_cap = ftot;
        }
    }
    // everything that allocates or uploads the configuration happens before the batch is sliced
    int rc = upload_cfg(b, st);
    if (rc) return rc;
    b->in_cycle_host = true;
    struct HostFlag {
        dwbc_batch *b;
        ~HostFlag() { b->in_cycle_host = false; }
    } host_flag{b};
    // The path is chosen once, before the batch is sliced (allocations must see the whole batch): the one-kernel cycle when the
    // output group allows it and a specialisation exists, the generic kernels otherwise.
    const bool lean = lean_staged(b) && lean_which(b, stages) >= 0;
    const int keep_outputs = b->outputs;
    struct OutputsGuard {
        dwbc_batch *b;
        int keep;
        ~OutputsGuard() { b->outputs = keep; b->list_off = 0; b->count_slot = 0; }
    } outputs_guard{b, keep_outputs};
    b->outputs = lean ? DWBC_OUT_TORQUES : DWBC_OUT_ALL;
    if (!lean) {
        rc = ensure_full(b);
        if (rc) return rc;
    }
    if (stages & (DWBC_SOLVE_TASK | DWBC_SOLVE_HQP | DWBC_SOLVE_REDIST | DWBC_SOLVE_REDIST_ANALYTIC)) {
        rc = alloc_tasks(b, !lean);
        if (rc) return rc;
    }
    // The H2D of slice k+1 and the D2H of slice k-1 run under the kernels of slice k.  Slice size: the generic kernels (DWBC_OUT_ALL) liked
    // slices of 4096 states, at most 16 (profiles/r01_host_slices.json); the torque path wants few, large slices -- hqp_kernel only sees the
    // QP-active ~44 % of a slice, and 4096 states leave it less than two waves of CTAs (profiles/r02_host_slices.json: 65 536 states in
    // 19.2 / 18.0 / 18.2 / 18.3 / 20.2 / 23.3 ms with 1 / 2 / 4 / 8 / 16 / 32 slices) -- so slices of 32 768 states, at most 4.
    int K = b->host_slices > 0 ? b->host_slices : (lean ? (int)std::min<size_t>(4, B / 32768) : (int)std::min<size_t>(16, B / 4096));
    K = std::max(1, std::min<int>(K, (int)B));
    if (K > 1 && !b->pipe_ready) {
        for (int i = 0; i < 2; i++) {
            CK(cudaStreamCreateWithFlags(&b->pipe_stream[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&b->pipe_join[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&b->pipe_fork, cudaEventDisableTiming));
        b->pipe_ready = true;
    }
    if (K > 1) {
        CK(cudaEventRecord(b->pipe_fork, st));
        for (int i = 0; i < 2; i++) CK(cudaStreamWaitEvent(b->pipe_stream[i], b->pipe_fork, 0));
    }
    const Work full = b->w;
    const size_t per = (B + K - 1) / K;
    rc = DWBC_OK;
    for (int k = 0; k < K && rc == DWBC_OK; k++) {
        const size_t s0 = (size_t)k * per, cnt = std::min(per, B - std::min(B, s0));
        if (cnt == 0) break;
        cudaStream_t sk = K > 1 ? b->pipe_stream[k & 1] : st;
        b->w = slice_work(full, b->cfg, s0, (int)cnt);
        b->B = (int)cnt;
        b->list_off = s0;
        b->count_slot = k & 1;
        rc = [&]() -> int {
            if (need_tasks) {
                size_t off = 0;
                for (int l = 0; l < b->cfg.nlevels; l++) {
                    const size_t t = b->cfg.tasks[l].dof;
                    CK(cudaMemcpyAsync(b->d_fstar + off + s0 * t, fstar_levels_host[l] + s0 * t, cnt * t * sizeof(double), cudaMemcpyHostToDevice, sk));
                    off += t * B;
                }
            }
            CK(cudaMemcpyAsync(b->d_q + s0 * (n + 1), q_host + s0 * (n + 1), sizeof(double) * cnt * (n + 1), cudaMemcpyHostToDevice, sk));
            CK(cudaMemcpyAsync(b->d_qd + s0 * n, qdot_host + s0 * n, sizeof(double) * cnt * n, cudaMemcpyHostToDevice, sk));
            CK(cudaMemcpyAsync(b->d_flags + s0 * nc, flags_host + s0 * nc, cnt * nc, cudaMemcpyHostToDevice, sk));
            int r = dwbc_batch_update_kinematics(b, b->d_q + s0 * (n + 1), b->d_qd + s0 * n, sk);
            if (r) return r;
            r = dwbc_batch_set_contact(b, b->d_flags + s0 * nc, sk);
            if (r) return r;
            r = dwbc_batch_calc_contact_constraint(b, sk);
            if (r) return r;
            if (need_tasks) {
                size_t off = 0;
                for (int l = 0; l < b->cfg.nlevels; l++) {
                    const size_t t = b->cfg.tasks[l].dof;
                    r = dwbc_batch_set_task(b, l, b->d_fstar + off + s0 * t, nullptr, 0, sk);
                    if (r) return r;
                    off += t * B;
                }
            }
            r = dwbc_batch_solve(b, stages, nullptr, nullptr, nullptr, nullptr, sk);
            if (r) return r;
            // results go to the host straight from the workspace records of the slice
            if (tau_grav_host) CK(cudaMemcpyAsync(tau_grav_host + s0 * m, b->w.tau_grav, sizeof(double) * cnt * m, cudaMemcpyDeviceToHost, sk));
            if (tau_task_host) CK(cudaMemcpyAsync(tau_task_host + s0 * m, b->w.tau_task, sizeof(double) * cnt * m, cudaMemcpyDeviceToHost, sk));
            if (tau_contact_host) CK(cudaMemcpyAsync(tau_contact_host + s0 * m, b->w.tau_contact, sizeof(double) * cnt * m, cudaMemcpyDeviceToHost, sk));
            if (status_host) CK(cudaMemcpyAsync(status_host + s0, b->w.status, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, sk));
            return DWBC_OK;
        }();
    }
    // back to the whole batch: the handle keeps describing all B states (get_matrix, diagnostics, later device-pointer calls)
    b->w = full;
    b->B = (int)B;
    b->last_qd = b->d_qd;
    if (need_tasks) {
        size_t off = 0;
        for (int l = 0; l < b->cfg.nlevels; l++) {
            b->w.fstar[l] = b->d_fstar + off;
            off += (size_t)b->cfg.tasks[l].dof * B;
        }
    }
    if (K > 1) {
        for (int i = 0; i < 2; i++) {
            cudaEventRecord(b->pipe_join[i], b->pipe_stream[i]);
            cudaStreamWaitEvent(st, b->pipe_join[i], 0);
        }
    }
    CK(cudaStreamSynchronize(st));
    return rc;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ a9: reduced-dimension variant
namespace {
void free_reduced_tasks(dwbc_batch *b) {
    for (void *p : b->red_task_allocs) cudaFree(p);
    b->red_task_allocs.clear();
    b->red_task_sig.clear();
    for (int l = 0; l < MAX_LEVELS; l++) {
        b->wr.Jt[l] = b->wr.Lamt[l] = b->wr.Jkt[l] = b->wr.Null[l] = b->wr.Ntq[l] = b->wr.torque_h[l] = nullptr;
        b->wr.fstar_qp[l] = b->wr.contact_qp[l] = nullptr;
        b->wr.active[l] = nullptr;
        b->rx.JtR[l] = b->rx.LamNC[l] = nullptr;
    }
    b->rx.fonR = b->rx.tprevNC = b->rx.NtqNC = b->rx.fqpNC = nullptr;
    b->rx.stNC = nullptr;
    b->d_cfgNC = nullptr;
}
void free_reduced(dwbc_batch *b) {
    free_reduced_tasks(b);
    for (void *p : b->red_allocs) cudaFree(p);
    b->red_allocs.clear();
    b->lqr = LqpOut{};
    b->lqr_allocated = b->have_lqp_r = false;
    b->lqnc_y = b->lqnc_tau_nc = b->lqnc_tau = nullptr;
    b->lqnc_status = b->lqnc_iters = nullptr;
    b->have_lqp_r_nc = false;
    b->red_rn = 0;
    b->d_cfgR = nullptr;
}
int alloc_reduced(dwbc_batch *b, int rn, int rm) {
    if (b->red_rn == rn) return DWBC_OK;
    free_reduced(b);
    Work &r = b->wr;
    const Work &w = b->w;
    r = Work();
    std::memset(&r, 0, sizeof(r));
    r.B = w.B; r.n = rn; r.m = rm; r.nb = w.nb;
    r.A = w.A; r.Bnl = w.Bnl; r.pose = w.pose; r.com = w.com; r.linkvel = w.linkvel; r.cdof = w.cdof; r.cidx = w.cidx;  // shared with the full model
    const size_t B = w.B, n = rn, m = rm;
    std::vector<void *> *L = &b->red_allocs;
    ReducedX &x = b->rx;
    bool ok = dev_alloc(b, &r.Ainv, B * n * n, L) && dev_alloc(b, &r.G, B * n, L) && dev_alloc(b, &r.Jc, B * CMAX * n, L) &&
              dev_alloc(b, &r.Lam, B * CMAX * CMAX, L) && dev_alloc(b, &r.Jbar, B * CMAX * n, L) && dev_alloc(b, &r.AinvNc, B * n * n, L) &&
              dev_alloc(b, &r.W, B * m * m, L) && dev_alloc(b, &r.Winv, B * m * m, L) && dev_alloc(b, &r.V2, B * RMAX * m, L) &&
              dev_alloc(b, &r.NwJw, B * m * RMAX, L) && dev_alloc(b, &r.PC, B * CMAX, L) && dev_alloc(b, &r.tau_grav, B * m, L) &&
              dev_alloc(b, &r.tau_task, B * m, L) && dev_alloc(b, &r.tau_contact, B * m, L) && dev_alloc(b, &r.status, B, L) &&
              dev_alloc(b, &x.JI, B * 6 * NCMAX, L) && dev_alloc(b, &x.JIinvT, B * 6 * NCMAX, L) && dev_alloc(b, &x.SInc, B * 16, L) &&
              dev_alloc(b, &x.AR, B * n * n, L) && dev_alloc(b, &x.Jbkt, B * m * 6, L) && dev_alloc(b, &x.tauR, B * m, L) &&
              dev_alloc(b, &x.tauNC, B * NCMAX, L) && dev_alloc(b, &b->d_cfgR, 1, L) && dev_alloc(b, &r.active[MAX_LEVELS], B * ACTIVE_WORDS, L) &&
              dev_alloc(b, &r.cf_redis, B * RMAX, L) && dev_alloc(b, &r.qp_iters, B, L);
    if (!ok) {
        free_reduced(b);
        return DWBC_ERR_CUDA;
    }
    b->red_rn = rn;
    return DWBC_OK;
}
}  // namespace

extern "C" {

int dwbc_batch_reduced_dynamics(dwbc_batch_t *b, const uint8_t *pattern_host, void *stream) {
    if (!b || !pattern_host) return fail(DWBC_ERR_ARG, "reduced_dynamics: null argument");
    if (!b->have_contact) return fail(DWBC_ERR_ORDER, "reduced_dynamics: call update_kinematics and set_contact first");
    NEED_ALL(b, "reduced_dynamics");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const DevModel &md = b->hm.dev;
    // index sets of the pattern, reference src/dwbc.cpp:2755-2823
    ReducedCfg rc;
    std::memset(&rc, 0, sizeof(rc));
    std::vector<char> co(md.nb, 0);
    co[0] = 1;
    for (int i = 0; i < b->cfg.ncontact; i++) {
        if (!pattern_host[i]) continue;
        if (rc.ncontact_active >= MAX_ACTIVE_CONTACTS) return fail(DWBC_ERR_LIMIT, "reduced_dynamics: too many active contacts in the pattern");
        rc.pattern[rc.ncontact_active++] = i;
        for (int l = b->cfg.contacts[i].link; l != 0; l = md.parent[l]) co[l] = 1;
    }
    if (rc.ncontact_active == 0) return fail(DWBC_ERR_ARG, "reduced_dynamics: the pattern has no active contact");
    int co_dof = 0;
    for (int l = 1; l < md.nb; l++) co_dof += co[l];
    for (int l = 1; l < md.nb; l++)
        if ((co[l] != 0) != (l <= co_dof))
            return fail(DWBC_ERR_ARG,
                        "reduced_dynamics: the contact-chain joints are not a prefix of q (the reference's block indexing needs [base | contact chain | "
                        "non-contact chain]: double support or left-foot single support)");
    rc.co_dof = co_dof;
    rc.vc_dof = co_dof + 6;
    rc.nc_dof = md.n - rc.vc_dof;
    rc.rm = co_dof + 6;
    rc.rn = rc.rm + 6;
    for (int l = 0; l < md.nb; l++)
        if (!co[l]) rc.nc_mask |= 1ull << l;
    if (rc.nc_dof <= 0) return fail(DWBC_ERR_ARG, "reduced_dynamics: empty non-contact chain");
    int rcode = alloc_reduced(b, rc.rn, rc.rm);
    if (rcode) return rcode;
    b->rc = rc;
    {
        size_t smem = sizeof(double) * reduced_dyn_smem_doubles(rc.rn, rc.nc_dof);
        CK(cudaFuncSetAttribute(reduced_dynamics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelScope ks(b, DWBC_K_REDUCED, st);
        DWBC_LAUNCH(reduced_dynamics_kernel, b->B, TEAM, smem, st, b->d_model, b->w, b->wr, b->rx, b->rc);
    }
    CK(cudaGetLastError());
    b->have_red_dyn = true;
    b->have_red_cc = false;
    return DWBC_OK;
}

int dwbc_batch_reduced_calc_contact_constraint(dwbc_batch_t *b, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "reduced_calc_contact_constraint: null handle");
    if (!b->have_red_dyn) return fail(DWBC_ERR_ORDER, "reduced_calc_contact_constraint: call reduced_dynamics first");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    ContactSmem L(b->wr.n, b->wr.m);
    size_t smem = sizeof(double) * L.total;
    CK(cudaFuncSetAttribute(contact_dynamics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelScope ks(b, DWBC_K_CONTACT_DYN, st);
        DWBC_LAUNCH(contact_dynamics_kernel, b->B, TEAM, smem, st, b->d_model, b->wr);
    }
    CK(cudaGetLastError());
    b->have_red_cc = true;
    return DWBC_OK;
}

int dwbc_batch_set_reduced_torque_limit(dwbc_batch_t *b, const double *tlim_host, int count) {
    if (!b) return fail(DWBC_ERR_ARG, "set_reduced_torque_limit: null handle");
    if (!tlim_host || count <= 0) {
        b->red_tlim_set = false;
        b->red_tlim.clear();
        return DWBC_OK;
    }
    if (count > MAXB) return fail(DWBC_ERR_LIMIT, "set_reduced_torque_limit: too many entries");
    b->red_tlim.assign(tlim_host, tlim_host + count);
    b->red_tlim_set = true;
    return DWBC_OK;
}

int dwbc_batch_reduced_dims(const dwbc_batch_t *b, int *co_dof, int *nc_dof, int *vc_dof, int *rsys, int *rmod) {
    if (!b || !b->have_red_dyn) return fail(DWBC_ERR_ORDER, "reduced_dims: call reduced_dynamics first");
    if (co_dof) *co_dof = b->rc.co_dof;
    if (nc_dof) *nc_dof = b->rc.nc_dof;
    if (vc_dof) *vc_dof = b->rc.vc_dof;
    if (rsys) *rsys = b->rc.rn;
    if (rmod) *rmod = b->rc.rm;
    return DWBC_OK;
}

int dwbc_batch_reduced_solve(dwbc_batch_t *b, int stages, double *tau_grav_dev, double *tau_task_dev, double *tau_contact_dev, int32_t *status_dev,
                             void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "reduced_solve: null handle");
    if (!b->have_red_cc) return fail(DWBC_ERR_ORDER, "reduced_solve: call reduced_calc_contact_constraint first");
    if ((stages & DWBC_SOLVE_REDIST_ANALYTIC) && (stages & DWBC_SOLVE_REDIST)) stages &= ~DWBC_SOLVE_REDIST_ANALYTIC;  // as dwbc_batch_solve
    if (stages & DWBC_SOLVE_HQP) stages |= DWBC_SOLVE_TASK;
    if ((stages & (DWBC_SOLVE_REDIST | DWBC_SOLVE_REDIST_ANALYTIC)) && !(stages & DWBC_SOLVE_TASK))
        return fail(DWBC_ERR_ARG, "reduced_solve: the contact redistribution needs DWBC_SOLVE_TASK");
    const bool use_qp = (stages & (DWBC_SOLVE_HQP | DWBC_SOLVE_REDIST)) != 0;
    if (use_qp && b->red_tlim_set && (int)b->red_tlim.size() != b->rc.rm)
        return fail(DWBC_ERR_ARG, "reduced_solve: task qp torque limit size is not matched with model size (reduced_model_dof_)");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rcode = upload_cfg(b, st);
    if (rcode) return rcode;
    const DevModel &md = b->hm.dev;
    const size_t B = b->B, m = b->w.m;
    ReducedCfg &rc = b->rc;
    Work &wr = b->wr;
    const int rn = rc.rn, rm = rc.rm;
    rc.nlevels = 0;
    rc.k_contact = 0;
    int tmax = 1, tmaxR = 1;
    CK(cudaMemsetAsync(wr.tau_task, 0, sizeof(double) * B * rm, st));
    if (stages & DWBC_SOLVE_TASK) {
        const BatchCfg &c = b->cfg;
        if (c.nlevels == 0) return fail(DWBC_ERR_ORDER, "reduced_solve: no task registered");
        for (int l = 0; l < c.nlevels; l++)
            if (!b->w.fstar[l]) return fail(DWBC_ERR_ORDER, "reduced_solve: set_task has not been called for every level");
        rcode = alloc_tasks(b);
        if (rcode) return rcode;
        // task classification, reference src/dwbc.cpp:3166-3227
        std::string sig = std::to_string(rn);
        for (int l = 0; l < c.nlevels; l++) {
            const TaskCfg &t = c.tasks[l];
            bool cmm = false, red = false, non = false;
            if (t.custom) {
                if (b->w.jcustom_per_state[l])
                    return fail(DWBC_ERR_ARG, "reduced_solve: a custom task with per-state Jacobians cannot be classified for the whole batch");
                std::vector<double> J((size_t)t.dof * md.n);
                CK(cudaMemcpyAsync(J.data(), b->w.jcustom[l], sizeof(double) * J.size(), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                double nl2 = 0.0, nr2 = 0.0;
                for (int j = 0; j < md.n; j++)
                    for (int i = 0; i < t.dof; i++) (j < rc.vc_dof ? nl2 : nr2) += J[i + (size_t)j * t.dof] * J[i + (size_t)j * t.dof];
                red = std::sqrt(nl2) > 1.0e-4;
                non = std::sqrt(nr2) > 1.0e-4;
            } else {
                for (int k = 0; k < t.link_count; k++) {
                    const int link = c.tlinks[t.link_start + k].link;
                    if (link == md.nb) cmm = true;
                    else if (!((rc.nc_mask >> link) & 1ull)) red = true;
                    else non = true;
                }
            }
            int cls;
            if (cmm) cls = 0;  // src/task.cpp:105: cmm_task wins
            else if (red && !non) cls = 1;
            else if (!red && non) cls = 2;
            else return fail(DWBC_ERR_ARG, "UNDEFINED TASK TYPE (a level mixes contact-chain and non-contact links, reference src/task.cpp:129-136)");
            rc.cls[l] = cls;
            sig += "," + std::to_string(t.dof) + ":" + std::to_string(cls);
            tmax = t.dof > tmax ? t.dof : tmax;
        }
        rc.nlevels = c.nlevels;
        int k = 0;
        while (k < c.nlevels && rc.cls[k] != 2) k++;
        for (int l = k; l < c.nlevels; l++)
            if (rc.cls[l] != 2) return fail(DWBC_ERR_ARG, "reduced_solve: COM / contact-chain levels must precede the non-contact levels");
        if (k == 0) return fail(DWBC_ERR_ARG, "reduced_solve: the first level cannot be a non-contact task (the reference reads ts_[-1])");
        rc.k_contact = k;
        if (sig != b->red_task_sig) {
            free_reduced_tasks(b);
            std::vector<void *> *L = &b->red_task_allocs;
            for (int l = 0; l < c.nlevels; l++) {
                const size_t t = c.tasks[l].dof;
                bool ok;
                if (l < k)
                    ok = dev_alloc(b, &wr.Jt[l], B * t * rn, L) && dev_alloc(b, &wr.Lamt[l], B * t * t, L) && dev_alloc(b, &wr.Jkt[l], B * rm * t, L) &&
                         dev_alloc(b, &wr.Null[l], B * rm * rm, L) && dev_alloc(b, &wr.Ntq[l], B * rm * t, L) && dev_alloc(b, &wr.torque_h[l], B * rm, L) &&
                         dev_alloc(b, &b->rx.JtR[l], B * t * rn, L) && dev_alloc(b, &wr.fstar_qp[l], B * t, L) && dev_alloc(b, &wr.contact_qp[l], B * RMAX, L) &&
                         dev_alloc(b, &wr.active[l], B * ACTIVE_WORDS, L);
                else
                    ok = dev_alloc(b, &b->rx.LamNC[l], B * t * t, L);
                if (!ok) return DWBC_ERR_CUDA;
            }
            if (k < c.nlevels) {  // workspace of CalcSingleTaskTorqueWithQP_R_NC (one 6-dof pseudo level for hqp_kernel)
                ReducedX &x = b->rx;
                Work &q = b->wq;
                q = Work();
                std::memset(&q, 0, sizeof(q));
                bool ok = dev_alloc(b, &x.fonR, B * 6, L) && dev_alloc(b, &x.tprevNC, B * rm, L) && dev_alloc(b, &x.NtqNC, B * rm * 6, L) &&
                          dev_alloc(b, &x.stNC, B, L) && dev_alloc(b, &x.fqpNC, B * 6, L) && dev_alloc(b, &q.tau_task, B * rm, L) &&
                          dev_alloc(b, &q.tau_contact, B * rm, L) && dev_alloc(b, &q.contact_qp[0], B * RMAX, L) &&
                          dev_alloc(b, &q.active[0], B * ACTIVE_WORDS, L) && dev_alloc(b, &q.torque_h[0], B * rm, L) &&
                          dev_alloc(b, &q.Lamt[0], B * 36, L) && dev_alloc(b, &q.Jkt[0], B * rm * 6, L) && dev_alloc(b, &q.qp_iters, B, L) &&
                          dev_alloc(b, &b->d_cfgNC, 1, L);
                if (!ok) return DWBC_ERR_CUDA;
            }
            b->red_task_sig = sig;
        }
        // reduced hierarchy = the first k levels with per-state custom Jacobians
        b->cfgR = b->cfg;
        b->cfgR.nlevels = k;
        b->cfgR.tlim_set = b->red_tlim_set ? 1 : 0;  // the full-model limit has the wrong size for the reduced QP (src/dwbc.cpp:3462-3468)
        for (int i = 0; i < (int)b->red_tlim.size() && i < MAXB; i++) b->cfgR.tlim[i] = b->red_tlim[i];
        for (int l = 0; l < k; l++) {
            b->cfgR.tasks[l].custom = 1;
            b->cfgR.tasks[l].link_count = 0;
            wr.jcustom[l] = b->rx.JtR[l];
            wr.jcustom_per_state[l] = 1;
            wr.fstar[l] = b->w.fstar[l];
            tmaxR = c.tasks[l].dof > tmaxR ? c.tasks[l].dof : tmaxR;
        }
        CK(cudaMemcpyAsync(b->d_cfgR, &b->cfgR, sizeof(BatchCfg), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        {
            size_t smem = sizeof(double) * (size_t)(tmaxR | 1) * md.n;
            CK(cudaFuncSetAttribute(reduced_task_jacobian_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            KernelScope ks(b, DWBC_K_REDUCED, st);
            DWBC_LAUNCH(reduced_task_jacobian_kernel, b->B, TEAM, smem, st, b->d_model, b->d_cfg, b->w, b->wr, b->rx, b->rc, tmaxR);
        }
        {
            TaskSmem L(rn, rm, b->cfgR.tasks, b->cfgR.nlevels);
            size_t smem = sizeof(double) * L.total;
            CK(cudaFuncSetAttribute(task_space_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            KernelScope ks(b, DWBC_K_TASK, st);
            DWBC_LAUNCH(task_space_kernel, b->B, TEAM, smem, st, b->d_model, b->d_cfgR, b->wr, (stages & DWBC_SOLVE_HQP) ? 1 : 0, k < c.nlevels ? 1 : 0);
        }
        if (use_qp) {
            // CalcSingleTaskTorqueWithQP_R / CalcContactRedistributeR: the same QPs on the reduced matrices (src/dwbc.cpp:3448-3599, 4777-4925)
            int nvmax = RMAX > 0 ? RMAX : 1;
            for (int l = 0; l < k; l++) nvmax = c.tasks[l].dof + RMAX > nvmax ? c.tasks[l].dof + RMAX : nvmax;
            HqpSmem L(rn, rm, nvmax);
            size_t smem = sizeof(double) * L.total;
            CK(cudaFuncSetAttribute(hqp_kernel_t<0, 0, 0, -1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            // with non-contact levels the redistribution needs torque_task_R_ after their sums: it runs after reduced_nc_kernel
            const int redist_now = ((stages & DWBC_SOLVE_REDIST) && k == c.nlevels) ? 2 : 0;
            if ((stages & DWBC_SOLVE_HQP) || redist_now) {
                KernelScope ks(b, DWBC_K_HQP, st);
                DWBC_LAUNCH((hqp_kernel_t<0, 0, 0, -1, -1>), b->B, TEAM, smem, st, b->d_model, b->d_cfgR, b->wr, (stages & DWBC_SOLVE_HQP) ? 1 : 0, redist_now, nvmax, (const int *)nullptr, (const int *)nullptr);
            }
        }
    }
    {
        ReducedNcSmem L(md.n, rm, tmax);
        size_t smem = sizeof(double) * L.total;
        CK(cudaFuncSetAttribute(reduced_nc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        KernelScope ks(b, DWBC_K_REDUCED, st);
        DWBC_LAUNCH(reduced_nc_kernel, b->B, TEAM, smem, st, b->d_model, b->d_cfg, b->w, b->wr, b->rx, b->rc, tmax,
                    ((stages & DWBC_SOLVE_HQP) ? 1 : 0) | (((stages & DWBC_SOLVE_REDIST) && rc.k_contact == rc.nlevels) ? 2 : 0));
    }
    if ((stages & DWBC_SOLVE_HQP) && rc.k_contact < rc.nlevels) {
        // CalcSingleTaskTorqueWithQP_R_NC: the level QP once more, on a 6-dof pseudo level whose f*, Ntorque and torque_prev were
        // prepared by reduced_nc_kernel (zero Lambda / J_kt: the pseudo level has no torque_h of its own)
        Work q = b->wq;
        const Work &r = b->wr;
        q.B = r.B; q.n = r.n; q.m = r.m; q.nb = r.nb;
        q.pose = r.pose; q.cdof = r.cdof; q.cidx = r.cidx; q.Jbar = r.Jbar; q.PC = r.PC; q.NwJw = r.NwJw;
        q.tau_grav = b->rx.tprevNC;
        q.status = b->rx.stNC;
        q.Ntq[0] = b->rx.NtqNC;
        q.fstar[0] = b->rx.fonR;
        q.fstar_qp[0] = b->rx.fqpNC;
        BatchCfg cn = b->cfgR;
        cn.nlevels = 1;
        cn.tasks[0].custom = 1;
        cn.tasks[0].dof = 6;
        cn.tasks[0].link_count = 0;
        CK(cudaMemcpyAsync(b->d_cfgNC, &cn, sizeof(BatchCfg), cudaMemcpyHostToDevice, st));
        CK(cudaStreamSynchronize(st));
        const int nvmax = 6 + RMAX;
        HqpSmem L(rn, rm, nvmax);
        size_t smem = sizeof(double) * L.total;
        CK(cudaFuncSetAttribute(hqp_kernel_t<0, 0, 0, -1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            KernelScope ks(b, DWBC_K_HQP, st);
            DWBC_LAUNCH((hqp_kernel_t<0, 0, 0, -1, -1>), b->B, TEAM, smem, st, b->d_model, b->d_cfgNC, q, 1, 0, nvmax, (const int *)nullptr, (const int *)nullptr);
        }
        {
            KernelScope ks(b, DWBC_K_REDUCED, st);
            DWBC_LAUNCH(reduced_ncqp_apply_kernel, b->B, 64, 0, st, b->w, b->rx, b->rc);
        }
    }
    if (stages & DWBC_SOLVE_REDIST_ANALYTIC) {
        // analytic branch of CalcContactRedistributeR (reference src/dwbc.cpp:4938-4990): the closed form of
        // CalcContactRedistribute(hqp = false) on J_CR_INV_T, P_CR, V2_R with torque_input = torque_grav_R_ + torque_task_R_
        Work d = b->wr;
        d.tau_task = b->rx.tauR;
        CK(cudaMemsetAsync(d.tau_contact, 0, sizeof(double) * B * rm, st));
        const int wpb = WARPS_PER_BLOCK, grid = grid_for(b->B, wpb);
        const size_t smem = sizeof(double) * wpb * (MAXB + CMAX + 6 * 7 + 8);
        {
            KernelScope ks(b, DWBC_K_POST, st);
            DWBC_LAUNCH(redistribute_analytic_kernel, grid, wpb * 32, smem, st, b->d_model, b->d_cfgR, d);
        }
        {
            KernelScope ks(b, DWBC_K_REDUCED, st);
            DWBC_LAUNCH(reduced_redist_apply_kernel, b->B, 64, 0, st, b->w, b->wr, b->rc);
        }
    }
    if ((stages & DWBC_SOLVE_REDIST) && rc.k_contact < rc.nlevels) {
        // CalcContactRedistributeR(torque_grav_R_ + torque_task_R_) with the final torque_task_R_ (reference src/dwbc.cpp:3758-3770)
        Work d = b->wr;
        d.tau_task = b->rx.tauR;
        const int nvmax = RMAX > 0 ? RMAX : 1;
        HqpSmem L(rn, rm, nvmax);
        size_t smem = sizeof(double) * L.total;
        CK(cudaFuncSetAttribute(hqp_kernel_t<0, 0, 0, -1, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            KernelScope ks(b, DWBC_K_HQP, st);
            DWBC_LAUNCH((hqp_kernel_t<0, 0, 0, -1, -1>), b->B, TEAM, smem, st, b->d_model, b->d_cfgR, d, 0, 2, nvmax, (const int *)nullptr, (const int *)nullptr);
        }
        {
            KernelScope ks(b, DWBC_K_REDUCED, st);
            DWBC_LAUNCH(reduced_redist_apply_kernel, b->B, 64, 0, st, b->w, b->wr, b->rc);
        }
    }
    CK(cudaGetLastError());
    if (tau_grav_dev) CK(cudaMemcpyAsync(tau_grav_dev, b->w.tau_grav, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (tau_task_dev) CK(cudaMemcpyAsync(tau_task_dev, b->w.tau_task, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (tau_contact_dev) CK(cudaMemcpyAsync(tau_contact_dev, b->w.tau_contact, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (status_dev) CK(cudaMemcpyAsync(status_dev, b->w.status, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    return DWBC_OK;
}

int dwbc_batch_get_reduced_matrix(dwbc_batch_t *b, int which, int level, double *out_dev, int *rows, int *cols, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "get_reduced_matrix: null handle");
    if (!b->have_red_dyn) return fail(DWBC_ERR_ORDER, "get_reduced_matrix: call reduced_dynamics first");
    CK(cudaSetDevice(b->device));
    const Work &r = b->wr;
    const ReducedX &x = b->rx;
    const int rn = b->rc.rn, rm = b->rc.rm, nc = b->rc.nc_dof;
    const double *src = nullptr;
    int rr = 0, cc = 0, stride = 0;  // stride != 0: the per-state record is larger than rr*cc
    int t = 0;
    if (which >= DWBC_RMAT_J_TASK_R) {
        if (level < 0 || level >= b->rc.nlevels) return fail(DWBC_ERR_ARG, "get_reduced_matrix: level out of range (or reduced_solve not called)");
        t = b->cfg.tasks[level].dof;
    }
    const bool contact_level = level < b->rc.k_contact;
    switch (which) {
    case DWBC_RMAT_A_R_INV: src = r.Ainv; rr = rn; cc = rn; break;
    case DWBC_RMAT_A_R: src = x.AR; rr = rn; cc = rn; break;
    case DWBC_RMAT_J_I_NC: src = x.JI; rr = 6; cc = nc; stride = 6 * NCMAX; break;
    case DWBC_RMAT_J_I_NC_INV_T: src = x.JIinvT; rr = 6; cc = nc; stride = 6 * NCMAX; break;
    case DWBC_RMAT_G_R: src = r.G; rr = rn; cc = 1; break;
    case DWBC_RMAT_J_CR: src = r.Jc; rr = CMAX; cc = rn; break;
    case DWBC_RMAT_LAMBDA_CR: src = r.Lam; rr = CMAX; cc = CMAX; break;
    case DWBC_RMAT_J_CR_INV_T: src = r.Jbar; rr = CMAX; cc = rn; break;
    case DWBC_RMAT_A_R_INV_N_CR: src = r.AinvNc; rr = rn; cc = rn; break;
    case DWBC_RMAT_W_R: src = r.W; rr = rm; cc = rm; break;
    case DWBC_RMAT_W_R_INV: src = r.Winv; rr = rm; cc = rm; break;
    case DWBC_RMAT_NWJW_R: src = r.NwJw; rr = rm; cc = RMAX; break;
    case DWBC_RMAT_P_CR: src = r.PC; rr = CMAX; cc = 1; break;
    case DWBC_RMAT_TORQUE_GRAV_R: src = r.tau_grav; rr = rm; cc = 1; break;
    case DWBC_RMAT_TORQUE_TASK_R: src = x.tauR; rr = rm; cc = 1; break;
    case DWBC_RMAT_J_BASE_R_KT: src = x.Jbkt; rr = rm; cc = 6; break;
    case DWBC_RMAT_SI_NC: src = x.SInc; rr = 13; cc = 1; stride = 16; break;
    case DWBC_RMAT_J_TASK_R: src = contact_level ? x.JtR[level] : nullptr; rr = t; cc = rn; break;
    case DWBC_RMAT_LAMBDA_TASK: src = contact_level ? r.Lamt[level] : x.LamNC[level]; rr = t; cc = t; break;
    case DWBC_RMAT_J_KT_R: src = contact_level ? r.Jkt[level] : nullptr; rr = rm; cc = t; break;
    case DWBC_RMAT_NULL_TASK_R: src = contact_level ? r.Null[level] : nullptr; rr = rm; cc = rm; break;
    default: return fail(DWBC_ERR_ARG, "get_reduced_matrix: unknown matrix id");
    }
    if (rows) *rows = rr;
    if (cols) *cols = cc;
    if (!out_dev) return DWBC_OK;
    if (!src) return fail(DWBC_ERR_ARG, "get_reduced_matrix: not defined for this level");
    cudaStream_t st = (cudaStream_t)stream;
    if (!stride)
        CK(cudaMemcpyAsync(out_dev, src, sizeof(double) * (size_t)b->B * rr * cc, cudaMemcpyDeviceToDevice, st));
    else
        CK(cudaMemcpy2DAsync(out_dev, sizeof(double) * rr * cc, src, sizeof(double) * stride, sizeof(double) * rr * cc, b->B, cudaMemcpyDeviceToDevice, st));
    return DWBC_OK;
}


// ---- a10: LQP variant ----
namespace {
// shared by ConfigureLQP + CalcControlTorqueLQP (full model) and ConfigureLQP_R + CalcControlTorqueLQP_R (reduced model)
int launch_lqp(dwbc_batch *b, bool reduced, double *y_dev, double *tau_dev, int32_t *status_dev, cudaStream_t st) {
    const BatchCfg &cfg = reduced ? b->cfgR : b->cfg;
    Work w = reduced ? b->wr : b->w;
    LqpOut &lq = reduced ? b->lqr : b->lq;
    bool &allocated = reduced ? b->lqr_allocated : b->lqp_allocated;
    if (reduced) {
        w.A = b->rx.AR;   // A_R (the reduced workspace shares the full A_ for the contact stage)
        w.Bnl = b->wr.G;  // ConfigureLQP_R uses the gravity vector G_R only (reference src/dwbc.cpp:4526, 4538-4539, 4546)
    }
    const size_t B = b->B, n = w.n, m = w.m;
    if (n + CMAX - 6 > LQP_WCAP || (n + CMAX - 6) * (n + CMAX - 5) / 2 > (size_t)LQP_CHOL_OWN * TEAM)
        return fail(DWBC_ERR_ARG, "calc_control_torque_lqp: model too large for the LQP kernel (n + cmax - 6 <= 48)");
    if (!allocated) {
        std::vector<void *> *L = reduced ? &b->red_allocs : nullptr;
        if (!dev_alloc(b, &lq.y, B * LQP_MAX_LEVELS * (n + CMAX), L) || !dev_alloc(b, &lq.v1, B * (2 * m + 10 * MAX_ACTIVE_CONTACTS), L) ||
            !dev_alloc(b, &lq.tau, B * m, L) || !dev_alloc(b, &lq.status, B, L) || !dev_alloc(b, &lq.iters, B, L))
            return DWBC_ERR_CUDA;
        allocated = true;
    }
    int tmax = 1;
    for (int l = 0; l < cfg.nlevels; l++) tmax = cfg.tasks[l].dof > tmax ? cfg.tasks[l].dof : tmax;
    LqpParams prm{};
    prm.anorm_src = b->w.A;
    prm.anorm_n = b->w.n;
    for (size_t i = 0; i < m && i < (size_t)MAXB; i++) prm.tlim[i] = LQP_TORQUE_LIMIT;
    if (reduced && m >= 4) prm.tlim[m - 4] = 600.0;  // reference src/dwbc.cpp:4529-4531
    const LqpSmem L((int)n, (int)m, tmax);
    const size_t smem = sizeof(double) * L.total;
    CK(cudaFuncSetAttribute(lqp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelScope ks(b, DWBC_K_LQP, st);
        DWBC_LAUNCH(lqp_kernel, b->B, TEAM, smem, st, b->d_model, reduced ? b->d_cfgR : b->d_cfg, w, lq, tmax, prm);
    }
    CK(cudaGetLastError());
    const int nlev = 2 + cfg.nlevels;
    if (y_dev)
        CK(cudaMemcpy2DAsync(y_dev, sizeof(double) * (n + CMAX), lq.y + (size_t)(nlev - 1) * (n + CMAX), sizeof(double) * LQP_MAX_LEVELS * (n + CMAX),
                             sizeof(double) * (n + CMAX), B, cudaMemcpyDeviceToDevice, st));
    if (tau_dev) CK(cudaMemcpyAsync(tau_dev, lq.tau, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (status_dev) CK(cudaMemcpyAsync(status_dev, lq.status, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    return DWBC_OK;
}
}  // namespace

int dwbc_batch_calc_control_torque_lqp(dwbc_batch_t *b, double *y_dev, double *tau_dev, int32_t *status_dev, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "calc_control_torque_lqp: null handle");
    if (!b->have_cc) return fail(DWBC_ERR_ORDER, "calc_control_torque_lqp: call calc_contact_constraint first");
    NEED_ALL(b, "calc_control_torque_lqp");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_cfg(b, st);
    if (rc) return rc;
    for (int l = 0; l < b->cfg.nlevels; l++)
        if (!b->w.fstar[l]) return fail(DWBC_ERR_ORDER, "calc_control_torque_lqp: set_task has not been called for every level");
    rc = launch_lqp(b, false, y_dev, tau_dev, status_dev, st);
    if (rc) return rc;
    b->have_lqp = true;
    return DWBC_OK;
}

int dwbc_batch_calc_control_torque_lqp_r(dwbc_batch_t *b, double *y_dev, double *tau_dev, int32_t *status_dev, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "calc_control_torque_lqp_r: null handle");
    if (!b->have_red_cc || b->rc.k_contact <= 0)
        return fail(DWBC_ERR_ORDER, "calc_control_torque_lqp_r: call reduced_dynamics, reduced_calc_contact_constraint and reduced_solve(DWBC_SOLVE_TASK) first");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_lqp(b, true, y_dev, tau_dev, status_dev, st);
    if (rc) return rc;
    b->have_lqp_r = true;
    return DWBC_OK;
}

int dwbc_batch_calc_control_torque_lqp_r_nc(dwbc_batch_t *b, double *qdd_nc_dev, double *tau_dev, int32_t *status_dev, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "calc_control_torque_lqp_r_nc: null handle");
    if (!b->have_lqp_r) return fail(DWBC_ERR_ORDER, "calc_control_torque_lqp_r_nc: call calc_control_torque_lqp_r first (its answer is the input)");
    const BatchCfg &c = b->cfg;
    // ConfigureLQP_R_NC reads ts_[1] as THE non-contact task (reference src/dwbc.cpp:4738-4745): a 6D link task on the non-contact chain
    if (c.nlevels < 2 || c.tasks[1].custom || c.tasks[1].dof != 6 || c.tasks[1].link_count != 1 || c.tlinks[c.tasks[1].link_start].dof != 6 ||
        b->rc.cls[1] != 2)
        return fail(DWBC_ERR_ARG, "calc_control_torque_lqp_r_nc: hierarchy level 1 must be one 6D link task on the non-contact chain (the reference hard-codes ts_[1])");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t B = b->B, m = b->w.m, nc = b->rc.nc_dof;
    if (nc + 6 > (size_t)LQP_WCAP || nc * (nc + 1) / 2 > (size_t)LQP_CHOL_OWN * TEAM) return fail(DWBC_ERR_LIMIT, "calc_control_torque_lqp_r_nc: non-contact chain too long");
    if (!b->lqnc_y) {
        std::vector<void *> *L = &b->red_allocs;
        if (!dev_alloc(b, &b->lqnc_y, B * 2 * nc, L) || !dev_alloc(b, &b->lqnc_tau_nc, B * nc, L) || !dev_alloc(b, &b->lqnc_tau, B * m, L) ||
            !dev_alloc(b, &b->lqnc_status, B, L) || !dev_alloc(b, &b->lqnc_iters, B, L))
            return DWBC_ERR_CUDA;
    }
    const int rn = b->rc.rn, nlevR = 2 + b->cfgR.nlevels;
    LqpNcArgs a{};
    a.stride_yR = LQP_MAX_LEVELS * (rn + CMAX);
    a.yR = b->lqr.y + (size_t)(nlevR - 1) * (rn + CMAX);
    a.rn = rn;
    a.tauR = b->lqr.tau;
    a.rm = b->rc.rm;
    a.co_dof = b->rc.co_dof;
    a.vc_dof = b->rc.vc_dof;
    a.nc = (int)nc;
    a.statusR = b->lqr.status;
    a.levelsR = nlevR;
    a.ync = b->lqnc_y;
    a.tau_nc = b->lqnc_tau_nc;
    a.tau = b->lqnc_tau;
    a.status = b->lqnc_status;
    a.iters = b->lqnc_iters;
    const LqpSmem L((int)nc, (int)nc, 6);
    const size_t smem = sizeof(double) * L.total;
    CK(cudaFuncSetAttribute(lqp_nc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        KernelScope ks(b, DWBC_K_LQP, st);
        DWBC_LAUNCH(lqp_nc_kernel, b->B, TEAM, smem, st, b->d_model, b->d_cfg, b->w, (const double *)b->rx.JI, a);
    }
    CK(cudaGetLastError());
    if (qdd_nc_dev) CK(cudaMemcpy2DAsync(qdd_nc_dev, sizeof(double) * nc, b->lqnc_y + nc, sizeof(double) * 2 * nc, sizeof(double) * nc, B, cudaMemcpyDeviceToDevice, st));
    if (tau_dev) CK(cudaMemcpyAsync(tau_dev, b->lqnc_tau, sizeof(double) * B * m, cudaMemcpyDeviceToDevice, st));
    if (status_dev) CK(cudaMemcpyAsync(status_dev, b->lqnc_status, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, st));
    b->have_lqp_r_nc = true;
    return DWBC_OK;
}

int dwbc_batch_get_lqp(dwbc_batch_t *b, int which, int level, double *out_dev, int *rows, int *cols, void *stream) {
    if (!b) return fail(DWBC_ERR_ARG, "get_lqp: null handle");
    const bool reduced = (which & DWBC_LQP_REDUCED) != 0;
    which &= ~DWBC_LQP_REDUCED;
    if (!(reduced ? b->have_lqp_r : b->have_lqp)) return fail(DWBC_ERR_ORDER, "get_lqp: call calc_control_torque_lqp[_r] first");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const Work &w = reduced ? b->wr : b->w;
    const LqpOut &lq = reduced ? b->lqr : b->lq;
    const size_t B = b->B, n = w.n, m = w.m, nvs = n + CMAX;
    const int nlev = 2 + (reduced ? b->cfgR.nlevels : b->cfg.nlevels);
    int rr = 0;
    switch (which) {
    case DWBC_LQP_Y:
        if (level < 0 || level >= nlev) return fail(DWBC_ERR_ARG, "get_lqp: level out of range");
        rr = (int)nvs;
        break;
    case DWBC_LQP_V1: rr = (int)(2 * m + 10 * MAX_ACTIVE_CONTACTS); break;
    case DWBC_LQP_TORQUE: rr = (int)m; break;
    default: return fail(DWBC_ERR_ARG, "get_lqp: unknown id");
    }
    if (rows) *rows = rr;
    if (cols) *cols = 1;
    if (!out_dev) return DWBC_OK;
    if (which == DWBC_LQP_Y)
        CK(cudaMemcpy2DAsync(out_dev, sizeof(double) * nvs, lq.y + (size_t)level * nvs, sizeof(double) * LQP_MAX_LEVELS * nvs, sizeof(double) * nvs, B,
                             cudaMemcpyDeviceToDevice, st));
    else
        CK(cudaMemcpyAsync(out_dev, which == DWBC_LQP_V1 ? lq.v1 : lq.tau, sizeof(double) * B * rr, cudaMemcpyDeviceToDevice, st));
    return DWBC_OK;
}

int dwbc_batch_get_lqp_iterations(dwbc_batch_t *b, int32_t *out_dev, void *stream) {
    if (!b || !out_dev) return fail(DWBC_ERR_ARG, "get_lqp_iterations: null argument");
    if (!b->have_lqp) return fail(DWBC_ERR_ORDER, "get_lqp_iterations: call calc_control_torque_lqp first");
    CK(cudaSetDevice(b->device));
    CK(cudaMemcpyAsync(out_dev, b->lq.iters, sizeof(int32_t) * b->B, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return DWBC_OK;
}


// ---- f* from trajectories, MuJoCo layout, float32 torque command ----
namespace {
int traj_slot(dwbc_batch *b, int level, int link_index, const char *who) {
    if (!b) return fail(DWBC_ERR_ARG, std::string(who) + ": null handle");
    if (level < 0 || level >= b->cfg.nlevels || b->cfg.tasks[level].custom) return fail(DWBC_ERR_ARG, std::string(who) + ": level out of range or TASK_CUSTOM");
    if (link_index < 0 || link_index >= b->cfg.tasks[level].link_count) return fail(DWBC_ERR_ARG, std::string(who) + ": task link index out of range");
    return b->cfg.tasks[level].link_start + link_index;
}
}  // namespace

int dwbc_batch_set_task_gain(dwbc_batch_t *b, int level, int link_index, const double *gains18_host) {
    const int k = traj_slot(b, level, link_index, "set_task_gain");
    if (k < 0) return k;
    if (!gains18_host) return fail(DWBC_ERR_ARG, "set_task_gain: null gains");
    std::memcpy(b->traj[k].gains, gains18_host, sizeof(double) * 18);
    return DWBC_OK;
}
int dwbc_batch_set_trajectory_quintic(dwbc_batch_t *b, int level, int link_index, const double *time_dev, const double *pos_init_dev,
                                      const double *vel_init_dev, const double *pos_desired_dev, const double *vel_desired_dev) {
    const int k = traj_slot(b, level, link_index, "set_trajectory_quintic");
    if (k < 0) return k;
    if (b->cfg.tlinks[k].mode > DWBC_TASK_LINK_POSITION_CUSTOM_FRAME) return fail(DWBC_ERR_ARG, "set_trajectory_quintic: the task link has no position rows");
    if (!time_dev || !pos_init_dev || !vel_init_dev || !pos_desired_dev || !vel_desired_dev) return fail(DWBC_ERR_ARG, "set_trajectory_quintic: null argument");
    auto &t = b->traj[k];
    t.pos_time = time_dev; t.pos_init = pos_init_dev; t.vel_init = vel_init_dev; t.pos_des = pos_desired_dev; t.vel_des = vel_desired_dev;
    return DWBC_OK;
}
int dwbc_batch_set_trajectory_rotation(dwbc_batch_t *b, int level, int link_index, const double *time_dev, const double *rot_init_dev,
                                       const double *rot_desired_dev) {
    const int k = traj_slot(b, level, link_index, "set_trajectory_rotation");
    if (k < 0) return k;
    const int mode = b->cfg.tlinks[k].mode;
    if (mode > DWBC_TASK_LINK_6D_CUSTOM_FRAME && mode < DWBC_TASK_LINK_ROTATION) return fail(DWBC_ERR_ARG, "set_trajectory_rotation: the task link has no rotation rows");
    if (!time_dev || !rot_init_dev || !rot_desired_dev) return fail(DWBC_ERR_ARG, "set_trajectory_rotation: null argument");
    auto &t = b->traj[k];
    t.rot_time = time_dev; t.rot_init = rot_init_dev; t.rot_des = rot_desired_dev;
    return DWBC_OK;
}
int dwbc_batch_calc_fstar(dwbc_batch_t *b, int level, const double *control_time_dev, double *fstar_dev, void *stream) {
    if (!b || !control_time_dev || !fstar_dev) return fail(DWBC_ERR_ARG, "calc_fstar: null argument");
    if (!b->have_kin) return fail(DWBC_ERR_ORDER, "calc_fstar: call update_kinematics first");
    NEED_ALL(b, "calc_fstar");
    if (level < 0 || level >= b->cfg.nlevels || b->cfg.tasks[level].custom) return fail(DWBC_ERR_ARG, "calc_fstar: level out of range or TASK_CUSTOM");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const TaskCfg &tc = b->cfg.tasks[level];
    TrajLevel lv{};
    lv.nlinks = tc.link_count;
    int row = 0;
    for (int i = 0; i < tc.link_count; i++) {
        const TaskLinkCfg &c = b->cfg.tlinks[tc.link_start + i];
        const auto &h = b->traj[tc.link_start + i];
        TrajLink &t = lv.links[i];
        t.mode = c.mode; t.link = c.link; t.row = row; t.dof = c.dof;
        std::memcpy(t.point, c.point, sizeof(t.point));
        std::memcpy(t.gains, h.gains, sizeof(t.gains));
        t.pos_time = h.pos_time; t.pos_init = h.pos_init; t.vel_init = h.vel_init; t.pos_des = h.pos_des; t.vel_des = h.vel_des;
        t.rot_time = h.rot_time; t.rot_init = h.rot_init; t.rot_des = h.rot_des;
        row += c.dof;
    }
    const long long total = (long long)b->B * lv.nlinks;
    const int threads = 128, grid = (int)((total + threads - 1) / threads);
    {
        KernelScope ks(b, DWBC_K_POST, st);
        DWBC_LAUNCH(fstar_pd_kernel, grid, threads, 0, st, b->d_model, b->w, b->last_qd, lv, tc.dof, control_time_dev, fstar_dev);
    }
    CK(cudaGetLastError());
    return DWBC_OK;
}

int dwbc_batch_update_kinematics_mujoco(dwbc_batch_t *b, const double *qpos_dev, const double *qvel_dev, void *stream) {
    if (!b || !qpos_dev || !qvel_dev) return fail(DWBC_ERR_ARG, "update_kinematics_mujoco: null argument");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = b->w.n;
    if (!b->mj_q && (!dev_alloc(b, &b->mj_q, (size_t)b->B * (n + 1)) || !dev_alloc(b, &b->mj_qd, (size_t)b->B * n))) return DWBC_ERR_CUDA;
    const long long total = (long long)b->B * (n + 1);
    const int threads = 256, grid = (int)((total + threads - 1) / threads);
    {
        KernelScope ks(b, DWBC_K_EXPORT, st);
        DWBC_LAUNCH(mujoco_to_rbdl_kernel, grid, threads, 0, st, b->B, n, qpos_dev, qvel_dev, b->mj_q, b->mj_qd);
    }
    return dwbc_batch_update_kinematics(b, b->mj_q, b->mj_qd, stream);
}

int dwbc_batch_get_torque_command_f32(dwbc_batch_t *b, float *out_dev, void *stream) {
    if (!b || !out_dev) return fail(DWBC_ERR_ARG, "get_torque_command_f32: null argument");
    if (!b->have_cc) return fail(DWBC_ERR_ORDER, "get_torque_command_f32: no solve has been run");
    CK(cudaSetDevice(b->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)b->B * b->w.m;
    const int threads = 256, grid = (int)((total + threads - 1) / threads);
    {
        KernelScope ks(b, DWBC_K_EXPORT, st);
        DWBC_LAUNCH(torque_command_f32_kernel, grid, threads, 0, st, total, b->w.tau_grav, b->w.tau_task, b->w.tau_contact, out_dev);
    }
    CK(cudaGetLastError());
    return DWBC_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ multi-GPU host path
// the batch shards embarrassingly -- contiguous slices [g B/G, (g+1) B/G) of every batch array, one host
// thread + one stream + one dwbc_batch_t per GPU, no collective; every GPU reads and writes its own slice of the caller's (pinned)
// host buffers and the host call returns when all slices are done ("the host gathers the results": the slices land in place).
#ifndef DWBC_CPU_EMUL
#include <condition_variable>
#include <mutex>
#include <thread>
#else
#include <condition_variable>
#include <mutex>
#include <thread>
#endif

struct dwbc_multi {
    struct Shard {
        dwbc_batch *b = nullptr;
        int device = 0;
        size_t first = 0, count = 0;
        cudaStream_t stream = nullptr;
        std::vector<void *> dev_allocs;  // custom-task Jacobians uploaded to this GPU
        // worker
        std::thread th;
        std::mutex mu;
        std::condition_variable cv;
        bool has_job = false, done = false, quit = false;
        int rc = 0;
        std::string err;
        std::function<int(Shard &)> job;
        double last_ms = 0.0;
    };
    std::vector<std::unique_ptr<Shard>> shards;
    int B = 0, nlevels = 0;
    int tdof[MAX_LEVELS] = {0, 0, 0, 0};
    int n = 0, m = 0, ncontact = 0;
};

namespace {
void multi_worker(dwbc_multi::Shard *s) {
    cudaSetDevice(s->device);
    std::unique_lock<std::mutex> lk(s->mu);
    while (true) {
        s->cv.wait(lk, [&] { return s->has_job || s->quit; });
        if (s->quit) return;
        s->has_job = false;
        lk.unlock();
        const auto t0 = std::chrono::steady_clock::now();
        int rc = s->job(*s);
        const auto t1 = std::chrono::steady_clock::now();
        lk.lock();
        s->rc = rc;
        s->last_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (rc) s->err = g_err;  // g_err is thread_local: carry the message to the calling thread
        s->done = true;
        s->cv.notify_all();
    }
}
// runs job on every shard's own host thread and waits for all of them; returns the first error
int multi_run(dwbc_multi *mh, const std::function<int(dwbc_multi::Shard &)> &job) {
    for (auto &sp : mh->shards) {
        std::lock_guard<std::mutex> lk(sp->mu);
        sp->job = job;
        sp->done = false;
        sp->has_job = true;
        sp->cv.notify_all();
    }
    int rc = 0;
    for (auto &sp : mh->shards) {
        std::unique_lock<std::mutex> lk(sp->mu);
        sp->cv.wait(lk, [&] { return sp->done; });
        if (sp->rc && !rc) {
            rc = sp->rc;
            g_err = "gpu " + std::to_string(sp->device) + ": " + sp->err;
        }
    }
    return rc;
}
template <class F>
int multi_each(dwbc_multi *mh, F f) {  // configuration calls: cheap, run on the calling thread
    for (auto &sp : mh->shards) {
        const int rc = f(*sp);
        if (rc < 0) return rc;
    }
    return DWBC_OK;
}
}  // namespace

extern "C" {

dwbc_multi_t *dwbc_multi_create(const dwbc_model_t *m, int batch, int ngpus, const int *devices) {
    if (!m || batch <= 0 || ngpus <= 0 || ngpus > batch) {
        g_err = "dwbc_multi_create: bad arguments";
        return nullptr;
    }
    dwbc_multi *mh = new (std::nothrow) dwbc_multi();
    if (!mh) return nullptr;
    mh->B = batch;
    mh->n = m->hm.dev.n;
    mh->m = m->hm.dev.m;
    for (int g = 0; g < ngpus; g++) {
        // balanced contiguous slices: every GPU gets floor or ceil of batch / ngpus states (none is left empty while ngpus <= batch)
        const size_t first = (size_t)batch * g / ngpus, count = (size_t)batch * (g + 1) / ngpus - first;
        if (count == 0) continue;
        auto sp = std::make_unique<dwbc_multi::Shard>();
        sp->device = devices ? devices[g] : g;
        sp->first = first;
        sp->count = count;
        sp->b = dwbc_batch_create(m, (int)count, sp->device);
        if (!sp->b || !cuda_ok(cudaStreamCreateWithFlags(&sp->stream, cudaStreamNonBlocking), "cudaStreamCreateWithFlags")) {
            if (sp->b) dwbc_batch_destroy(sp->b);
            for (auto &o : mh->shards) {
                dwbc_batch_destroy(o->b);
            }
            delete mh;
            return nullptr;
        }
        mh->shards.push_back(std::move(sp));
    }
    for (auto &sp : mh->shards) sp->th = std::thread(multi_worker, sp.get());
    return mh;
}
void dwbc_multi_destroy(dwbc_multi_t *mh) {
    if (!mh) return;
    for (auto &sp : mh->shards) {
        {
            std::lock_guard<std::mutex> lk(sp->mu);
            sp->quit = true;
            sp->cv.notify_all();
        }
        if (sp->th.joinable()) sp->th.join();
        cudaSetDevice(sp->device);
        for (void *p : sp->dev_allocs) cudaFree(p);
        if (sp->stream) cudaStreamDestroy(sp->stream);
        dwbc_batch_destroy(sp->b);
    }
    delete mh;
}
int dwbc_multi_gpus(const dwbc_multi_t *mh) { return mh ? (int)mh->shards.size() : 0; }
int dwbc_multi_shard(dwbc_multi_t *mh, int g, dwbc_batch_t **batch, int *device, int *first_state, int *count) {
    if (!mh || g < 0 || g >= (int)mh->shards.size()) return fail(DWBC_ERR_ARG, "multi_shard: bad index");
    auto &s = *mh->shards[g];
    if (batch) *batch = s.b;
    if (device) *device = s.device;
    if (first_state) *first_state = (int)s.first;
    if (count) *count = (int)s.count;
    return DWBC_OK;
}
int dwbc_multi_add_contact(dwbc_multi_t *mh, int link, int contact_type, const double point[3], const double normal[3], double lx, double ly) {
    if (!mh) return fail(DWBC_ERR_ARG, "multi_add_contact: null handle");
    int idx = 0;
    const int rc = multi_each(mh, [&](dwbc_multi::Shard &s) { return idx = dwbc_batch_add_contact(s.b, link, contact_type, point, normal, lx, ly); });
    if (rc < 0) return rc;
    mh->ncontact = idx + 1;
    return idx;
}
int dwbc_multi_add_task(dwbc_multi_t *mh, int level, int task_mode, int link, const double point[3]) {
    if (!mh) return fail(DWBC_ERR_ARG, "multi_add_task: null handle");
    const int rc = multi_each(mh, [&](dwbc_multi::Shard &s) { return dwbc_batch_add_task(s.b, level, task_mode, link, point); });
    if (rc < 0) return rc;
    const BatchCfg &c = mh->shards[0]->b->cfg;
    mh->nlevels = c.nlevels;
    for (int l = 0; l < c.nlevels; l++) mh->tdof[l] = c.tasks[l].dof;
    return DWBC_OK;
}
int dwbc_multi_add_custom_task(dwbc_multi_t *mh, int level, int task_dof, const double *jtask_host) {
    if (!mh || !jtask_host) return fail(DWBC_ERR_ARG, "multi_add_custom_task: null argument");
    const int rc = multi_each(mh, [&](dwbc_multi::Shard &s) -> int {
        int r = dwbc_batch_add_custom_task(s.b, level, task_dof);
        if (r) return r;
        CK(cudaSetDevice(s.device));
        void *raw = nullptr;
        const size_t bytes = sizeof(double) * (size_t)task_dof * s.b->w.n;
        CK(cudaMalloc(&raw, bytes));
        s.dev_allocs.push_back(raw);
        CK(cudaMemcpy(raw, jtask_host, bytes, cudaMemcpyHostToDevice));
        s.b->w.jcustom[level] = (const double *)raw;  // shared by all states of the shard
        s.b->w.jcustom_per_state[level] = 0;
        return DWBC_OK;
    });
    if (rc < 0) return rc;
    const BatchCfg &c = mh->shards[0]->b->cfg;
    mh->nlevels = c.nlevels;
    for (int l = 0; l < c.nlevels; l++) mh->tdof[l] = c.tasks[l].dof;
    return DWBC_OK;
}
int dwbc_multi_set_torque_limit(dwbc_multi_t *mh, const double *tlim_host) {
    if (!mh) return fail(DWBC_ERR_ARG, "multi_set_torque_limit: null handle");
    return multi_each(mh, [&](dwbc_multi::Shard &s) { return dwbc_batch_set_torque_limit(s.b, tlim_host); });
}
int dwbc_multi_set_outputs(dwbc_multi_t *mh, int outputs) {
    if (!mh) return fail(DWBC_ERR_ARG, "multi_set_outputs: null handle");
    return multi_each(mh, [&](dwbc_multi::Shard &s) { return dwbc_batch_set_outputs(s.b, outputs); });
}

int dwbc_multi_cycle_host(dwbc_multi_t *mh, int stages, const double *q_host, const double *qdot_host, const double *fstar_host,
                          const uint8_t *flags_host, double *tau_grav_host, double *tau_task_host, double *tau_contact_host,
                          int32_t *status_host) {
    if (!mh || !q_host || !qdot_host || !flags_host) return fail(DWBC_ERR_ARG, "multi_cycle_host: null argument");
    const size_t n = mh->n, m = mh->m, nc = mh->ncontact, B = mh->B;
    return multi_run(mh, [=](dwbc_multi::Shard &s) -> int {
        const double *lv[MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr};
        if (fstar_host) {
            size_t off = 0;
            for (int l = 0; l < mh->nlevels; l++) {
                lv[l] = fstar_host + off + s.first * mh->tdof[l];
                off += (size_t)mh->tdof[l] * B;
            }
        }
        auto adv = [&](double *p, size_t rec) { return p ? p + s.first * rec : nullptr; };
        return dwbc_batch_cycle_host_levels(s.b, stages, q_host + s.first * (n + 1), qdot_host + s.first * n, fstar_host ? lv : nullptr,
                                            flags_host + s.first * nc, adv(tau_grav_host, m), adv(tau_task_host, m), adv(tau_contact_host, m),
                                            status_host ? status_host + s.first : nullptr, s.stream);
    });
}
// host wall time (ms) each shard's thread spent in the last dwbc_multi_cycle_host: out[g], g < dwbc_multi_gpus
int dwbc_multi_last_shard_ms(const dwbc_multi_t *mh, double *out) {
    if (!mh || !out) return fail(DWBC_ERR_ARG, "multi_last_shard_ms: null argument");
    for (size_t g = 0; g < mh->shards.size(); g++) out[g] = mh->shards[g]->last_ms;
    return DWBC_OK;
}
// pinned host memory visible to every GPU (the copies of dwbc_multi_cycle_host overlap with kernels only from pinned memory)
void *dwbc_host_alloc(size_t bytes) {
    void *p = nullptr;
#ifdef DWBC_CPU_EMUL
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
#else
    if (!cuda_ok(cudaHostAlloc(&p, bytes, cudaHostAllocPortable), "cudaHostAlloc")) return nullptr;
#endif
    return p;
}
void dwbc_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ host-side staging helpers
// For hosts that do not link the CUDA runtime themselves (the RobotData shim include/dwbc_compat/dwbc.h, the pybind module): device
// buffers and copies through the C ABI.  Plumbing only.
extern "C" {
void *dwbc_device_alloc(int device, size_t bytes) {
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
    void *p = nullptr;
    if (!cuda_ok(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc")) return nullptr;
    cudaMemset(p, 0, bytes ? bytes : 1);
    return p;
}
void dwbc_device_free(int device, void *p) {
    if (!p) return;
    cudaSetDevice(device);
    cudaFree(p);
}
int dwbc_memcpy_h2d(int device, void *dst_dev, const void *src_host, size_t bytes, void *stream) {
    CK(cudaSetDevice(device));
    CK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return DWBC_OK;
}
int dwbc_memcpy_d2h(int device, void *dst_host, const void *src_dev, size_t bytes, void *stream) {
    CK(cudaSetDevice(device));
    CK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    return DWBC_OK;
}
int dwbc_device_count(void) {
#ifdef DWBC_CPU_EMUL
    return 1;
#else
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
#endif
}
}  // extern "C"
