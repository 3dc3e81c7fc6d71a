/*
 * dwbc_b200 -- C ABI of the batched whole-body-control hot path (B200 / sm_100a).
 *
 * This header is the drop-in boundary.  Each entry point names the reference interface it replaces
 * (paths relative to the saga0619/libdwbc checkout).  One `dwbc_batch_t` plays the role of B
 * independent `DWBC::RobotData` objects (reference include/dwbc.h:59-430) that share a model, a contact
 * list, a task hierarchy and a torque limit, and differ only in their per-cycle state.
 *
 * Conventions
 *   - Everything is FP64.  Per-state vectors are stored state-major: x[b*len + i].
 *   - Per-state matrices are column-major (Eigen::MatrixXd layout, reference tests/dwbc_test_util.h:23-26)
 *     and state-major: M[b*rows*cols + i + j*rows].
 *   - q has n+1 entries: [p(3) | quat xyz(3) | joints(m) | quat w]; qdot has n entries
 *     [v_world(3) | omega_body(3) | joint rates(m)]   (reference src/pybind/rl_wbc_bridge.cpp:88-98).
 *   - Pointers named *_dev must be device pointers on the batch's GPU, valid until the stream work that
 *     uses them has completed; the library never copies host<->device behind the caller's back except
 *     in the explicitly named *_host entry points.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls on one handle are
 *     stream-ordered; different handles are independent.  A handle is not thread-safe (like RobotData).
 *   - Return value: 0 on success, a negative DWBC_ERR_* code otherwise; dwbc_last_error() gives text.
 *     Per-state outcomes (the reference's 1/0 ints) are reported in status[b], see DWBC_ST_*.
 *   - There is no CPU fallback: every compute entry point launches CUDA kernels.
 */
#ifndef DWBC_B200_H
#define DWBC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dwbc_model dwbc_model_t;
typedef struct dwbc_batch dwbc_batch_t;

enum {
    DWBC_OK = 0,
    DWBC_ERR_ARG = -1,      /* bad argument (null pointer, size, unknown link, level out of order ...) */
    DWBC_ERR_MODEL = -2,    /* URDF could not be read / unsupported model */
    DWBC_ERR_CUDA = -3,     /* CUDA runtime error (no device, out of memory, launch failure) */
    DWBC_ERR_ORDER = -4,    /* call-order contract violated (e.g. solve before update_kinematics) */
    DWBC_ERR_LIMIT = -5     /* exceeds a compiled-in limit (contacts, levels, task dof) */
};

/* reference include/dwbc_contact_constraint.h:19-25 */
enum { DWBC_CONTACT_6D = 0, DWBC_CONTACT_POINT = 1, DWBC_CONTACT_LINK = 2, DWBC_CONTACT_LINE = 3 };

/* reference include/dwbc_task.h:23-33 (LINK_MODE) */
enum {
    DWBC_TASK_LINK_6D = 0,
    DWBC_TASK_LINK_6D_COM_FRAME = 1,
    DWBC_TASK_LINK_6D_CUSTOM_FRAME = 2,
    DWBC_TASK_LINK_POSITION = 3,
    DWBC_TASK_LINK_POSITION_COM_FRAME = 4,
    DWBC_TASK_LINK_POSITION_CUSTOM_FRAME = 5,
    DWBC_TASK_LINK_ROTATION = 6,
    DWBC_TASK_LINK_ROTATION_CUSTOM_FRAME = 7
};

/* stage selection for dwbc_batch_solve */
enum {
    DWBC_SOLVE_GRAV = 1,      /* CalcGravCompensation            reference src/dwbc.cpp:875-889 */
    DWBC_SOLVE_TASK = 2,      /* CalcTaskSpace + CalcTaskControlTorque(hqp=false)  :795-873 */
    DWBC_SOLVE_HQP = 4,       /* ... with hqp=true: per-level QP  :941-1127 (implies DWBC_SOLVE_TASK) */
    DWBC_SOLVE_REDIST = 8,    /* CalcContactRedistribute(hqp=true) :1372-1568 */
    DWBC_SOLVE_REDIST_ANALYTIC = 16 /* CalcContactRedistribute(hqp=false) :1570-1619 + ContactRedistributetwomod src/wbd.cpp:273-404
                                       (two-foot closed form; ignored when DWBC_SOLVE_REDIST is also set) */
};

/* per-state status bits (1 = the corresponding reference call would have returned 1) */
enum {
    DWBC_ST_CONTACT_OK = 1,   /* CalcContactConstraint */
    DWBC_ST_TASK_OK = 2,      /* CalcTaskControlTorque */
    DWBC_ST_REDIST_OK = 4,    /* CalcContactRedistribute */
    DWBC_ST_QP_SLOW = 16      /* diagnostic: a QP's proximal-point loop stopped at its iteration cap (the feasible iterate was kept) */
};

/* matrices retrievable with dwbc_batch_get_matrix (public RobotData fields of the reference) */
enum {
    DWBC_MAT_A = 0,           /* A_            n x n */
    DWBC_MAT_A_INV,           /* A_inv_        n x n */
    DWBC_MAT_G,               /* G_            n x 1 */
    DWBC_MAT_B,               /* B_            n x 1 */
    DWBC_MAT_J_C,             /* J_C           cmax x n   (rows >= contact_dof are zero) */
    DWBC_MAT_LAMBDA_C,        /* Lambda_contact cmax x cmax */
    DWBC_MAT_J_C_INV_T,       /* J_C_INV_T     cmax x n */
    DWBC_MAT_N_C,             /* N_C           n x n      (materialised on request) */
    DWBC_MAT_A_INV_N_C,       /* A_inv_N_C     n x n */
    DWBC_MAT_W,               /* W             m x m */
    DWBC_MAT_W_INV,           /* W_inv         m x m */
    DWBC_MAT_V2,              /* V2            rmax x m   (orthonormal rows; basis is not unique) */
    DWBC_MAT_NWJW,            /* NwJw          m x rmax */
    DWBC_MAT_P_C,             /* P_C           cmax x 1 */
    DWBC_MAT_TORQUE_GRAV,     /* torque_grav_  m x 1 */
    DWBC_MAT_TORQUE_TASK,     /* torque_task_  m x 1 */
    DWBC_MAT_TORQUE_CONTACT,  /* torque_contact_ m x 1 */
    DWBC_MAT_COM,             /* com_pos       3 x 1 */
    DWBC_MAT_LINK_POS,        /* link_[i].xpos 3 x nb */
    DWBC_MAT_LINK_ROT,        /* link_[i].rotm 9 x nb (each column a column-major 3x3) */
    DWBC_MAT_LINK_VEL,        /* link_[i].v, .w 6 x nb */
    DWBC_MAT_CF_REDIS,        /* cf_redis_qp_  rmax x 1 */
    /* per hierarchy level (argument `level`) */
    DWBC_MAT_J_TASK = 64,     /* ts_[l].J_task_      t x n */
    DWBC_MAT_LAMBDA_TASK,     /* ts_[l].Lambda_task_ t x t */
    DWBC_MAT_J_KT,            /* ts_[l].J_kt_        m x t */
    DWBC_MAT_NULL_TASK,       /* ts_[l].Null_task_   m x m */
    DWBC_MAT_FSTAR_QP,        /* ts_[l].f_star_qp_   t x 1 */
    DWBC_MAT_CONTACT_QP,      /* ts_[l].contact_qp_  rmax x 1 */
    DWBC_MAT_TORQUE_H         /* ts_[l].torque_h_    m x 1 */
};

const char *dwbc_last_error(void);
/* number of CUDA kernels launched by this library since process start (bench.py's gpu_launches) */
long long dwbc_launch_count(void);

/* ---- model: RobotData::LoadModelData(urdf_path, floating)  reference src/dwbc.cpp:102-123 ---- */
dwbc_model_t *dwbc_model_load_urdf(const char *urdf_path, int floating);
void dwbc_model_destroy(dwbc_model_t *m);
/* system_dof_, model_dof_, number of links (without the COM pseudo link), total_mass_ */
int dwbc_model_info(const dwbc_model_t *m, int *system_dof, int *model_dof, int *n_links, double *total_mass);
/* case-insensitive link lookup (reference src/dwbc.cpp:393-404); "COM" gives n_links; -1 if unknown */
int dwbc_model_link_id(const dwbc_model_t *m, const char *name);
const char *dwbc_model_link_name(const dwbc_model_t *m, int link);

/* ---- batch handle ---- */
dwbc_batch_t *dwbc_batch_create(const dwbc_model_t *m, int batch, int device);
void dwbc_batch_destroy(dwbc_batch_t *b);
int dwbc_batch_size(const dwbc_batch_t *b);
/* compile-time capacities: cmax = 6*max active contacts, rmax = cmax-6 */
int dwbc_batch_limits(const dwbc_batch_t *b, int *cmax, int *rmax, int *max_levels, int *max_task_dof);

/* ---- setup (once): same meaning as the RobotData calls they replace ---- */
/* AddContactConstraint(link, type, point, normal, lx, ly)   reference src/dwbc.cpp:377-391.
 * Returns the contact index (>= 0). Only DWBC_CONTACT_6D is supported on this path. */
int dwbc_batch_add_contact(dwbc_batch_t *b, int link, int contact_type, const double point[3], const double normal[3], double lx, double ly);
/* ContactConstraint::SetFrictionRatio   reference src/contact_constraint.cpp:93-97 (default 0.2, 0.2) */
int dwbc_batch_set_friction(dwbc_batch_t *b, int contact, double friction_ratio, double friction_ratio_z);
/* AddTaskSpace(heirarchy, task_mode, link, point): a new level if heirarchy == number of levels, otherwise
 * the link is appended to that level (AddTaskLink)   reference src/dwbc.cpp:531-600 */
int dwbc_batch_add_task(dwbc_batch_t *b, int level, int task_mode, int link, const double point[3]);
/* AddTaskSpace(heirarchy, TASK_CUSTOM, task_dof)   reference src/dwbc.cpp:522-530 */
int dwbc_batch_add_custom_task(dwbc_batch_t *b, int level, int task_dof);
int dwbc_batch_clear_tasks(dwbc_batch_t *b);
/* SetTorqueLimit(tlim)   reference src/dwbc.cpp:254-258; tlim is a HOST array of model_dof entries */
int dwbc_batch_set_torque_limit(dwbc_batch_t *b, const double *tlim_host);

/* ---- per control cycle, device buffers, in the reference's call order ---- */
/* UpdateKinematics(q, qdot, qddot)   reference src/dwbc.cpp:279-371.  q_dev [B x (n+1)], qdot_dev [B x n]. */
int dwbc_batch_update_kinematics(dwbc_batch_t *b, const double *q_dev, const double *qdot_dev, void *stream);
/* SetContact(flags...)   reference include/dwbc.h:432-474. flags_dev [B x ncontact], 0/1 per registered contact. */
int dwbc_batch_set_contact(dwbc_batch_t *b, const uint8_t *flags_dev, void *stream);
/* CalcContactConstraint()   reference src/dwbc.cpp:456-478 */
int dwbc_batch_calc_contact_constraint(dwbc_batch_t *b, void *stream);
/* SetTaskSpace(heirarchy, f_star, J_task)   reference src/dwbc.cpp:662-678.  fstar_dev [B x t].
 * jtask_dev: NULL for link tasks; for a custom task either [t x n] shared by all states (jtask_per_state = 0)
 * or [B x t x n] (jtask_per_state = 1), column-major. */
int dwbc_batch_set_task(dwbc_batch_t *b, int level, const double *fstar_dev, const double *jtask_dev, int jtask_per_state, void *stream);
/* CalcGravCompensation / CalcTaskControlTorque / CalcContactRedistribute as selected by `stages`.
 * Any output pointer may be NULL.  torques are [B x m]; status [B] (DWBC_ST_* bits). */
int dwbc_batch_solve(dwbc_batch_t *b, int stages, double *tau_grav_dev, double *tau_task_dev, double *tau_contact_dev,
                     int32_t *status_dev, void *stream);
/* Copies a per-state matrix into out_dev ([B x rows x cols], column-major per state). rows/cols may be NULL. */
int dwbc_batch_get_matrix(dwbc_batch_t *b, int which, int level, double *out_dev, int *rows, int *cols, void *stream);
/* Active rows of the level's task QP (level >= 0) or of the redistribution QP (level = -1) as a bit mask
 * per state: out_dev [B x 4] uint32 words, bit i of word i/32 = row i (rows ordered as in the reference:
 * +torque (m), -torque (m), contact (10 per active contact)). */
int dwbc_batch_get_active_set(dwbc_batch_t *b, int level, uint32_t *out_dev, void *stream);

/* getContactForce(command_torque)   reference src/dwbc.cpp:891-896, src/wbd.cpp:268-271:
 * cf = J_C_INV_T[:, 6:] tau - P_C.  torque_dev [B x m], or NULL for torque_grav_ + torque_task_ + torque_contact_ of the last
 * solve.  cf_dev [B x cmax], rows >= contact_dof are zero. */
int dwbc_batch_get_contact_force(dwbc_batch_t *b, const double *torque_dev, double *cf_dev, void *stream);
/* getZMP(contact_force)   reference src/dwbc.cpp:898-939.  cf_dev [B x cmax] (stacked active contacts), zmp_dev [B x 3]. */
int dwbc_batch_get_zmp(dwbc_batch_t *b, const double *cf_dev, double *zmp_dev, void *stream);

/* ---- a9: reduced-dimension variant (Ahn 2024), reference src/dwbc.cpp:2752-3446, in the call order of
 * reference tests/sp_test/redu_dyn_test.cpp:262-267: update_kinematics -> set_contact -> reduced_dynamics ->
 * reduced_calc_contact_constraint -> set_task -> reduced_solve.
 * The non-contact chain is condensed to a 6-dof body: n 39 -> 24 and m 33 -> 18 in double support.  As in the reference
 * the q order [base | contact chain | non-contact chain] must be contiguous (double support or left-foot single support),
 * and one call handles ONE contact pattern: `pattern_host` [ncontact] gives the flags every state is expected to show;
 * states with other flags get status 0.  Levels must be ordered COM / contact-chain tasks first, then non-contact tasks;
 * a level mixing both is the reference's "UNDEFINED TASK TYPE" and is rejected (DWBC_ERR_ARG). */
/* ReducedDynamicsCalculate()   reference src/dwbc.cpp:2752-2989 */
int dwbc_batch_reduced_dynamics(dwbc_batch_t *b, const uint8_t *pattern_host, void *stream);
/* ReducedCalcContactConstraint()   reference src/dwbc.cpp:3077-3142 */
int dwbc_batch_reduced_calc_contact_constraint(dwbc_batch_t *b, void *stream);
/* ReducedCalcGravCompensation (DWBC_SOLVE_GRAV) + ReducedCalcTaskSpace + ReducedCalcTaskControlTorque (DWBC_SOLVE_TASK: hqp=false;
 * DWBC_SOLVE_HQP: hqp=true, CalcSingleTaskTorqueWithQP_R) + ReducedCalcContactRedistribute(hqp=true) (DWBC_SOLVE_REDIST,
 * CalcContactRedistributeR)   reference src/dwbc.cpp:3144-3599, 3758-3770, 4777-4925.  Outputs as dwbc_batch_solve (full
 * 33-joint torques; torque_contact_[0:contact_dof_] = torque_contact_R_[0:contact_dof_]).
 * DWBC_SOLVE_HQP / DWBC_SOLVE_REDIST cover every hierarchy the closed form covers (non-contact levels go through
 * CalcSingleTaskTorqueWithQP_R_NC, src/dwbc.cpp:3601-3756); DWBC_SOLVE_REDIST_ANALYTIC is the analytic branch of
 * CalcContactRedistributeR (:4938-4990). */
int dwbc_batch_reduced_solve(dwbc_batch_t *b, int stages, double *tau_grav_dev, double *tau_task_dev, double *tau_contact_dev,
                             int32_t *status_dev, void *stream);
/* Torque limit of the reduced QPs: reduced_model_dof_ entries (contact-chain joints, then the 6 coordinates of the condensed
 * non-contact body), as CalcSingleTaskTorqueWithQP_R compares torque_limit_.size() with reduced_model_dof_ (reference
 * src/dwbc.cpp:3462-3468).  The full-model limit of dwbc_batch_set_torque_limit is not used by the reduced QPs.  count = 0 clears it. */
int dwbc_batch_set_reduced_torque_limit(dwbc_batch_t *b, const double *tlim_host, int count);
/* co_dof, nc_dof, vc_dof, reduced_system_dof_, reduced_model_dof_ of the current pattern (reference src/dwbc.cpp:2818-2823) */
int dwbc_batch_reduced_dims(const dwbc_batch_t *b, int *co_dof, int *nc_dof, int *vc_dof, int *reduced_system_dof, int *reduced_model_dof);
/* public RobotData fields of the reduced variant */
enum {
    DWBC_RMAT_A_R_INV = 0,    /* A_R_inv        rn x rn */
    DWBC_RMAT_A_R,            /* A_R            rn x rn */
    DWBC_RMAT_J_I_NC,         /* J_I_nc_        6 x nc */
    DWBC_RMAT_J_I_NC_INV_T,   /* J_I_nc_inv_T   6 x nc */
    DWBC_RMAT_G_R,            /* G_R            rn x 1 */
    DWBC_RMAT_J_CR,           /* J_CR           cmax x rn */
    DWBC_RMAT_LAMBDA_CR,      /* Lambda_CR      cmax x cmax */
    DWBC_RMAT_J_CR_INV_T,     /* J_CR_INV_T     cmax x rn */
    DWBC_RMAT_A_R_INV_N_CR,   /* A_R_inv_N_CR   rn x rn */
    DWBC_RMAT_W_R,            /* W_R            rm x rm */
    DWBC_RMAT_W_R_INV,        /* W_R_inv        rm x rm */
    DWBC_RMAT_NWJW_R,         /* NwJw_R         rm x rmax */
    DWBC_RMAT_P_CR,           /* P_CR           cmax x 1 */
    DWBC_RMAT_TORQUE_GRAV_R,  /* torque_grav_R_ rm x 1 */
    DWBC_RMAT_TORQUE_TASK_R,  /* torque_task_R_ rm x 1 */
    DWBC_RMAT_J_BASE_R_KT,    /* J_base_R_kt_   rm x 6 */
    DWBC_RMAT_SI_NC,          /* [mass_nc_, com_pos_nc_ (3), inertia_nc_ (9, row-major)]  13 x 1 */
    /* per hierarchy level */
    DWBC_RMAT_J_TASK_R = 64,  /* ts_[l].J_task_R_      t x rn  (COM / contact-chain levels) */
    DWBC_RMAT_LAMBDA_TASK,    /* ts_[l].Lambda_task_   t x t */
    DWBC_RMAT_J_KT_R,         /* ts_[l].J_kt_R_        rm x t  (COM / contact-chain levels) */
    DWBC_RMAT_NULL_TASK_R     /* ts_[l].Null_task_R_   rm x rm (COM / contact-chain levels) */
};
int dwbc_batch_get_reduced_matrix(dwbc_batch_t *b, int which, int level, double *out_dev, int *rows, int *cols, void *stream);

/* ---- a10: LQP / centroidal simplified-WBD variant (Ahn 2025): ConfigureLQP + CalcControlTorqueLQP in one call,
 * reference src/dwbc.cpp:4304-4454, src/dwbc_hqp.cpp:23-85, 291-403.  Call order as reference
 * tests/sp_test/jacc_compare.cpp:367-384: update_kinematics -> set_contact -> calc_contact_constraint -> set_task (every
 * level) -> calc_control_torque_lqp.  Decision vector y = [qdd (n); Fc (contact_dof)].  Levels: 0 floating-base dynamics
 * (preset answer), 1 contact rows, 2.. one per task level.  Every level is solved exactly (the reference uses OSQP with
 * eps 1e-3; parity unpinned).
 * y_dev [B x (n + cmax)] final answer (entries >= n + contact_dof are zero), tau_dev [B x m] joint torque
 * A[6:,:] qdd + Jc^T[6:,:] Fc + B[6:], status_dev [B]: number of levels (= 2 + task levels) when every level was solved,
 * -(levels solved) when a level failed (the answer of the last solved level is returned, as HQP::solveSequential leaves it),
 * 0 when the contact constraint of the state is invalid.  Any output pointer may be NULL. */
int dwbc_batch_calc_control_torque_lqp(dwbc_batch_t *b, double *y_dev, double *tau_dev, int32_t *status_dev, void *stream);
/* ConfigureLQP_R + CalcControlTorqueLQP_R (reference src/dwbc.cpp:4504-4632, 4456-4478): the same hierarchy on the reduced model
 * (A_R, A_R_inv, G_R, J_CR, J_task J_R_INV_T^T; torque bound 200 with 600 on row m-4; cost A_R / |A_|), task levels = the COM /
 * contact-chain levels.  Call after reduced_dynamics, reduced_calc_contact_constraint and reduced_solve(DWBC_SOLVE_TASK) (which
 * classifies the levels and forms the reduced task Jacobians).  y_dev [B x (reduced_system_dof + cmax)], tau_dev
 * [B x reduced_model_dof]. */
int dwbc_batch_calc_control_torque_lqp_r(dwbc_batch_t *b, double *y_dev, double *tau_dev, int32_t *status_dev, void *stream);
/* ConfigureLQP_R_NC + CalcControlTorqueLQP_R_NC (reference src/dwbc.cpp:4634-4760, 4479-4502): the non-contact chain's own two-level
 * hierarchy over its nc_dof joint accelerations, fed by the answer of dwbc_batch_calc_control_torque_lqp_r (call that first):
 * level 0 centroidal equality J_I_nc qdd = fstar_gnc + torque rows |A[nc,nc] qdd + G[nc]| <= 200; level 1 the 6D task of hierarchy
 * level 1 (the reference hard-codes ts_[1]) in the pelvis-relative form + |qdd| <= 5.  qdd_nc_dev [B x nc_dof] final answer;
 * tau_dev [B x m] the stitched joint torque of the only caller (reference tests/sp_test/jacc_compare.cpp:543-569): contact-chain joints
 * from the reduced LQP, the others A[nc,nc] qdd + G[nc]; status_dev [B]: 2 both levels solved, -k level k failed, 0 no input. */
int dwbc_batch_calc_control_torque_lqp_r_nc(dwbc_batch_t *b, double *qdd_nc_dev, double *tau_dev, int32_t *status_dev, void *stream);
enum { DWBC_LQP_REDUCED = 256 }; /* OR into `which` of dwbc_batch_get_lqp for the results of the reduced variant */
enum {
    DWBC_LQP_Y = 0,      /* hqp_hs_[level].y_ans_   (n + cmax) x 1, level in [0, 2 + task levels) */
    DWBC_LQP_V1,         /* hqp_hs_[1].v_ans_       (2m + 10*max_active_contacts) x 1: friction/ZMP rows, then +accel, -accel */
    DWBC_LQP_TORQUE      /* m x 1 */
};
int dwbc_batch_get_lqp(dwbc_batch_t *b, int which, int level, double *out_dev, int *rows, int *cols, void *stream);
/* active-set iterations of the last LQP call per state: out_dev [B] int32 */
int dwbc_batch_get_lqp_iterations(dwbc_batch_t *b, int32_t *out_dev, void *stream);

/* ---- f* generation (task PD on quintic / slerp trajectories), the step before the path.
 * TaskLink::SetTaskGain / SetTrajectoryQuintic / SetTrajectoryRotation / GetFstarPosPD / GetFstarRotPD
 * (reference src/task.cpp:223-334), QuinticSpline and GetPhi (src/math.cpp:127-183, 272-287), per LINK_MODE as in
 * UpdateTaskSpace (src/dwbc.cpp:714-785).  `link_index` counts the task links of the level in AddTaskSpace order.
 * gains18_host = [pos_p, pos_d, pos_a, rot_p, rot_d, rot_a] (3 each), shared by all states.  Trajectory parameters are
 * per state and stay caller-owned device buffers: time_dev [B x 2] = (start, end), positions / velocities [B x 3],
 * rotations [B x 9] row-major.  (The reference's twist_init / twist_desired are stored but never used by GetFstarRotPD.) */
int dwbc_batch_set_task_gain(dwbc_batch_t *b, int level, int link_index, const double *gains18_host);
int dwbc_batch_set_trajectory_quintic(dwbc_batch_t *b, int level, int link_index, const double *time_dev, const double *pos_init_dev,
                                      const double *vel_init_dev, const double *pos_desired_dev, const double *vel_desired_dev);
int dwbc_batch_set_trajectory_rotation(dwbc_batch_t *b, int level, int link_index, const double *time_dev, const double *rot_init_dev,
                                       const double *rot_desired_dev);
/* f* of `level` at control_time_dev [B] from the link states of the last update_kinematics (whose qdot buffer must still be
 * valid): writes the rows of every task link whose trajectory is set into fstar_dev [B x t] (other rows are left as they
 * are, like f_star_ in the reference); pass fstar_dev to dwbc_batch_set_task afterwards. */
int dwbc_batch_calc_fstar(dwbc_batch_t *b, int level, const double *control_time_dev, double *fstar_dev, void *stream);

/* ---- section 8(f) row 3: the RL caller's wire format (reference src/pybind/rl_wbc_bridge.cpp:88-107, 131-139).
 * update_kinematics from MuJoCo-ordered qpos [B x (n+1)] = [xyz | qw qx qy qz | joints] and qvel [B x n]. */
int dwbc_batch_update_kinematics_mujoco(dwbc_batch_t *b, const double *qpos_dev, const double *qvel_dev, void *stream);
/* getTorqueCommand(): float32 (torque_grav_ + torque_task_ + torque_contact_) of the last solve, out_dev [B x m]. */
int dwbc_batch_get_torque_command_f32(dwbc_batch_t *b, float *out_dev, void *stream);

/* ---- whole cycle with HOST buffers (pinned or pageable): H2D, all stages, D2H, synchronised on return.
 * This is what the RobotData shim, the pybind bridge and bench.py's end-to-end timing call
 * (one call = UpdateKinematics ... CalcContactRedistribute of reference tests/dwbc_test.cpp:61-130 for every state).
 * fstar_host: concatenation over levels of [B x t_l] blocks. flags_host [B x ncontact].
 * Large batches are cut into slices of consecutive states that alternate between two internal streams ordered after `stream`,
 * so the copies of one slice run under the kernels of another (pinned host buffers needed for that overlap); results are
 * identical to the unsliced call because states are independent. */
int dwbc_batch_cycle_host(dwbc_batch_t *b, int stages, const double *q_host, const double *qdot_host, const double *fstar_host,
                          const uint8_t *flags_host, double *tau_grav_host, double *tau_task_host, double *tau_contact_host,
                          int32_t *status_host, void *stream);
/* The same with one f* pointer per hierarchy level (fstar_levels_host[l] -> [B x t_l]): what a caller uses whose level blocks are not
 * one concatenation, e.g. a slice of a larger batch (dwbc_multi_cycle_host). */
int dwbc_batch_cycle_host_levels(dwbc_batch_t *b, int stages, const double *q_host, const double *qdot_host, const double *const *fstar_levels_host,
                                 const uint8_t *flags_host, double *tau_grav_host, double *tau_task_host, double *tau_contact_host,
                                 int32_t *status_host, void *stream);
/* Number of slices dwbc_batch_cycle_host cuts the batch into: 0 = automatic (one per 4096 states, at most 16), 1 = no slicing. */
int dwbc_batch_set_host_slices(dwbc_batch_t *b, int slices);

/* ---- one host batch over several GPUs of one box: contiguous slices of the batch, one host thread + one
 * stream + one dwbc_batch_t per GPU, no collective on the data path; every GPU reads and writes its own slice of the caller's host
 * buffers and the call returns when every slice is done.  `devices` = NULL means GPUs 0 .. ngpus-1.  The setup calls mirror the
 * dwbc_batch_* ones and are applied to every shard (a custom task takes its [t x n] column-major Jacobian, shared by all states, from
 * the host).  Host buffers should come from dwbc_host_alloc (pinned, visible to every GPU) for the copies to overlap the kernels. */
typedef struct dwbc_multi dwbc_multi_t;
dwbc_multi_t *dwbc_multi_create(const dwbc_model_t *m, int batch, int ngpus, const int *devices);
void dwbc_multi_destroy(dwbc_multi_t *mh);
int dwbc_multi_gpus(const dwbc_multi_t *mh);
/* shard g: its handle (for per-GPU queries), device, first state and state count */
int dwbc_multi_shard(dwbc_multi_t *mh, int g, dwbc_batch_t **batch, int *device, int *first_state, int *count);
int dwbc_multi_add_contact(dwbc_multi_t *mh, int link, int contact_type, const double point[3], const double normal[3], double lx, double ly);
int dwbc_multi_add_task(dwbc_multi_t *mh, int level, int task_mode, int link, const double point[3]);
int dwbc_multi_add_custom_task(dwbc_multi_t *mh, int level, int task_dof, const double *jtask_host);
int dwbc_multi_set_torque_limit(dwbc_multi_t *mh, const double *tlim_host);
int dwbc_multi_set_outputs(dwbc_multi_t *mh, int outputs);
/* dwbc_batch_cycle_host for the whole batch: same argument layout ([B x ...] host arrays, f* = concatenation over levels of [B x t_l]) */
int dwbc_multi_cycle_host(dwbc_multi_t *mh, int stages, const double *q_host, const double *qdot_host, const double *fstar_host,
                          const uint8_t *flags_host, double *tau_grav_host, double *tau_task_host, double *tau_contact_host,
                          int32_t *status_host);
/* host wall time (ms) of every shard's thread in the last dwbc_multi_cycle_host: out[g], g < dwbc_multi_gpus */
int dwbc_multi_last_shard_ms(const dwbc_multi_t *mh, double *out);
void *dwbc_host_alloc(size_t bytes);
void dwbc_host_free(void *p);

/* ---- output groups: which RobotData fields a cycle materialises in HBM ----
 * DWBC_OUT_ALL      every public field of the reference's RobotData (A_, A_inv_, J_C, Lambda_contact, J_C_INV_T, A_inv_N_C, W, W_inv,
 *                   V2, NwJw, per level J_task_, Lambda_task_, J_kt_, Null_task_, torque_h_ ...) is written for every state and can be
 *                   read with dwbc_batch_get_matrix: ~110 KB of workspace and ~190 KB of HBM traffic per state and cycle.
 * DWBC_OUT_TORQUES  torque_grav_ / torque_task_ / torque_contact_, status, f_star_qp_ / contact_qp_ / cf_redis_qp_ and the active sets
 *                   only.  The staged calls (update_kinematics, set_contact, calc_contact_constraint, set_task) then only RECORD
 *                   their arguments (the buffers must stay valid until dwbc_batch_solve returns) and dwbc_batch_solve runs the whole
 *                   cycle of a state in one kernel with every intermediate matrix in shared memory, followed by the active-set QP
 *                   kernel over the states whose QPs are active.  ~19 KB of workspace per state.  dwbc_batch_get_matrix of any other
 *                   field returns DWBC_ERR_ORDER.  Models / hierarchies without a compiled specialisation of that kernel fall back to
 *                   the generic kernels transparently (dwbc_batch_last_path tells which ran).
 * DWBC_OUT_AUTO     (default) DWBC_OUT_ALL for the staged calls, DWBC_OUT_TORQUES inside dwbc_batch_cycle_host. */
enum { DWBC_OUT_AUTO = -1, DWBC_OUT_TORQUES = 0, DWBC_OUT_ALL = 1 };
int dwbc_batch_set_outputs(dwbc_batch_t *b, int outputs);
/* 1 if the last dwbc_batch_solve / dwbc_batch_cycle_host ran the one-kernel cycle, 0 if it ran the generic kernels */
int dwbc_batch_last_path(const dwbc_batch_t *b);
/* device memory (bytes) the handle has allocated so far: inputs staging, outputs and the per-state workspace of the output group in use
 * (allocation is lazy: the public-field workspace of DWBC_OUT_ALL only exists once that group has been used) */
long long dwbc_batch_workspace_bytes(const dwbc_batch_t *b);
/* Diagnostics of the one-kernel cycle: stamps_dev [B x 32] int64 (device) receives the SM clock (clock64) at the end of every phase of
 * every state (phase list next to LEAN_STAMP in csrc/kernels_lean.cuh); NULL switches it off.  tools/lean_phases.py prints the table. */
int dwbc_batch_set_lean_timing(dwbc_batch_t *b, long long *stamps_dev);

/* ---- staging for hosts that do not link the CUDA runtime (include/dwbc_compat/dwbc.h, the pybind module `dwbc`): plumbing only ---- */
void *dwbc_device_alloc(int device, size_t bytes);            /* zero-initialised device buffer, NULL on failure */
void dwbc_device_free(int device, void *p);
int dwbc_memcpy_h2d(int device, void *dst_dev, const void *src_host, size_t bytes, void *stream);  /* synchronous on return */
int dwbc_memcpy_d2h(int device, void *dst_host, const void *src_dev, size_t bytes, void *stream);
int dwbc_device_count(void);

/* ---- diagnostics ---- */
/* Best DFMA-chain throughput (TFLOP/s, 2 flop per FMA) observed over ~`seconds` of back-to-back launches on
 * `device`: the FP64-pipe roofline denominator used by bench.py. */
int dwbc_probe_fp64(int device, double seconds, double *tflops);
/* Per-kernel device time: with profiling enabled every kernel launch of this handle is bracketed by CUDA events
 * recorded on the launch stream (no host synchronisation is added to the hot path).  dwbc_batch_get_kernel_times
 * synchronises the pending events and returns, per DWBC_K_* kernel, the summed duration in ms (only launches made while
 * profiling was on) and the launch count since the last reset.  bench.py's roofline numbers come from here. */
enum { DWBC_K_RBD = 0, DWBC_K_AINV, DWBC_K_CONTACT_JAC, DWBC_K_CONTACT_DYN, DWBC_K_TASK, DWBC_K_HQP, DWBC_K_EXPORT, DWBC_K_POST, DWBC_K_REDUCED, DWBC_K_LQP, DWBC_K_LEAN, DWBC_K_COUNT };
int dwbc_batch_set_profiling(dwbc_batch_t *b, int enable);
int dwbc_batch_get_kernel_times(dwbc_batch_t *b, double *ms_sum, long long *launches, int reset);
const char *dwbc_kernel_name(int k);
/* Total active-set iterations of the last dwbc_batch_solve per state: out_dev [B] int32. */
int dwbc_batch_get_qp_iterations(dwbc_batch_t *b, int32_t *out_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DWBC_B200_H */
