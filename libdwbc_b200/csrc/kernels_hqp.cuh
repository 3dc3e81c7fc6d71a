// Stages a6-a8: the prioritised torque QPs (one per hierarchy level) and the contact-force
// redistribution QP, solved by a batched dense FP64 active-set method -- one CTA (team) per robot state.
//
// Reference: RobotData::CalcTaskControlTorque hqp=true branch and CalcSingleTaskTorqueWithQP
// (src/dwbc.cpp:818-853, 941-1127), CalcContactRedistribute QP branch (src/dwbc.cpp:1377-1568),
// CQuadraticProgram::SolveQPoases (src/qp_wrapper.cpp:192-380; qpOASES itself is third-party).
//
// QP of level l (variables x = [x_t (t); x_c (r)], rows ordered as in the reference):
//     min 1/2 |x_t|^2   s.t.   [ U ; -U ; -Atemp U ] x <= [ tlim - base ; tlim + base ; -(bA0 - Atemp base) ]
//     U = [Ntorque_l | NwJw],  base = tau_prev + Ntorque_l f*,  Atemp = A_const A_rot Jbar^T[:,6:],  bA0 = A_const A_rot P_C
// The constraint matrix is never formed: only U (m x nv) and Atemp U (10k x nv) are kept.
//
// Semantics: qpOASES with setToMPC() regularises the semi-definite Hessian with eps ~ 2e-13, i.e. it returns the
// LEXICOGRAPHIC optimum (min |x_t|^2, then min |x_c|^2 among the minimisers).  That is computed
// here exactly, without an ill-conditioned epsilon:
//   0. all right-hand sides >= 0                          -> x = 0 (the common case, no factorisation at all)
//   A. min |x_c|^2 s.t. the rows with x_t = 0 feasible    -> x = (0, x_c)
//   B. otherwise proximal-point iterations on  1/2|x_t|^2 + delta/2 |x_c - c_k|^2  (each strictly convex), which
//      converge linearly with ratio ~delta/(delta+sigma) to the unique x_t*, and then
//   C. min |x_c|^2 on the optimal face: rows with a positive multiplier are pinned as equalities.
// Every strictly convex sub-problem is a least-distance problem  min 1/2|y|^2, a~_i.y <= b~_i  in scaled variables
// y = sqrt(d) (x - x0) and is solved by the Goldfarb-Idnani dual active-set method (Math. Prog. 27, 1983) with an
// identity Hessian: Q (orthogonal, nv x nv) and R (upper triangular) of the active normals are updated by one
// Householder reflection per added constraint and Givens rotations per dropped constraint.
// Team mapping (gi_solve): solver state (y, multipliers, active rows, Q, R) lives in shared memory; one thread per constraint
// row in the scan, four threads per entry in the products with Q, the whole team in the rank-one update of Q.
#pragma once
#include "batch.h"
#include "kernels_rbd.cuh"
#include "team_linalg.cuh"
#include "warp_linalg.cuh"

namespace dwbc {

constexpr int QP_NV_MAX = 32;  // one lane per variable
constexpr int QP_NC_MAX = 10 * MAX_ACTIVE_CONTACTS;
constexpr int GI_SCRATCH_DOUBLES = 448;  // >= GiScratch::DOUBLES (static_assert below)

struct HqpSmem {
    int ldm, ldq, ldc, U, AU, At, Q, R, b, tol, rin, bt, gi, vec, total;
    // nvmax: largest number of QP variables of the configured hierarchy (max task dof + RMAX); the arena is sized for it
    __host__ __device__ HqpSmem(int n, int m, int nvmax) {
        ldm = m | 1;
        ldq = nvmax | 1;
        ldc = QP_NC_MAX | 1;  // 21
        int o = 0;
        U = o; o += ldm * nvmax;
        AU = o; o += ldc * nvmax;
        // Atemp (nc x m) is only needed while a level's QP is assembled; it shares its storage with the solver's Q and R
        const int qr = 2 * ldq * nvmax, at = ldc * m;
        At = o; Q = o; R = o + ldq * nvmax; o += qr > at ? qr : at;
        const int mc = 2 * m + QP_NC_MAX;
        b = o; o += mc + 2;
        tol = o; o += mc + 2;
        rin = o; o += mc + 2;   // 1 / scaled row norm
        bt = o; o += mc + 2;    // right-hand side of the current sub-problem
        gi = o; o += GI_SCRATCH_DOUBLES;  // GiScratch
        vec = o; o += 6 * (n + 2) + 2;
        total = o;
    }
};

// ten inequality rows of one 6D contact applied to a wrench given in world axes (reference src/wbd.cpp:59-97,
// src/dwbc.cpp:1055-1075): rows 0-3 ZMP, rows 4-9 friction.  R = world <- contact link (row-major).
__device__ __forceinline__ void contact_rows(const ContactCfg &c, const double *R, Vec3 f, Vec3 mo, double *out) {
    const Vec3 fl = mtv(R, f), ml = mtv(R, mo);
    out[0] = -c.lx * fl.z - ml.y;
    out[1] = -c.lx * fl.z + ml.y;
    out[2] = -c.ly * fl.z - ml.x;
    out[3] = -c.ly * fl.z + ml.x;
    out[4] = fl.x - c.mu * fl.z;
    out[5] = -fl.x - c.mu * fl.z;
    out[6] = fl.y - c.mu * fl.z;
    out[7] = -fl.y - c.mu * fl.z;
    out[8] = ml.z - c.muz * fl.z;
    out[9] = -ml.z - c.muz * fl.z;
}

// View of one QP: rows i < mt: +U / -U (torque limits, mt = 2m or 0), then -AU (contact rows).
struct QpView {
    const double *U, *AU;
    int ldm, ldc, m, mt, nc;  // mt = number of torque rows (0 or 2m), nc = contact rows
    int col0, nv;             // active column window [col0, col0 + nv)
    __device__ __forceinline__ int rows() const { return mt + nc; }
    // a_i[col0 + j]
    __device__ __forceinline__ double a(int i, int j) const {
        if (i < mt) return i < m ? U[i + (col0 + j) * ldm] : -U[i - m + (col0 + j) * ldm];
        return -AU[i - mt + (col0 + j) * ldc];
    }
    // row i as (pointer to its first window entry, column stride, sign)
    __device__ __forceinline__ void row(int i, const double *&p, int &ld, double &sgn) const {
        if (i < mt) {
            p = U + (i < m ? i : i - m) + col0 * ldm;
            ld = ldm;
            sgn = i < m ? 1.0 : -1.0;
        } else {
            p = AU + (i - mt) + col0 * ldc;
            ld = ldc;
            sgn = -1.0;
        }
    }
    // sum_j a(i, j) x[j], two independent partial sums.  Deliberately a compact loop: this is inlined at every row scan of the
    // solver, and the solver's hot loop has to stay within the 32 KB L1.5 instruction cache (profiles/r01_ncu_warm_prox.md:
    // no_instruction was the top stall of hqp_kernel).
    __device__ __forceinline__ double dot(int i, const double *x) const {
        const double *p;
        int ld;
        double sgn;
        row(i, p, ld, sgn);
        double s0 = 0.0, s1 = 0.0;
        int j = 0;
        DWBC_UNROLL(1)
        for (; j + 1 < nv; j += 2) {
            s0 = fma(p[j * ld], x[j], s0);
            s1 = fma(p[(j + 1) * ld], x[j + 1], s1);
        }
        if (j < nv) s0 = fma(p[j * ld], x[j], s0);
        return sgn * (s0 + s1);
    }
};

constexpr int GI_WARPS_MAX = 8;  // TEAM <= 256

// per-state scratch of the active-set solver (shared memory); every array has QP_NV_MAX entries unless noted
struct GiScratch {
    double *y, *u, *dxs, *dv, *zv, *rkv, *isds, *xs, *xp, *xj;
    double *gcs;   // 2 * QP_NV_MAX: Givens (cos, sin) of a drop step
    double *scal;  // 8
    double *red;   // GI_WARPS_MAX
    int *act, *slot;  // QP_NV_MAX each
    int *ired;        // 2 * GI_WARPS_MAX
    uint32_t *pinw, *amask;  // ACTIVE_WORDS each
    double *bt;              // rows doubles
    static constexpr int DOUBLES = 10 * QP_NV_MAX + 2 * QP_NV_MAX + 8 + GI_WARPS_MAX + (2 * QP_NV_MAX + 2 * GI_WARPS_MAX + 2 * ACTIVE_WORDS + 1) / 2;
    __device__ __forceinline__ GiScratch(double *base, double *bt_) {
        y = base; u = y + QP_NV_MAX; dxs = u + QP_NV_MAX; dv = dxs + QP_NV_MAX; zv = dv + QP_NV_MAX; rkv = zv + QP_NV_MAX;
        isds = rkv + QP_NV_MAX; xs = isds + QP_NV_MAX; xp = xs + QP_NV_MAX; xj = xp + QP_NV_MAX;
        gcs = xj + QP_NV_MAX; scal = gcs + 2 * QP_NV_MAX; red = scal + 8;
        act = (int *)(red + GI_WARPS_MAX); slot = act + QP_NV_MAX; ired = slot + QP_NV_MAX;
        pinw = (uint32_t *)(ired + 2 * GI_WARPS_MAX); amask = pinw + ACTIVE_WORDS;
        bt = bt_;
    }
};

static_assert(GiScratch::DOUBLES <= GI_SCRATCH_DOUBLES, "GiScratch arena");
static_assert(TEAM / 32 <= GI_WARPS_MAX && TEAM % 32 == 0, "team size");

struct GiResult {
    int status;  // 1 solved, 0 infeasible / iteration limit
    int iters;
    int q;       // size of the final active set (S.act / S.u hold its rows and multipliers)
};

__device__ __forceinline__ bool bit_test(const uint32_t *w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }

// ---- team reductions through shared memory (every thread gets the result; two barriers each) ----
__device__ __forceinline__ double team_max(double v, const GiScratch &S) {
    v = warp_max(v);
    if (lane_id() == 0) S.red[t_id() >> 5] = v;
    tsync();
    double r = S.red[0];
    for (int w = 1; w < (t_n() >> 5); w++) r = S.red[w] > r ? S.red[w] : r;
    tsync();
    return r;
}
// two maxima with one pair of barriers (red has GI_WARPS_MAX slots: first half for a, second half for b)
__device__ __forceinline__ void team_max2(double &a, double &b, const GiScratch &S) {
    static_assert(2 * (TEAM / 32) <= GI_WARPS_MAX, "team_max2 needs two slots per warp");
    a = warp_max(a);
    b = warp_max(b);
    const int nw = t_n() >> 5;
    if (lane_id() == 0) S.red[t_id() >> 5] = a, S.red[nw + (t_id() >> 5)] = b;
    tsync();
    a = S.red[0], b = S.red[nw];
    for (int w = 1; w < nw; w++) {
        a = S.red[w] > a ? S.red[w] : a;
        b = S.red[nw + w] > b ? S.red[nw + w] : b;
    }
    tsync();
}
__device__ __forceinline__ double team_sum(double v, const GiScratch &S) {
    v = warp_sum(v);
    if (lane_id() == 0) S.red[t_id() >> 5] = v;
    tsync();
    double r = 0.0;
    for (int w = 0; w < (t_n() >> 5); w++) r += S.red[w];
    tsync();
    return r;
}

// Goldfarb-Idnani on  min 1/2|y|^2  s.t.  sum_j a(i,j) dx_j <= bt_i,  dx_j = isd_j y_j  (dx = x - x0), executed by the WHOLE
// team: the row scan takes one row per thread, the products with Q four threads per output, the rank-one update of Q the
// whole team; only the back substitution with R (a dependent chain of q steps) runs on warp 0.  All control decisions are
// taken from values in shared memory after a barrier, so every thread follows the same path.
// isd_j = isd_t for j < tsplit, isd_c otherwise.  `pin` (may be null): rows to be pinned as equalities, added first and never
// dropped.  On return S.dxs holds dx, S.act / S.u the active rows and their multipliers (logical positions < q).
// Deliberately NOT inlined (instruction-cache footprint, profiles/r01_ncu_summary.md).  Synchronises on entry and exit.
// boff: offset of this problem's first row in the right-hand-side array b (2m when the torque rows are left out, mt = 0)
template <int CN, int CM, int CNV, int CMT, int CC>
__device__ __noinline__ GiResult gi_solve(double *sm, int n, int m, int nvmax, int mt, int nc, int col0, int nv, int use_bt, double isd_t,
                                          double isd_c, int tsplit, int use_pin, int max_iter, double ztol, int boff) {
    if constexpr (CN > 0) n = CN, m = CM, nvmax = CNV;  // dimensions known at compile time: strides and offsets fold into the instructions
    if constexpr (CMT >= 0) mt = CMT;
    if constexpr (CC > 0) nc = 10 * (CC / 6);  // contact class of the launch: row counts and the contact window are constants
    // Everything is re-derived from the arena base and a few integers: structs passed by reference to a non-inlined function
    // live in local memory, and with most of L1 carved out as shared memory those loads went to L2 (ncu: 20 % L1 hit rate,
    // long-scoreboard the second largest stall of the first team version).
    const HqpSmem L(n, m, nvmax);
    const GiScratch S(sm + L.gi, sm + L.bt);
    const QpView V{sm + L.U, sm + L.AU, L.ldm, L.ldc, m, mt, nc, col0, nv};
    const double *bt = use_bt ? sm + L.bt : sm + L.b + boff, *tolv = sm + L.tol;
    double *rin = sm + L.rin, *Q = sm + L.Q, *R = sm + L.R;
    const int ldq = L.ldq;
    const uint32_t *pin = use_pin ? S.pinw : nullptr;
    const int tid = t_id(), nt = CN > 0 ? TEAM : t_n(), lane = lane_id(), warp = tid >> 5, nw = nt >> 5;  // (the specialised launches use TEAM threads)
    const int mc = V.rows();
    double *y = S.y, *u = S.u, *dxs = S.dxs, *dv = S.dv, *zv = S.zv, *rkv = S.rkv, *isds = S.isds, *gcs = S.gcs, *scal = S.scal;
    int *act = S.act, *slot = S.slot, *ired = S.ired;
    uint32_t *amask = S.amask;
    tsync();
    DWBC_UNROLL(1)
    for (int e = tid; e < ldq * nv; e += nt) Q[e] = 0.0;
    if (tid < nv) {
        isds[tid] = tid < tsplit ? isd_t : isd_c;
        y[tid] = 0.0;
        dxs[tid] = 0.0;
        u[tid] = 0.0;
        act[tid] = -1;
        slot[tid] = -1;
    }
    if (tid < ACTIVE_WORDS) amask[tid] = 0u;
    tsync();
    if (tid < nv) Q[tid + tid * ldq] = 1.0;
    // 1 / (scaled row norm): violations are compared as distances in y-space.  Rows whose norm is below 1e-10 of the
    // largest row norm are numerically zero (what rounding leaves of a constraint already active at a higher level) and
    // never enter the active set (rin = 0); the oracle applies the same rule.
    double smax = 0.0;
    DWBC_UNROLL(1)
    for (int i = tid; i < mc; i += nt) {
        double s = 0.0;
        DWBC_UNROLL(1)
        for (int j = 0; j < nv; j++) {
            const double v = V.a(i, j) * isds[j];
            s = fma(v, v, s);
        }
        rin[i] = s;
        smax = s > smax ? s : smax;
    }
    smax = team_max(smax, S);
    // A row whose normal is numerically zero IN THIS COLUMN WINDOW cannot be moved by the sub-problem's variables: if it is
    // violated the sub-problem is infeasible (e.g. stage A of solve_level, x_t = 0: a torque-limit row that only the task
    // variables can relieve).  ztol: solve_level's margin for rows that are zero in every column, 1e-9 (1 + max|b|).
    bool zero_row_violated = false;
    DWBC_UNROLL(1)
    for (int i = tid; i < mc; i += nt) {
        const double s = rin[i];
        const bool live = s > 0.0 && s > 1e-20 * smax;
        rin[i] = live ? 1.0 / sqrt(s) : 0.0;
        if (!live && !(pin && bit_test(pin, i)) && bt[i] < -fmax(tolv[i], ztol)) zero_row_violated = true;
    }
    {
        const bool wz = __any_sync(DWBC_FULL, zero_row_violated);
        if (lane == 0) ired[warp] = wz ? 1 : 0;
    }
    tsync();
    GiResult res{0, 0, 0};
    {
        bool anyz = false;
        DWBC_UNROLL(1)
        for (int wI = 0; wI < nw; wI++) anyz = anyz || ired[wI];
        tsync();
        if (anyz) return res;  // status 0: infeasible
    }
    int q = 0, neq = 0;
    uint32_t used = 0;  // physical R columns in use
    int p = -1;  // row being added
    double sp = 0.0, sigma = 1.0, up = 0.0;
    int pin_scan = 0;
    DWBC_UNROLL(1)
    for (int it = 0; it < max_iter; it++) {
        res.iters = it + 1;
        if (p < 0) {
            // ---- choose the next row: pinned rows first (index order), then the most violated inequality ----
            int found = -1;
            if (pin) {
                while (pin_scan < mc) {
                    const int i = pin_scan++;
                    if (bit_test(pin, i)) {
                        found = i;
                        break;
                    }
                }
            }
            if (found >= 0) {
                double part = lane < nv ? V.a(found, lane) * dxs[lane] : 0.0;
                part = warp_sum(part);  // identical in every warp
                const double sl = bt[found] - part;
                p = found;
                sigma = sl <= 0.0 ? 1.0 : -1.0;  // reach the equality from whichever side we are on
                sp = -fabs(sl);
            } else {
                double best = 0.0;
                int bi = -1;
                DWBC_UNROLL(1)
                for (int i = tid; i < mc; i += nt) {
                    const double s1 = bt[i] - V.dot(i, dxs);
                    if (!bit_test(amask, i) && s1 < -tolv[i] && s1 * rin[i] < best) best = s1 * rin[i], bi = i;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(DWBC_FULL, best, o);
                    const int oi = __shfl_xor_sync(DWBC_FULL, bi, o);
                    if (oi >= 0 && (bi < 0 || ob < best || (ob == best && oi < bi))) best = ob, bi = oi;
                }
                if (lane == 0) S.red[warp] = best, ired[warp] = bi;
                tsync();
                best = S.red[0], bi = ired[0];
                DWBC_UNROLL(1)
                for (int w = 1; w < nw; w++) {
                    const double ob = S.red[w];
                    const int oi = ired[w];
                    if (oi >= 0 && (bi < 0 || ob < best || (ob == best && oi < bi))) best = ob, bi = oi;
                }
                if (bi < 0) {
                    res.status = 1;  // no violated row left: optimal
                    break;
                }
                p = bi;
                sigma = 1.0;
                sp = best / rin[p];
            }
            up = 0.0;
        }
        // ---- step direction: n+ = -sigma a~_p, d = Q^T n+ (four threads per entry) ----
        {
            const double *rp;
            int rld;
            double rs;
            V.row(p, rp, rld, rs);
            rs *= -sigma;
            const int part = tid & 3;
            DWBC_UNROLL(1)
            for (int k0 = 0; k0 < nv; k0 += nt >> 2) {
                const int k = k0 + (tid >> 2);
                double a0 = 0.0;
                if (k < nv)
                    DWBC_UNROLL(1)
                    for (int j = part; j < nv; j += 4) a0 = fma(Q[j + k * ldq], rp[j * rld] * isds[j], a0);
                a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
                a0 += __shfl_xor_sync(DWBC_FULL, a0, 2);
                if (k < nv && part == 0) dv[k] = rs * a0;
            }
        }
        tsync();
        // ---- z = Q_2 d_2 (four threads per entry); r = R^{-1} d_1 and the dual step bound on warp 0 ----
        const double dk = lane < nv ? dv[lane] : 0.0;
        const double zz = warp_sum((lane >= q && lane < nv) ? dk * dk : 0.0);  // identical in every warp
        const double nn = warp_sum(dk * dk);
        {
            const int part = tid & 3;
            DWBC_UNROLL(1)
            for (int j0 = 0; j0 < nv; j0 += nt >> 2) {
                const int j = j0 + (tid >> 2);
                double a0 = 0.0;
                if (j < nv)
                    DWBC_UNROLL(1)
                    for (int k = q + part; k < nv; k += 4) a0 = fma(Q[j + k * ldq], dv[k], a0);
                a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
                a0 += __shfl_xor_sync(DWBC_FULL, a0, 2);
                if (j < nv && part == 0) zv[j] = a0;
            }
        }
        if (warp == 0) {
            const int myslot = lane < q ? slot[lane] : 0;
            const double invd = lane < q ? 1.0 / R[lane + myslot * ldq] : 0.0;
            double rk = (lane < q) ? dk : 0.0;
            DWBC_UNROLL(1)
            for (int k = q - 1; k >= 0; k--) {
                const int sk = __shfl_sync(DWBC_FULL, myslot, k);
                const double val = __shfl_sync(DWBC_FULL, rk, k) * __shfl_sync(DWBC_FULL, invd, k);
                if (lane == k) rk = val;
                if (lane < k) rk = fma(-R[lane + sk * ldq], val, rk);
            }
            // largest dual step that keeps the multipliers of the droppable active rows non-negative
            double t1 = 1e300;
            int l1 = -1;
            if (lane >= neq && lane < q && rk > 0.0) t1 = u[lane] / rk, l1 = lane;
            for (int o = 16; o > 0; o >>= 1) {
                const double ot = __shfl_xor_sync(DWBC_FULL, t1, o);
                const int ol = __shfl_xor_sync(DWBC_FULL, l1, o);
                if (ol >= 0 && (l1 < 0 || ot < t1 || (ot == t1 && ol < l1))) t1 = ot, l1 = ol;
            }
            if (lane < q) rkv[lane] = rk;
            if (lane == 0) scal[0] = t1, ired[GI_WARPS_MAX] = l1;
        }
        tsync();
        const double t1 = scal[0];
        const int l1 = ired[GI_WARPS_MAX];
        const bool dependent = (q >= nv) || !(zz > 1e-22 * nn);
        if (dependent && l1 < 0) {
            // no primal and no dual step possible
            const bool is_pin = pin && bit_test(pin, p);
            // A redundant row at a degenerate vertex (more active rows than variables) shows up here "violated" by what rounding
            // leaves: judge it as a distance in y-space (the geometry the iteration works in -- with a small proximal weight the
            // contact columns are scaled up by 1e3 and a residual of 4e-9 in row units is 4e-12 of distance), relative to |y|.
            const double yy = warp_sum(lane < nv ? y[lane] * y[lane] : 0.0);  // identical in every warp
            if (is_pin || -sp <= tolv[p] || -sp * rin[p] <= 1e-10 * (1.0 + sqrt(yy))) {
                // a pinned row that depends on the rows pinned so far (consistent by construction), or a row
                // that is violated by less than its tolerance: nothing to do for it
                p = -1;
                tsync();
                continue;
            }
#ifdef DWBC_CPU_EMUL
            if (tid == 0 && getenv("DWBC_QP_DEBUG"))
                fprintf(stderr, "gi_solve infeasible: it %d row %d sp %.3e tol %.3e q %d nv %d zz %.3e nn %.3e col0 %d\n", it, p, sp, tolv[p], q, nv, zz, nn, col0);
#endif
            res.status = 0;  // infeasible
            break;
        }
        const double t2 = dependent ? 1e300 : -sp / zz;
        const double t = t1 < t2 ? t1 : t2;
        if (tid < q) u[tid] = fma(-t, rkv[tid], u[tid]);
        up += t;
        if (!dependent && tid < nv) {
            const double yn = fma(t, zv[tid], y[tid]);
            y[tid] = yn;
            dxs[tid] = yn * isds[tid];
        }
        if (!dependent && t2 <= t1) {
            // ---- full step: row p becomes active.  One Householder reflection on Q_2 maps d_2 to alpha e_q:
            // h = d_2 - alpha e_q,  Q_2 <- Q_2 - w h^T,  w = tau Q_2 h = tau (z - alpha Q[:, q]) ----
            const double dq = dv[q];
            const double alpha = dq > 0.0 ? -sqrt(zz) : sqrt(zz);
            const double tau = 1.0 / (zz - alpha * dq);
            const int part = tid & 3;
            DWBC_UNROLL(1)
            for (int j0 = 0; j0 < nv; j0 += nt >> 2) {
                const int j = j0 + (tid >> 2);
                const double wj = j < nv ? tau * (zv[j] - alpha * Q[j + q * ldq]) : 0.0;
                __syncwarp();  // the four threads of row j have read Q[j, q] before it is overwritten
                if (j < nv)
                    DWBC_UNROLL(1)
                    for (int k = q + part; k < nv; k += 4) {
                        const double hk = k == q ? dq - alpha : dv[k];
                        Q[j + k * ldq] = fma(-wj, hk, Q[j + k * ldq]);
                    }
            }
            const int ns = __ffs((int)~used) - 1;
            used |= 1u << ns;
            if (tid < q) R[tid + ns * ldq] = dv[tid];
            if (tid == q) {
                R[q + ns * ldq] = alpha;
                act[q] = p;
                u[q] = up;
                slot[q] = ns;
            }
            if (tid == 0) amask[p >> 5] |= 1u << (p & 31);
            if (pin && bit_test(pin, p)) neq = q + 1;
            q++;
            p = -1;
            tsync();
        } else {
            // ---- partial (or purely dual) step: the blocking row at logical position l1 leaves the active set ----
            const int l = l1;
            const int dropped = act[l];
            const int dslot = slot[l];
            if (warp == 0) {
                // Givens rotations of rows k, k+1 of R: lane j owns logical column j (only it touches that column)
                double *col = R + (lane < q ? slot[lane] : 0) * ldq;
                DWBC_UNROLL(1)
                for (int k = l; k + 1 < q; k++) {
                    double a = 0.0, bb = 0.0;
                    if (lane == k + 1) a = col[k], bb = col[k + 1];
                    a = __shfl_sync(DWBC_FULL, a, k + 1);
                    bb = __shfl_sync(DWBC_FULL, bb, k + 1);
                    const double h = sqrt(a * a + bb * bb);
                    const double cs = h > 0.0 ? a / h : 1.0, sn = h > 0.0 ? bb / h : 0.0;
                    if (lane > k && lane < q) {
                        const double r0 = col[k], r1 = col[k + 1];
                        col[k] = cs * r0 + sn * r1;
                        col[k + 1] = -sn * r0 + cs * r1;
                    }
                    if (lane == 0) gcs[2 * k] = cs, gcs[2 * k + 1] = sn;
                }
            }
            tsync();  // every thread has read act[l] / slot[l]; the rotations are published
            if (tid < nv) {
                double q0 = Q[tid + l * ldq];
                DWBC_UNROLL(1)
                for (int k = l; k + 1 < q; k++) {
                    const double cs = gcs[2 * k], sn = gcs[2 * k + 1];
                    const double q1 = Q[tid + (k + 1) * ldq];
                    Q[tid + k * ldq] = cs * q0 + sn * q1;
                    q0 = -sn * q0 + cs * q1;
                }
                if (l + 1 < q) Q[tid + (q - 1) * ldq] = q0;
            }
            if (warp == 0) {
                const int a2 = lane + 1 < QP_NV_MAX ? act[lane + 1] : -1;
                const int s2 = lane + 1 < QP_NV_MAX ? slot[lane + 1] : -1;
                const double u2 = lane + 1 < QP_NV_MAX ? u[lane + 1] : 0.0;
                __syncwarp();
                if (lane >= l && lane < q - 1) act[lane] = a2, slot[lane] = s2, u[lane] = u2;
                if (lane == q - 1) act[lane] = -1, slot[lane] = -1, u[lane] = 0.0;
            }
            used &= ~(1u << dslot);
            if (tid == 0) amask[dropped >> 5] &= ~(1u << (dropped & 31));
            q--;
            // slack of row p at the new point
            double part = lane < nv ? V.a(p, lane) * dxs[lane] : 0.0;
            part = warp_sum(part);
            sp = sigma * (bt[p] - part);
            if (sp > 0.0) sp = 0.0;
            tsync();
        }
    }
    res.q = q;
#ifdef DWBC_CPU_EMUL
    if (tid == 0 && getenv("DWBC_QP_TRACE")) fprintf(stderr, "gi_solve nv %d col0 %d pin %d iters %d status %d q %d\n", nv, col0, use_pin, res.iters, res.status, q);
#endif
    tsync();
    return res;
}

// Re-solve after a change of the right-hand side only (the next proximal-point iterate): with the factorisation Q, R and the
// active set left by the previous gi_solve, the candidate  y = Q_1 w,  R^T w = -bt_W  is the least-norm point on the active
// rows and  lambda = R^{-1} w  its multipliers.  If lambda >= 0 and every other row is satisfied this is the (unique) optimum
// of the new sub-problem -- at the cost of about one active-set iteration; otherwise nothing is modified and the caller falls
// back to a cold gi_solve.  Same team conventions as gi_solve; synchronises on entry and exit.
template <int CN, int CM, int CNV, int CMT, int CC>
__device__ __noinline__ bool gi_warm(double *sm, int n, int m, int nvmax, int mt, int nc, int col0, int nv, int q, int use_bt) {
    if constexpr (CN > 0) n = CN, m = CM, nvmax = CNV;  // dimensions known at compile time: strides and offsets fold into the instructions
    if constexpr (CMT >= 0) mt = CMT;
    if constexpr (CC > 0) nc = 10 * (CC / 6);  // contact class of the launch: row counts and the contact window are constants
    const HqpSmem L(n, m, nvmax);
    const GiScratch S(sm + L.gi, sm + L.bt);
    const QpView V{sm + L.U, sm + L.AU, L.ldm, L.ldc, m, mt, nc, col0, nv};
    const double *bt = use_bt ? sm + L.bt : sm + L.b, *tolv = sm + L.tol, *rin = sm + L.rin, *Q = sm + L.Q, *R = sm + L.R;
    const int ldq = L.ldq;
    const int tid = t_id(), nt = CN > 0 ? TEAM : t_n(), lane = lane_id(), warp = tid >> 5, nw = nt >> 5;  // (the specialised launches use TEAM threads)
    const int mc = V.rows();
    double *wv = S.dv, *lam = S.rkv, *yn = S.zv;  // scratch vectors of gi_solve, free between calls
    tsync();
    if (warp == 0) {
        const int myslot = lane < q ? S.slot[lane] : 0;
        const double *col = R + myslot * ldq;
        const double invd = lane < q ? 1.0 / col[lane] : 0.0;
        double rhs = lane < q ? -bt[S.act[lane]] : 0.0;
        DWBC_UNROLL(1)
        for (int k = 0; k < q; k++) {  // forward substitution with R^T (column j of R is owned by lane j)
            const double val = __shfl_sync(DWBC_FULL, rhs, k) * __shfl_sync(DWBC_FULL, invd, k);
            if (lane == k) rhs = val;
            if (lane > k && lane < q) rhs = fma(-col[k], val, rhs);
        }
        const double w = rhs;
        double rk = w;
        DWBC_UNROLL(1)
        for (int k = q - 1; k >= 0; k--) {  // back substitution with R
            const int sk = __shfl_sync(DWBC_FULL, myslot, k);
            const double val = __shfl_sync(DWBC_FULL, rk, k) * __shfl_sync(DWBC_FULL, invd, k);
            if (lane == k) rk = val;
            if (lane < k) rk = fma(-R[lane + sk * ldq], val, rk);
        }
        const double lmax = warp_max(lane < q ? fabs(rk) : 0.0);
        const bool neg = lane < q && rk < -1e-12 * lmax;
        const bool anyneg = __any_sync(DWBC_FULL, neg);
        if (lane < q) wv[lane] = w, lam[lane] = rk;
        if (lane == 0) S.ired[GI_WARPS_MAX] = anyneg ? 1 : 0;
    }
    tsync();
    if (S.ired[GI_WARPS_MAX]) {
        tsync();
        return false;
    }
    {
        const int part = tid & 3;
        DWBC_UNROLL(1)
        for (int j0 = 0; j0 < nv; j0 += nt >> 2) {
            const int j = j0 + (tid >> 2);
            double a0 = 0.0;
            if (j < nv)
                DWBC_UNROLL(1)
                for (int k = part; k < q; k += 4) a0 = fma(Q[j + k * ldq], wv[k], a0);
            a0 += __shfl_xor_sync(DWBC_FULL, a0, 1);
            a0 += __shfl_xor_sync(DWBC_FULL, a0, 2);
            if (j < nv && part == 0) yn[j] = a0 * S.isds[j];  // dx = y o isd
        }
    }
    tsync();
    bool viol = false;
    DWBC_UNROLL(1)
    for (int i = tid; i < mc; i += nt) {
        const double s1 = bt[i] - V.dot(i, yn);
        if (!bit_test(S.amask, i) && s1 < -tolv[i] && rin[i] > 0.0) viol = true;
    }
    const bool wviol = __any_sync(DWBC_FULL, viol);
    if (lane == 0) S.ired[warp] = wviol ? 1 : 0;
    tsync();
    bool bad = false;
    DWBC_UNROLL(1)
    for (int w = 0; w < nw; w++) bad = bad || S.ired[w];
    if (!bad && tid < nv) {
        S.dxs[tid] = yn[tid];
        S.y[tid] = yn[tid] / S.isds[tid];
        if (tid < q) S.u[tid] = lam[tid] > 0.0 ? lam[tid] : 0.0;
    }
    tsync();
    return !bad;
}


// active_out[k] = bits of mask (ACTIVE_WORDS words in shared memory) shifted up by `shift` rows
__device__ __forceinline__ uint32_t shifted_word(const uint32_t *mask, int k, int shift) {
    const int ws = shift >> 5, bs = shift & 31;
    const int a = k - ws;
    uint32_t lo = (a >= 0 && a < ACTIVE_WORDS) ? mask[a] : 0u;
    uint32_t hi = (a - 1 >= 0 && a - 1 < ACTIVE_WORDS) ? mask[a - 1] : 0u;
    return bs ? ((lo << bs) | (hi >> (32 - bs))) : lo;
}

struct LevelOut {
    int status;
    int iters;
    int slow;  // 1: the proximal-point loop stopped at its iteration cap; the last (primal-feasible) iterate was kept
};

// Solves one QP (t task variables first, then r contact variables) with the whole team.  U (m x nv), AU (nc x nv) and b (rows)
// are in shared memory; the solution is left in S.xj.  active_out (global, ACTIVE_WORDS words): rows with zero slack at the
// solution.  Synchronises on entry and exit.
// boff / row_shift: the problem's rows start at b[boff] and are reported as rows row_shift.. (a solve without the torque rows: mt = 0,
// boff = row_shift = 2m keeps the reference's row numbering in the active set).
template <int CN, int CM, int CNV, int CMT, int CC>
__device__ __noinline__ LevelOut solve_level(double *sm, int n, int m, int nvmax, int t, int r, int mt, int nc, uint32_t *active_out, int boff,
                                             int row_shift, int c0, int skip_a) {
    if constexpr (CN > 0) n = CN, m = CM, nvmax = CNV;  // dimensions known at compile time: strides and offsets fold into the instructions
    if constexpr (CMT >= 0) mt = CMT;
    if constexpr (CC > 0) nc = 10 * (CC / 6);  // contact class of the launch: row counts and the contact window are constants
    if constexpr (CC > 0) r = CC - 6;
    const HqpSmem L(n, m, nvmax);
    const GiScratch S(sm + L.gi, sm + L.bt);
    const double *b = sm + L.b + boff;
    double *tolv = sm + L.tol, *rin = sm + L.rin;
    const int tid = t_id(), nt = CN > 0 ? TEAM : t_n(), lane = lane_id(), warp = tid >> 5, nw = nt >> 5;  // (the specialised launches use TEAM threads)
    const int nv = t + r, mc = mt + nc;
    LevelOut out{1, 0, 0};
    const QpView V{sm + L.U, sm + L.AU, L.ldm, L.ldc, m, mt, nc, c0, nv};
    double *xj = S.xj, *xs = S.xs, *xp = S.xp;
    tsync();
    if (tid < QP_NV_MAX) xj[tid] = 0.0;
    // row tolerances in the units of the rows, the quick exit, and the active rows of the quick exit.  Numerically zero
    // rows (norm <= 1e-10 of the largest) cannot be influenced by x: they only decide feasibility (b >= -1e-9 (1 + max|b|)).
    double smax = 0.0;
    DWBC_UNROLL(1)
    for (int i = tid; i < mc; i += nt) {
        double s = 0.0;
        DWBC_UNROLL(1)
        for (int j = 0; j < nv; j++) {
            const double v = V.a(i, j);
            s = fma(v, v, s);
        }
        rin[i] = s;
        smax = s > smax ? s : smax;
    }
    // magnitude of the right-hand sides: a numerically zero row is "violated" only beyond 1e-9 of it (its b is a difference of
    // quantities of that size, e.g. tlim - tau where an earlier level left the torque limit active to ~1e-10 relative)
    double babs = 0.0;
    DWBC_UNROLL(1)
    for (int i = tid; i < mc; i += nt) babs = fmax(babs, fabs(b[i]));
    team_max2(smax, babs, S);
    const double zero_row_tol = 1e-9 * (1.0 + babs);
    double bmin = 1e300;
    bool infeasible = false;
    DWBC_UNROLL(1)
    for (int i0 = 0; i0 < mc; i0 += nt) {
        const int i = i0 + tid;
        bool on = false;
        if (i < mc) {
            const double s = rin[i];
            const double sc = fmax(sqrt(s), fabs(b[i]));
            tolv[i] = 1e-10 * sc;
            if (s > 1e-20 * smax && s >= 1e-28)
                bmin = fmin(bmin, b[i] + 1e-10 * sc);
            else if (b[i] < -zero_row_tol)
                infeasible = true;
            on = s > 0.0 && b[i] < 1e-9 * (1.0 + fabs(b[i]));
        }
        const unsigned bal = __ballot_sync(DWBC_FULL, on);
        const int word = (i0 >> 5) + warp;
        if (lane == 0 && word < ACTIVE_WORDS) S.pinw[word] = bal;
    }
    bmin = warp_min(bmin);
    const bool winf = __any_sync(DWBC_FULL, infeasible);
    if (lane == 0) S.red[warp] = bmin, S.ired[warp] = winf ? 1 : 0;
    tsync();
    bool any_inf = false;
    bmin = 1e300;
    DWBC_UNROLL(1)
    for (int w = 0; w < nw; w++) {
        bmin = fmin(bmin, S.red[w]);
        any_inf = any_inf || S.ired[w];
    }
    tsync();
    if (any_inf) {
        out.status = 0;
        return out;
    }
    if (bmin >= 0.0) {
        // x = 0 satisfies every row: optimal for both stages, no factorisation needed
        if (tid < ACTIVE_WORDS && tid >= (mc + 31) / 32) S.pinw[tid] = 0u;
        tsync();
        if (active_out && tid < ACTIVE_WORDS) active_out[tid] = shifted_word(S.pinw, tid, row_shift);
        return out;
    }
    const int max_iter = 20 * (nv + 8);
    bool done = false;
    if (t == 0 || r == 0) {
        GiResult g = gi_solve<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, mt, nc, c0, nv, 0, 1.0, 1.0, nv, 0, max_iter, zero_row_tol, boff);
        out.iters += g.iters;
        out.status = g.status;
        if (tid < nv) xj[tid] = S.dxs[tid];
        done = true;
    }
    if (!done && !skip_a) {
        // ---- A: is x_t = 0 attainable?  (looser tolerances: the set may be a thin face inherited from the levels above)
        DWBC_UNROLL(1)
        for (int i = tid; i < mc; i += nt) tolv[i] *= 100.0;
        GiResult g = gi_solve<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, mt, nc, c0 + t, r, 0, 1.0, 1.0, r, 0, max_iter, zero_row_tol, boff);
        out.iters += g.iters;
        if (g.status) {
            if (tid < r) xj[t + tid] = S.dxs[tid];
            done = true;
        }
        DWBC_UNROLL(1)
        for (int i = tid; i < mc; i += nt) tolv[i] *= 0.01;
        tsync();
    }
    if (!done) {
        // ---- B: proximal point on 1/2|x_t|^2 + delta/2 |x_c - c|^2 ----
        double ntn = 0.0, ncn = 0.0;
        DWBC_UNROLL(1)
        for (int i = tid; i < mc; i += nt)
            DWBC_UNROLL(1)
            for (int j = 0; j < nv; j++) {
                const double v = V.a(i, j);
                if (j < t) ntn = fma(v, v, ntn);
                else ncn = fma(v, v, ncn);
            }
        ntn = team_sum(ntn, S);
        ncn = team_sum(ncn, S);
        double delta = 1e-5 * (ncn / r) / fmax(ntn / t, 1e-300);
        double isd_c = rsqrt(delta);
        double *bt = S.bt;
        if (tid < nv) xs[tid] = 0.0, xp[tid] = 0.0;  // xs: prox centre (zero on the task columns), xp: previous iterate
        tsync();
        bool conv = false, force_cold = false;
        double dd_prev = 0.0;
        int reductions = 0;
        GiResult g{0, 0, 0};
        DWBC_UNROLL(1)
        for (int k = 0; k < 40; k++) {
            DWBC_UNROLL(1)
            for (int i = tid; i < mc; i += nt) {  // bt = b - A x0 (x0 is zero on the task columns)
                double s = b[i];
                DWBC_UNROLL(1)
                for (int j = t; j < nv; j++) s = fma(-V.a(i, j), xs[j], s);
                bt[i] = s;
            }
            if (k > 0 && !force_cold && gi_warm<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, mt, nc, c0, nv, g.q, 1)) {
                out.iters += 1;
            } else {
                g = gi_solve<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, mt, nc, c0, nv, 1, 1.0, isd_c, t, 0, max_iter, zero_row_tol, boff);
                out.iters += g.iters;
            }
            force_cold = false;
            if (!g.status) {
                out.status = 0;
                break;
            }
            const double xn = lane < nv ? xs[lane] + S.dxs[lane] : 0.0;
            const double xo = lane < nv ? xp[lane] : 0.0;
            const double dd = warp_sum(lane < t ? (xn - xo) * (xn - xo) : 0.0);  // identical in every warp
            const double xx = warp_sum(lane < t ? xn * xn : 0.0);
            tsync();
            if (tid < nv) {
                xp[tid] = xn;
                xj[tid] = xn;
                xs[tid] = tid >= t ? xn : 0.0;
            }
            tsync();
            if (k > 0 && sqrt(dd) <= 1e-12 * (1.0 + sqrt(xx))) {
                conv = true;
                break;
            }
            // The iteration contracts with ratio delta / (delta + sigma), sigma = curvature of the optimal face: on a nearly
            // degenerate face (x_t = 0 attainable only in the limit) the default delta crawls (ratio 0.97 seen: 36 of 65 536 states of
            // the headline workload).  A step that shrinks by less than half switches to a 100 times smaller delta (cold re-solve:
            // the scaling of the factorisation changes); any delta gives the same limit, the unique lexicographic optimum.
            if (k >= 2 && reductions < 3 && dd > 0.25 * dd_prev) {
                delta *= 0.01;
                isd_c = rsqrt(delta);
                reductions++;
                force_cold = true;
            }
            dd_prev = dd;
        }
        // Every proximal iterate is primal feasible and the sequence decreases |x_t| monotonically: a state that has not met the
        // 1e-12 step test within the cap keeps its last iterate (qpOASES, with nWSR = 1000, would also return a point for a
        // feasible problem; reference src/dwbc.cpp:1080) and is flagged DWBC_ST_QP_SLOW instead of being reported as failed.
        if (out.status && !conv) out.slow = 1;
#ifdef DWBC_CPU_EMUL
        if (tid == 0 && getenv("DWBC_QP_TRACE")) fprintf(stderr, "prox t %d r %d iters_so_far %d conv %d reductions %d\n", t, r, out.iters, (int)conv, reductions);
#endif
        if (out.status) {
            // ---- C: min |x_c|^2 on the optimal face; rows with a positive multiplier are pinned ----
            const double umax = warp_max(lane < g.q ? S.u[lane] : 0.0);  // identical in every warp
            if (tid < ACTIVE_WORDS) S.pinw[tid] = 0u;
            tsync();
            if (tid == 0)
                DWBC_UNROLL(1)
                for (int k = 0; k < g.q; k++) {
                    const int ak = S.act[k];
                    if (ak >= 0 && S.u[k] > 1e-9 * umax) S.pinw[ak >> 5] |= 1u << (ak & 31);
                }
            DWBC_UNROLL(1)
            for (int i = tid; i < mc; i += nt) {
                double s = b[i];
                DWBC_UNROLL(1)
                for (int j = 0; j < t; j++) s = fma(-V.a(i, j), xj[j], s);
                bt[i] = s;
                tolv[i] *= 100.0;
            }
            GiResult gc = gi_solve<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, mt, nc, c0 + t, r, 1, 1.0, 1.0, r, 1, max_iter, zero_row_tol, boff);
            out.iters += gc.iters;
            if (!gc.status) out.status = 0;
            if (tid < r) xj[t + tid] = S.dxs[tid];
        }
    }
    tsync();
    if (!out.status && tid < QP_NV_MAX) xj[tid] = 0.0;
    tsync();
    // active rows at the solution: slack within 1e-9 (1 + |b|), rows with a zero normal excluded (the oracle's rule)
    DWBC_UNROLL(1)
    for (int i0 = 0; i0 < mc; i0 += nt) {
        const int i = i0 + tid;
        bool on = false;
        if (i < mc && out.status) {
            double s = -b[i], nn2 = 0.0;
            DWBC_UNROLL(1)
            for (int j = 0; j < nv; j++) {
                const double a = V.a(i, j);
                s = fma(a, xj[j], s);
                nn2 = fma(a, a, nn2);
            }
            on = nn2 > 0.0 && s > -1e-9 * (1.0 + fabs(b[i]));
        }
        const unsigned bal = __ballot_sync(DWBC_FULL, on);
        const int word = (i0 >> 5) + warp;
        if (lane == 0 && word < ACTIVE_WORDS) S.pinw[word] = bal;
    }
    if (tid < ACTIVE_WORDS && tid >= (mc + 31) / 32) S.pinw[tid] = 0u;
    tsync();
    if (active_out && tid < ACTIVE_WORDS) active_out[tid] = shifted_word(S.pinw, tid, row_shift);
    tsync();
    return out;
}

// The contact-variable sub-problem  min |x_c|^2  s.t. every row holds with x_t = 0  (stage A of solve_level; for the redistribution
// QP, which has no task variables, the whole problem).  Its matrix -- the columns [NwJw] of U and Atemp NwJw of AU, at columns
// [cc, cc + r) -- is the SAME for every level of the hierarchy and for the redistribution of one state; only the right-hand side b
// moves with the torque of the levels above.  So the caller assembles those r columns once per state and this function keeps the
// factorisation of the last solve (aq = number of active rows it left, -1: none) and first tries gi_warm on it: about one active-set
// iteration when the active rows stay the same from level to level, which is the common case (a friction / ZMP row that the first
// level leaves active stays active below).  Falls back to a cold gi_solve.  The window norms are kept in the bt array (unused here:
// the right-hand side is b itself).  t = task dof of this level (only to place the answer in S.xj = [0_t; x_c] and, for the few
// torque rows that end with zero slack, to evaluate the task part of their normal from Ntq in global memory); loose: the 100 times
// wider row tolerances of stage A.  status 0: x_t = 0 is not attainable (nothing usable is left in S.xj).
template <int CN, int CM, int CNV, int CMT, int CC>
__device__ __noinline__ LevelOut solve_contact_only(double *sm, int n, int m, int nvmax, int t, int r, int mt, int nc, uint32_t *active_out, int cc,
                                                    int *aq_io, int loose, const double *__restrict__ Ntq) {
    if constexpr (CN > 0) n = CN, m = CM, nvmax = CNV;  // dimensions known at compile time: strides and offsets fold into the instructions
    if constexpr (CMT >= 0) mt = CMT;
    if constexpr (CC > 0) nc = 10 * (CC / 6);  // contact class of the launch: row counts and the contact window are constants
    if constexpr (CC > 0) r = CC - 6;
    const HqpSmem L(n, m, nvmax);
    const GiScratch S(sm + L.gi, sm + L.bt);
    const double *b = sm + L.b;
    double *tolv = sm + L.tol, *nrm = sm + L.bt;
    const int tid = t_id(), nt = CN > 0 ? TEAM : t_n(), lane = lane_id(), warp = tid >> 5, nw = nt >> 5;  // (the specialised launches use TEAM threads)
    const int mc = mt + nc;
    const QpView V{sm + L.U, sm + L.AU, L.ldm, L.ldc, m, mt, nc, cc, r};
    LevelOut out{1, 0, 0};
    int aq = *aq_io;
    tsync();
    if (aq < 0) {  // (re)build the window norms: a full solve of an earlier level used bt
        DWBC_UNROLL(1)
        for (int i = tid; i < mc; i += nt) {
            double s = 0.0;
            DWBC_UNROLL(1)
            for (int j = 0; j < r; j++) {
                const double v = V.a(i, j);
                s = fma(v, v, s);
            }
            nrm[i] = s;
        }
    }
    double smax = 0.0, babs = 0.0;
    tsync();
    DWBC_UNROLL(1)
    for (int i = tid; i < mc; i += nt) smax = fmax(smax, nrm[i]), babs = fmax(babs, fabs(b[i]));
    team_max2(smax, babs, S);
    const double ztol = 1e-9 * (1.0 + babs);
    const double rel = loose ? 1e-8 : 1e-10;
    bool infeasible = false;
    DWBC_UNROLL(1)
    for (int i = tid; i < mc; i += nt) {
        const double s = nrm[i];
        const double tl = rel * fmax(sqrt(s), fabs(b[i]));
        tolv[i] = tl;
        // a row the contact variables cannot move (e.g. the torque limit of an upper-body joint) must already hold
        if (!(s > 1e-20 * smax && s >= 1e-28) && b[i] < -fmax(tl, ztol)) infeasible = true;
    }
    const bool winf = __any_sync(DWBC_FULL, infeasible);
    if (lane == 0) S.ired[warp] = winf ? 1 : 0;
    tsync();
    bool any_inf = false;
    DWBC_UNROLL(1)
    for (int w = 0; w < nw; w++) any_inf = any_inf || S.ired[w];
    tsync();
    if (any_inf) {
        out.status = 0;
        return out;
    }
    if (aq >= 0 && gi_warm<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, mt, nc, cc, r, aq, 0)) {
        out.iters = 1;
#ifdef DWBC_CPU_EMUL
        if (tid == 0 && getenv("DWBC_QP_TRACE")) fprintf(stderr, "contact_only warm t %d q %d\n", t, aq);
#endif
    } else {
#ifdef DWBC_CPU_EMUL
        if (tid == 0 && getenv("DWBC_QP_TRACE")) fprintf(stderr, "contact_only cold t %d (kept q %d)\n", t, aq);
#endif
        const GiResult g = gi_solve<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, mt, nc, cc, r, 0, 1.0, 1.0, r, 0, 20 * (r + 8), ztol, 0);
        out.iters = g.iters;
        out.status = g.status;
        aq = g.status ? g.q : -1;
        *aq_io = aq;
        if (!g.status) return out;
    }
    if (tid < QP_NV_MAX) S.xj[tid] = (tid >= t && tid < t + r) ? S.dxs[tid - t] : 0.0;
    tsync();
    // active rows at the solution (rule of solve_level: slack within 1e-9 (1 + |b|), rows with a zero normal excluded)
    DWBC_UNROLL(1)
    for (int i0 = 0; i0 < mc; i0 += nt) {
        const int i = i0 + tid;
        bool on = false;
        if (i < mc) {
            double sl = -b[i];
            DWBC_UNROLL(1)
            for (int j = 0; j < r; j++) sl = fma(V.a(i, j), S.xj[t + j], sl);
            if (sl > -1e-9 * (1.0 + fabs(b[i]))) {
                double nn2 = nrm[i];
                if (!(nn2 > 0.0) && i < mt && Ntq) {
                    const int row = i < m ? i : i - m;
                    DWBC_UNROLL(1)
                    for (int j = 0; j < t; j++) nn2 = fma(Ntq[row + j * m], Ntq[row + j * m], nn2);
                }
                on = nn2 > 0.0;
            }
        }
        const unsigned bal = __ballot_sync(DWBC_FULL, on);
        const int word = (i0 >> 5) + warp;
        if (lane == 0 && word < ACTIVE_WORDS) S.pinw[word] = bal;
    }
    if (tid < ACTIVE_WORDS && tid >= (mc + 31) / 32) S.pinw[tid] = 0u;
    tsync();
    if (active_out && tid < ACTIVE_WORDS) active_out[tid] = S.pinw[tid];
    tsync();
    return out;
}

// One CTA (team) per state: QP assembly (U, Atemp U, right-hand sides), the active-set iterations (gi_solve) and the torque
// updates all use every thread of the team.
// list / count (may be null): the compact list of states left QP-active by lean_cycle_kernel (kernels_lean.cuh); CTA i finishes state
// list[i] and CTAs beyond *count exit at once.  Without a list CTA i owns state i.
#ifndef DWBC_HQP_CTAS
#define DWBC_HQP_CTAS 7  // states per SM (72 registers); 6 (80 registers) measured slower in both rounds (round 2: 7.18 vs 6.89 ms)
#endif
// (Out-of-line warp / team reductions shrink this kernel from 9.7 k to 8.8 k SASS instructions but run slower: 7.33 vs 7.02 ms.)
template <int CN, int CM, int CNV, int CMT, int CC>
__global__ void __launch_bounds__(TEAM, DWBC_HQP_CTAS) hqp_kernel_t(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, int hqp, int redist,
                                                        int nvmax, const int *__restrict__ list, const int *__restrict__ count) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = CN > 0 ? TEAM : t_n();
    if (list && (int)blockIdx.x >= *count) return;
    const long long s = list ? list[blockIdx.x] : blockIdx.x;
    const int n = CN > 0 ? CN : w.n, m = CN > 0 ? CM : w.m, nb = w.nb;
    if constexpr (CN > 0) nvmax = CNV;
    const HqpSmem L(n, m, nvmax);
    double *sm = (double *)smem;
    double *U = sm + L.U, *AU = sm + L.AU, *At = sm + L.At, *b = sm + L.b;
    const GiScratch S(sm + L.gi, sm + L.bt);
    double *vec = sm + L.vec;
    const int vs = n + 2;
    double *tprev = vec, *base = vec + vs, *fs = vec + 2 * vs, *xs = vec + 4 * vs, *lf = vec + 5 * vs;
    const int ldm = L.ldm, ldc = L.ldc;
    int st = w.status[s];
    const bool contact_ok = st & 1;
    const bool task_ok = (st & 8) != 0;  // set by task_space_kernel when hqp = 1
    st &= ~8;
    const int c = CC > 0 ? CC : (contact_ok ? w.cdof[s] : 0);  // (a class launch only lists states with that many contact dof)
    const int r = c > 6 ? c - 6 : 0;
    const int kc = c / 6;
    const int nc = 10 * kc;
    const int mt = CMT >= 0 ? CMT : (cfg->tlim_set ? 2 * m : 0);
    int total_iters = 0;
    // diagnostics (builds with -DDWBC_HQP_TIMING only: the bookkeeping costs registers of a kernel compiled for seven states per SM):
    // cycles per phase accumulated in slots 21.. of the state's stamp record (tools/lean_phases.py)
#ifdef DWBC_HQP_TIMING
    long long *tm = w.hqp_timing ? w.hqp_timing + s * 32 : nullptr;
    long long tlast = tm ? clock64() : 0;
    auto lap = [&](int k) {
        if (tm && tid == 0) {
            const long long now = clock64();
            tm[k] += now - tlast;
            tlast = now;
        }
    };
    if (tm && tid < 11) tm[21 + tid] = 0;
#else
    constexpr long long *tm = nullptr;
    auto lap = [](int) {};
#endif
    double *tau_task = w.tau_task + s * m, *tau_contact = w.tau_contact + s * m;
    if (!contact_ok) {  // uniform over the CTA
        if (tid == 0) w.status[s] = st, w.qp_iters[s] = 0;
        return;
    }
    // ---- Atemp = A_const A_rot Jbar^T[:, 6:]  (nc x m).  Rebuilt when a QP needs it again: the solver's Q / R overwrite its storage
    // (a few hundred flops from L2-resident operands against 4-5 KB of shared memory per state) ----
    auto build_At = [&]() {
        const double *Jb = w.Jbar + s * (long long)CMAX * n;
        const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
        for (int e = tid; e < kc * m; e += nt) {
            const int k = e / m, j = e - k * m;
            const ContactCfg &cc = cfg->contacts[w.cidx[s * MAX_ACTIVE_CONTACTS + k]];
            const double *col = Jb + 6 * k + (long long)(6 + j) * CMAX;
            double o[10];
            contact_rows(cc, pose + cc.link * POSE_STRIDE, ld3(col), ld3(col + 3), o);
            for (int a = 0; a < 10; a++) At[10 * k + a + j * ldc] = o[a];
        }
    };
    // ---- fast path test: with x = 0, is every row of the QP strictly satisfied for the torque `base` (shared memory, visible)?
    // rows: tlim -/+ base > 0 and -(A_const A_rot (P_C - Jbar^T[:,6:] base)) > 0 (contact_rows is linear in the wrench).  Then x = 0
    // is the optimum of both lexicographic stages with an empty active set and none of U, Atemp, Atemp U, the row norms or the
    // factorisation is needed: most QPs of the workload end here (profiles: with f* = 0 the assembly alone cost 7.4 of 9.6 ms).
    // Stricter than solve_level's own quick exit (margin 1e-8 (1 + |b|)), so the results are the same.  Uniform over the CTA.
    // It also leaves the right-hand sides b of the QP (what the rows evaluate to at x = 0) for the solver.
    auto quick_feasible = [&]() -> bool {
        const double *Jb6 = w.Jbar + s * (long long)CMAX * n + 6 * CMAX;
        const double *pc = w.PC + s * CMAX;
        const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
        if (tid == 0) S.ired[0] = 1;
        tgemv<false>(c, m, Jb6, CMAX, base, [&](int i, double v) { lf[i] = pc[i] - v; });
        tsync();
        bool okq = true;
        if (mt)
            for (int i = tid; i < m; i += nt) {
                const double lo = cfg->tlim[i] - base[i], hi = cfg->tlim[i] + base[i];
                b[i] = lo;
                b[m + i] = hi;
                okq = okq && lo > 1e-8 * (1.0 + fabs(lo)) && hi > 1e-8 * (1.0 + fabs(hi));
            }
        for (int k = tid; k < kc; k += nt) {
            const ContactCfg &cc = cfg->contacts[w.cidx[s * MAX_ACTIVE_CONTACTS + k]];
            double o[10];
            contact_rows(cc, pose + cc.link * POSE_STRIDE, ld3(lf + 6 * k), ld3(lf + 6 * k + 3), o);
            for (int a = 0; a < 10; a++) {
                b[mt + 10 * k + a] = -o[a];
                okq = okq && (-o[a] > 1e-8 * (1.0 + fabs(o[a])));
            }
        }
        if (!okq) S.ired[0] = 0;
        tsync();
        const bool res = S.ired[0] != 0;
        tsync();
        return res;
    };
    for (int i = tid; i < m; i += nt) tprev[i] = w.tau_grav[s * m + i];
    tsync();
    const double *NwJw = w.NwJw + s * (long long)m * RMAX;
    // The hierarchy levels and the contact redistribution are the same QP (the redistribution has no task variables and its
    // own tau_in), so they share ONE copy of the assembly / solve / update code: index q < nl is a level, q == nl the
    // redistribution.  (Two inlined copies made this kernel 9.7 k SASS instructions and instruction fetch its top stall.)
    const int nl = hqp ? cfg->nlevels : 0;
    if (hqp) {
        for (int i = tid; i < m; i += nt) tau_task[i] = 0.0, tau_contact[i] = 0.0;
        for (int l = 0; l < nl; l++) {
            const int t = cfg->tasks[l].dof;
            for (int i = tid; i < t; i += nt) w.fstar_qp[l][s * t + i] = 0.0;
            for (int i = tid; i < RMAX; i += nt) w.contact_qp[l][s * RMAX + i] = 0.0;
            if (tid < ACTIVE_WORDS) w.active[l][s * ACTIVE_WORDS + tid] = 0u;
        }
    }
    if (redist) {
        if (tid < ACTIVE_WORDS) w.active[MAX_LEVELS][s * ACTIVE_WORDS + tid] = 0u;
        for (int i = tid; i < RMAX; i += nt) w.cf_redis[s * RMAX + i] = 0.0;
    }
    bool all_ok = task_ok;
    const int TC = nvmax - r;  // first contact column of U / AU
    bool cols_ready = false;   // the contact columns are assembled
    int aq = -1;               // active rows of the kept factorisation of the contact sub-problem (-1: none)
    DWBC_UNROLL(1)
    for (int q = 0; q < nl + (redist ? 1 : 0); q++) {
        const bool lev = q < nl;
        if (lev && !all_ok) continue;  // a failed level ends CalcTaskControlTorque (reference src/dwbc.cpp:836-837); the redistribution still runs
        const int t = lev ? cfg->tasks[q].dof : 0;
        const int nv = t + r;
        tsync();  // the previous QP is done with U / AU / b / fs; tau_task / tau_contact are visible to every thread
        if (!lev && r <= 0) {
            for (int i = tid; i < m; i += nt) tau_contact[i] = 0.0;  // reference src/dwbc.cpp:1562-1566
            st |= 4;
            continue;
        }
        const double *Ntq = lev ? w.Ntq[q] + s * (long long)m * t : NwJw;
        if (lev) {
            // base = tau_prev + Ntorque f*
            for (int i = tid; i < t; i += nt) fs[i] = w.fstar[q][s * t + i];
            tsync();
            tgemv<false>(m, t, Ntq, m, fs, [&](int i, double v) { base[i] = tprev[i] + v; });
        } else {
            // tau_in: torque_grav_ + torque_task_ + torque_contact_ (reference src/dwbc.cpp:1372-1380); the reduced variant passes
            // torque_grav_R_ + torque_task_R_ and ASSIGNS torque_contact_R_ = NwJw_R x (redist == 2, src/dwbc.cpp:3763, 4915)
            for (int i = tid; i < m; i += nt) {
                if (redist == 2) tau_contact[i] = 0.0;
                base[i] = w.tau_grav[s * m + i] + tau_task[i] + tau_contact[i];
            }
        }
        tsync();
        bool solved = true, have_x = false, xt_zero = true;
        lap(21);
        const bool qf = quick_feasible();
        lap(22);
        if (!qf) {
            if (tm && tid == 0) tm[29] += 1;
            uint32_t *aout = w.active[lev ? q : MAX_LEVELS] + s * ACTIVE_WORDS;
            LevelOut lo{0, 0, 0};
            if (r > 0) {
                // the contact columns [NwJw] / Atemp NwJw sit at the fixed columns [TC, TC + r) and are assembled once per state
                if (!cols_ready) {
                    tcopy(m, r, NwJw, m, U + TC * ldm, ldm);
                    build_At();
                    tsync();
                    tgemm_mma<2, false, false>(nc, r, m, At, ldc, U + TC * ldm, ldm, [&](int i, int j, double v) { AU[i + (TC + j) * ldc] = v; });
                    tsync();
                    cols_ready = true;
                    aq = -1;
                    lap(23);
                }
                lo = solve_contact_only<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, t, r, mt, nc, aout, TC, &aq, lev ? 1 : 0, lev ? Ntq : nullptr);
                total_iters += lo.iters;
                lap(24);
            }
            if (!lo.status && lev) {
                // x_t = 0 is not attainable (or there is no internal-force space): the whole problem, task columns [TC - t, TC)
                tcopy(m, t, Ntq, m, U + (TC - t) * ldm, ldm);
                build_At();
                tsync();
                tgemm_mma<2, false, false>(nc, t, m, At, ldc, U + (TC - t) * ldm, ldm, [&](int i, int j, double v) { AU[i + (TC - t + j) * ldc] = v; });
                tsync();
                lap(25);
                xt_zero = false;
                aq = -1;  // the solver reuses the storage of the kept factorisation
                // (Solving on the contact rows alone first and verifying the torque rows afterwards -- they are never active in the headline
                // workload -- was measured slower: 7.75 vs 7.11 ms per 65 536 states; the row scan is not what bounds this kernel.)
                lo = solve_level<CN, CM, CNV, CMT, CC>(sm, n, m, nvmax, t, r, mt, nc, aout, 0, 0, TC - t, r > 0 ? 1 : 0);
                total_iters += lo.iters;
                if (lo.slow) st |= 16;
                lap(26);
            }
            solved = lo.status != 0;
            have_x = solved;
        }
        if (!solved) {
            if (lev) {
                all_ok = false;  // reference: f_star_qp_ = contact_qp_ = 0 and return 0 (src/dwbc.cpp:1117-1125, 836-837)
            } else {
                for (int i = tid; i < m; i += nt) tau_contact[i] = 0.0;  // reference src/dwbc.cpp:1553-1559
            }
            continue;
        }
        // x = [x_t; x_c]  (zero on the fast path: the outputs were cleared above)
        if (tid < nv) {
            const double xj = have_x ? S.xj[tid] : 0.0;
            if (tid < t) {
                if (have_x) w.fstar_qp[q][s * t + tid] = xj;
                fs[tid] += xj;
            } else {
                xs[tid - t] = xj;
                if (have_x) (lev ? w.contact_qp[q] : w.cf_redis)[s * RMAX + tid - t] = xj;
            }
        }
        tsync();
        if (lev) {
            // torque_task += Null_{l-1} J_kt Lambda (f* + f*_qp) = Ntorque (f* + x_t);  tau_prev = tau_grav + torque_task.
            // With x_t = 0 (the x = 0 test or the contact-only solve decided the level) Ntorque f* is base - tau_prev: no second pass
            // over Ntorque in global memory.
#ifdef DWBC_HQP_NO_XTZERO
            xt_zero = false;
#endif
            if (xt_zero) {
                for (int i = tid; i < m; i += nt) {
                    const double tt = tau_task[i] + (base[i] - tprev[i]);
                    tau_task[i] = tt;
                    tprev[i] = w.tau_grav[s * m + i] + tt;
                }
            } else {
                tgemv<false>(m, t, Ntq, m, fs, [&](int i, double v) {
                    const double tt = tau_task[i] + v;
                    tau_task[i] = tt;
                    tprev[i] = w.tau_grav[s * m + i] + tt;
                });
            }
            // torque_h = J_kt Lambda_t (f* + f*_qp)   (a public field: not produced when only the torques were asked for)
            if (w.torque_h[q]) {
                const double *Lt = w.Lamt[q] + s * (long long)t * t, *Jk = w.Jkt[q] + s * (long long)m * t;
                tgemv<false>(t, t, Lt, t, fs, [&](int i, double v) { lf[i] = v; });
                tsync();
                tgemv<false>(m, t, Jk, m, lf, [&](int i, double v) { w.torque_h[q][s * m + i] = v; });
            }
        }
        // level: torque_contact = NwJw contact_qp (zero without an internal-force space: quirk 1 of SURVEY Appendix A);
        // redistribution: torque_contact += NwJw x
        if (r > 0) {
            if (have_x) tgemv<false>(m, r, NwJw, m, xs, [&](int i, double v) { tau_contact[i] = (lev ? 0.0 : tau_contact[i]) + v; });
            else if (lev)
                for (int i = tid; i < m; i += nt) tau_contact[i] = 0.0;
        }
        if (!lev) st |= 4;
    }
    if (hqp && all_ok) st |= 2;
    lap(21);
    if (tid == 0) {
        w.status[s] = st;
        w.qp_iters[s] = total_iters;
    }
}

}  // namespace dwbc
