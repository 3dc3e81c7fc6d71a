// Stages a6-a8: the prioritised torque QPs (one per hierarchy level) and the contact-force
// redistribution QP, solved by a batched dense FP64 active-set method -- one warp per robot state.
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
// Warp mapping: lane j owns variable j (row j of Q) and logical active-set position j (act, multiplier, R column slot).
#pragma once
#include "batch.h"
#include "kernels_rbd.cuh"
#include "team_linalg.cuh"
#include "warp_linalg.cuh"

namespace dwbc {

constexpr int QP_NV_MAX = 32;  // one lane per variable
constexpr int QP_NC_MAX = 10 * MAX_ACTIVE_CONTACTS;

struct HqpSmem {
    int ldm, ldq, ldc, U, AU, At, Q, R, b, tol, rin, bt, gi, vec, total;
    __host__ __device__ HqpSmem(int n, int m) {
        ldm = m | 1;
        ldq = QP_NV_MAX - 1;  // 31, odd
        ldc = QP_NC_MAX | 1;  // 21
        int o = 0;
        U = o; o += ldm * QP_NV_MAX;
        AU = o; o += ldc * QP_NV_MAX;
        At = o; o += ldc * m;
        Q = o; o += ldq * QP_NV_MAX;
        R = o; o += ldq * QP_NV_MAX;
        const int mc = 2 * m + QP_NC_MAX;
        b = o; o += mc + 2;
        tol = o; o += mc + 2;
        rin = o; o += mc + 2;   // 1 / scaled row norm
        bt = o; o += mc + 2;    // right-hand side of the current sub-problem
        gi = o; o += 6 * QP_NV_MAX + 8;  // GiScratch
        vec = o; o += 6 * (n + 2) + 2;  // + status / iteration words of the QP just solved
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
};

// per-warp scratch of the active-set solver (shared memory)
struct GiScratch {
    double *np, *dv, *hv, *dxs, *isds, *xs;  // QP_NV_MAX doubles each
    uint32_t *pinw;                           // ACTIVE_WORDS words
    double *bt;                               // rows doubles
    __device__ __forceinline__ GiScratch(double *base, double *bt_) {
        np = base; dv = base + QP_NV_MAX; hv = base + 2 * QP_NV_MAX; dxs = base + 3 * QP_NV_MAX; isds = base + 4 * QP_NV_MAX;
        xs = base + 5 * QP_NV_MAX;
        pinw = (uint32_t *)(base + 6 * QP_NV_MAX);
        bt = bt_;
    }
};

struct GiResult {
    int status;  // 1 solved, 0 infeasible / iteration limit
    int iters;
};

__device__ __forceinline__ bool bit_test(const uint32_t *w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }

// Goldfarb-Idnani on  min 1/2|y|^2  s.t.  sum_j a(i,j) dx_j <= bt_i,  dx_j = isd_j y_j  (dx = x - x0).
// Lane j returns dx_j.  `pin` (may be null): rows to be pinned as equalities, added first and never dropped.
// On return act_out / u_out hold, for logical position = lane, the active row (-1 if none) and its multiplier.
// Deliberately NOT inlined: the solver is called from several places of solve_level, itself called per level and for the
// redistribution QP; inlined, the kernel grew to 54k SASS instructions and the working warps spent 59 % of their
// non-barrier stall samples waiting for instruction fetch (ncu, profiles/r01_*).
__device__ __noinline__ GiResult gi_solve(const QpView &V, const double *bt, const double *tolv, double *rin, double isd, double *Q,
                                             double *R, int ldq, const GiScratch &S, const uint32_t *pin, double &dx_out, int &act_out,
                                             double &u_out, int max_iter) {
    const int lane = lane_id();
    const int nv = V.nv, mc = V.rows();
    double *np = S.np, *dv = S.dv, *hv = S.hv, *dxs = S.dxs, *isds = S.isds;
    if (lane < nv) isds[lane] = isd;
    for (int e = lane; e < ldq * nv; e += 32) Q[e] = 0.0;
    __syncwarp();
    if (lane < nv) Q[lane + lane * ldq] = 1.0;
    // 1 / (scaled row norm): violations are compared as distances in y-space.  Rows whose norm is below 1e-10 of the
    // largest row norm are numerically zero (what rounding leaves of a constraint already active at a higher level) and
    // never enter the active set (rin = 0); the oracle applies the same rule.
    double smax = 0.0;
    for (int i = lane; i < mc; i += 32) {
        double s = 0.0;
        for (int j = 0; j < nv; j++) {
            const double v = V.a(i, j) * isds[j];
            s = fma(v, v, s);
        }
        rin[i] = s;
        smax = s > smax ? s : smax;
    }
    smax = warp_max(smax);
    for (int i = lane; i < mc; i += 32) {
        const double s = rin[i];
        rin[i] = (s > 0.0 && s > 1e-20 * smax) ? 1.0 / sqrt(s) : 0.0;
    }
    __syncwarp();
    double y = 0.0;           // lane j: scaled variable j
    int act = -1, slot = -1;  // lane k: row and physical R column of logical active position k
    double u = 0.0;           // lane k: multiplier of logical position k
    int q = 0, neq = 0;
    uint32_t used = 0;                            // physical R columns in use (warp-uniform)
    uint32_t amask[ACTIVE_WORDS] = {0, 0, 0, 0};  // active rows (warp-uniform)
    GiResult res{0, 0};
    int p = -1;  // row being added
    double sp = 0.0, sigma = 1.0, up = 0.0;
    int pin_scan = 0;
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
            if (lane < nv) dxs[lane] = y * isd;
            __syncwarp();
            if (found >= 0) {
                double part = lane < nv ? V.a(found, lane) * dxs[lane] : 0.0;
                part = warp_sum(part);
                const double sl = bt[found] - part;
                p = found;
                sigma = sl <= 0.0 ? 1.0 : -1.0;  // reach the equality from whichever side we are on
                sp = -fabs(sl);
            } else {
                double best = 0.0;
                int bi = -1;
                if (V.mt > 0) {
                    for (int i = lane; i < V.m; i += 32) {
                        double s = 0.0;
                        for (int j = 0; j < nv; j++) s = fma(V.U[i + (V.col0 + j) * V.ldm], dxs[j], s);
                        const int i2 = i + V.m;
                        const double s1 = bt[i] - s, s2 = bt[i2] + s;
                        if (!bit_test(amask, i) && s1 < -tolv[i] && s1 * rin[i] < best) best = s1 * rin[i], bi = i;
                        if (!bit_test(amask, i2) && s2 < -tolv[i2] && s2 * rin[i2] < best) best = s2 * rin[i2], bi = i2;
                    }
                }
                for (int i = lane; i < V.nc; i += 32) {
                    double s = 0.0;
                    for (int j = 0; j < nv; j++) s = fma(V.AU[i + (V.col0 + j) * V.ldc], dxs[j], s);
                    const int row = V.mt + i;
                    const double s1 = bt[row] + s;
                    if (!bit_test(amask, row) && s1 < -tolv[row] && s1 * rin[row] < best) best = s1 * rin[row], bi = row;
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ob = __shfl_xor_sync(DWBC_FULL, best, o);
                    const int oi = __shfl_xor_sync(DWBC_FULL, bi, o);
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
        // ---- step direction: n+ = -sigma a~_p, d = Q^T n+, z = Q_2 d_2, r = R^{-1} d_1 ----
        __syncwarp();
        if (lane < nv) np[lane] = -sigma * V.a(p, lane) * isd;
        __syncwarp();
        double dk = 0.0;
        if (lane < nv) {
            for (int j = 0; j < nv; j++) dk = fma(Q[j + lane * ldq], np[j], dk);
            dv[lane] = dk;
        }
        __syncwarp();
        double z = 0.0;
        if (lane < nv)
            for (int k = q; k < nv; k++) z = fma(Q[lane + k * ldq], dv[k], z);
        const double zz = warp_sum((lane >= q && lane < nv) ? dk * dk : 0.0);
        const double nn = warp_sum(lane < nv ? dk * dk : 0.0);
        double rk = (lane < q) ? dk : 0.0;
        for (int k = q - 1; k >= 0; k--) {
            const int sk = __shfl_sync(DWBC_FULL, slot, k);
            const double val = __shfl_sync(DWBC_FULL, rk, k) / R[k + sk * ldq];
            if (lane == k) rk = val;
            if (lane < k) rk = fma(-R[lane + sk * ldq], val, rk);
        }
        const bool dependent = (q >= nv) || !(zz > 1e-22 * nn);
        // largest dual step that keeps the multipliers of the droppable active rows non-negative
        double t1 = 1e300;
        int l1 = -1;
        if (lane >= neq && lane < q && rk > 0.0) t1 = u / rk, l1 = lane;
        for (int o = 16; o > 0; o >>= 1) {
            const double ot = __shfl_xor_sync(DWBC_FULL, t1, o);
            const int ol = __shfl_xor_sync(DWBC_FULL, l1, o);
            if (ol >= 0 && (l1 < 0 || ot < t1 || (ot == t1 && ol < l1))) t1 = ot, l1 = ol;
        }
        if (dependent && l1 < 0) {
            // no primal and no dual step possible
            const bool is_pin = pin && bit_test(pin, p);
            if (is_pin || -sp <= tolv[p]) {
                // a pinned row that depends on the rows pinned so far (consistent by construction), or a row
                // that is violated by less than its tolerance: nothing to do for it
                p = -1;
                continue;
            }
            res.status = 0;  // infeasible
            break;
        }
        const double t2 = dependent ? 1e300 : -sp / zz;
        const double t = t1 < t2 ? t1 : t2;
        if (lane < q) u = fma(-t, rk, u);
        up += t;
        if (!dependent) y = fma(t, z, y);
        if (!dependent && t2 <= t1) {
            // ---- full step: row p becomes active.  One Householder reflection on Q_2 maps d_2 to alpha e_q ----
            const double dq = dv[q];
            const double alpha = dq > 0.0 ? -sqrt(zz) : sqrt(zz);
            if (lane >= q && lane < nv) hv[lane] = (lane == q) ? dk - alpha : dk;
            __syncwarp();
            const double tau = 1.0 / (zz - alpha * dq);
            if (lane < nv) {
                double wj = 0.0;
                for (int k = q; k < nv; k++) wj = fma(Q[lane + k * ldq], hv[k], wj);
                wj *= tau;
                for (int k = q; k < nv; k++) Q[lane + k * ldq] = fma(-wj, hv[k], Q[lane + k * ldq]);
            }
            const int ns = __ffs((int)~used) - 1;
            used |= 1u << ns;
            if (lane < q) R[lane + ns * ldq] = dk;
            if (lane == q) {
                R[q + ns * ldq] = alpha;
                act = p;
                u = up;
                slot = ns;
            }
            amask[p >> 5] |= 1u << (p & 31);
            if (pin && bit_test(pin, p)) neq = q + 1;
            q++;
            p = -1;
            __syncwarp();
        } else {
            // ---- partial (or purely dual) step: the blocking row at logical position l1 leaves the active set ----
            const int l = l1;
            const int dropped = __shfl_sync(DWBC_FULL, act, l);
            const int dslot = __shfl_sync(DWBC_FULL, slot, l);
            for (int k = l; k + 1 < q; k++) {
                // Givens rotation of rows k, k+1 of R: the column moving into position k becomes upper triangular
                const int sk1 = __shfl_sync(DWBC_FULL, slot, k + 1);
                const double a = R[k + sk1 * ldq], bb = R[k + 1 + sk1 * ldq];
                const double h = sqrt(a * a + bb * bb);
                const double cs = h > 0.0 ? a / h : 1.0, sn = h > 0.0 ? bb / h : 0.0;
                __syncwarp();
                if (lane > k && lane < q) {
                    double *col = R + slot * ldq;
                    const double r0 = col[k], r1 = col[k + 1];
                    col[k] = cs * r0 + sn * r1;
                    col[k + 1] = -sn * r0 + cs * r1;
                }
                if (lane < nv) {
                    const double q0 = Q[lane + k * ldq], q1 = Q[lane + (k + 1) * ldq];
                    Q[lane + k * ldq] = cs * q0 + sn * q1;
                    Q[lane + (k + 1) * ldq] = -sn * q0 + cs * q1;
                }
                __syncwarp();
            }
            {
                const int a2 = __shfl_down_sync(DWBC_FULL, act, 1);
                const int s2 = __shfl_down_sync(DWBC_FULL, slot, 1);
                const double u2 = __shfl_down_sync(DWBC_FULL, u, 1);
                if (lane >= l && lane < q - 1) act = a2, slot = s2, u = u2;
                if (lane == q - 1) act = -1, slot = -1, u = 0.0;
            }
            used &= ~(1u << dslot);
            amask[dropped >> 5] &= ~(1u << (dropped & 31));
            q--;
            // slack of row p at the new point
            if (lane < nv) dxs[lane] = y * isd;
            __syncwarp();
            double part = lane < nv ? V.a(p, lane) * dxs[lane] : 0.0;
            part = warp_sum(part);
            sp = sigma * (bt[p] - part);
            if (sp > 0.0) sp = 0.0;
        }
    }
    dx_out = lane < nv ? y * isd : 0.0;
    act_out = lane < q ? act : -1;
    u_out = lane < q ? u : 0.0;
    __syncwarp();
    return res;
}

struct LevelOut {
    int status;
    int iters;
};

// Solves one QP (t task variables first, then r contact variables).  U (m x nv), AU (nc x nv) and b (rows) are in
// shared memory; lane j returns x_j.  active_out (global, ACTIVE_WORDS words): rows with zero slack at the solution.
__device__ __noinline__ LevelOut solve_level(int t, int r, int m, int mt, int nc, const double *U, const double *AU, int ldm, int ldc,
                                                const double *b, double *tolv, double *rin, double *Q, double *R, int ldq,
                                                const GiScratch &S, double &x_out, uint32_t *active_out) {
    const int lane = lane_id();
    const int nv = t + r, mc = mt + nc;
    LevelOut out{1, 0};
    QpView V{U, AU, ldm, ldc, m, mt, nc, 0, nv};
    // row tolerances in the units of the rows, the quick exit, and the active rows of the quick exit.  Numerically zero
    // rows (norm <= 1e-10 of the largest) cannot be influenced by x: they only decide feasibility (b >= -1e-9).
    double smax = 0.0;
    for (int i = lane; i < mc; i += 32) {
        double s = 0.0;
        for (int j = 0; j < nv; j++) {
            const double v = V.a(i, j);
            s = fma(v, v, s);
        }
        rin[i] = s;
        smax = s > smax ? s : smax;
    }
    smax = warp_max(smax);
    double bmin = 1e300;
    bool infeasible = false;
    for (int i0 = 0; i0 < mc; i0 += 32) {
        const int i = i0 + lane;
        bool on = false;
        if (i < mc) {
            const double s = rin[i];
            const double sc = fmax(sqrt(s), fabs(b[i]));
            tolv[i] = 1e-10 * sc;
            if (s > 1e-20 * smax && s >= 1e-28)
                bmin = fmin(bmin, b[i] + 1e-10 * sc);
            else if (b[i] < -1e-9)
                infeasible = true;
            on = s > 0.0 && b[i] < 1e-9 * (1.0 + fabs(b[i]));
        }
        const unsigned bal = __ballot_sync(DWBC_FULL, on);
        if (lane == 0) S.pinw[i0 >> 5] = bal;
    }
    if (__any_sync(DWBC_FULL, infeasible)) {
        out.status = 0;
        x_out = 0.0;
        __syncwarp();
        return out;
    }
    bmin = warp_min(bmin);
    __syncwarp();
    double xj = 0.0;
    if (bmin >= 0.0) {
        // x = 0 satisfies every row: optimal for both stages, no factorisation needed
        if (active_out && lane < ACTIVE_WORDS) active_out[lane] = lane < (mc + 31) / 32 ? S.pinw[lane] : 0u;
        x_out = 0.0;
        __syncwarp();
        return out;
    }
    const int max_iter = 20 * (nv + 8);
    int act = -1;
    double u = 0.0, dx = 0.0;
    bool done = false;
    if (t == 0 || r == 0) {
        GiResult g = gi_solve(V, b, tolv, rin, 1.0, Q, R, ldq, S, nullptr, dx, act, u, max_iter);
        out.iters += g.iters;
        out.status = g.status;
        xj = dx;
        done = true;
    }
    if (!done) {
        // ---- A: is x_t = 0 attainable?  (looser tolerances: the set may be a thin face inherited from the levels above)
        for (int i = lane; i < mc; i += 32) tolv[i] *= 100.0;
        __syncwarp();
        QpView Vc = V;
        Vc.col0 = t;
        Vc.nv = r;
        GiResult g = gi_solve(Vc, b, tolv, rin, 1.0, Q, R, ldq, S, nullptr, dx, act, u, max_iter);
        out.iters += g.iters;
        if (g.status) {
            const double v = __shfl_sync(DWBC_FULL, dx, lane >= t ? lane - t : 0);
            xj = (lane >= t && lane < nv) ? v : 0.0;
            done = true;
        }
        for (int i = lane; i < mc; i += 32) tolv[i] *= 0.01;
        __syncwarp();
    }
    if (!done) {
        // ---- B: proximal point on 1/2|x_t|^2 + delta/2 |x_c - c|^2 ----
        double nt = 0.0, ncn = 0.0;
        for (int i = lane; i < mc; i += 32)
            for (int j = 0; j < nv; j++) {
                const double v = V.a(i, j);
                if (j < t) nt = fma(v, v, nt);
                else ncn = fma(v, v, ncn);
            }
        nt = warp_sum(nt);
        ncn = warp_sum(ncn);
        const double delta = 1e-5 * (ncn / r) / fmax(nt / t, 1e-300);
        const double isd = lane < t ? 1.0 : rsqrt(delta);
        double *bt = S.bt, *xs = S.xs;
        double x0 = 0.0, xprev = 0.0;
        bool conv = false;
        for (int k = 0; k < 30; k++) {
            if (lane < nv) xs[lane] = x0;
            __syncwarp();
            for (int i = lane; i < mc; i += 32) {  // bt = b - A x0 (x0 is zero on the task columns)
                double s = b[i];
                for (int j = t; j < nv; j++) s = fma(-V.a(i, j), xs[j], s);
                bt[i] = s;
            }
            __syncwarp();
            GiResult g = gi_solve(V, bt, tolv, rin, isd, Q, R, ldq, S, nullptr, dx, act, u, max_iter);
            out.iters += g.iters;
            if (!g.status) {
                out.status = 0;
                break;
            }
            const double xn = x0 + dx;
            const double dd = warp_sum(lane < t ? (xn - xprev) * (xn - xprev) : 0.0);
            const double xx = warp_sum(lane < t ? xn * xn : 0.0);
            xprev = xn;
            xj = xn;
            x0 = (lane >= t && lane < nv) ? xn : 0.0;
            if (k > 0 && sqrt(dd) <= 1e-12 * (1.0 + sqrt(xx))) {
                conv = true;
                break;
            }
        }
        if (out.status && !conv) out.status = 0;
        if (out.status) {
            // ---- C: min |x_c|^2 on the optimal face; rows with a positive multiplier are pinned ----
            const double umax = warp_max(u);
            if (lane < ACTIVE_WORDS) S.pinw[lane] = 0u;
            __syncwarp();
            const int mine = (act >= 0 && u > 1e-9 * umax) ? act : -1;
            for (int k = 0; k < nv; k++) {
                const int ak = __shfl_sync(DWBC_FULL, mine, k);
                if (ak >= 0 && lane == 0) S.pinw[ak >> 5] |= 1u << (ak & 31);
            }
            if (lane < nv) xs[lane] = xj;
            __syncwarp();
            for (int i = lane; i < mc; i += 32) {
                double s = b[i];
                for (int j = 0; j < t; j++) s = fma(-V.a(i, j), xs[j], s);
                bt[i] = s;
                tolv[i] *= 100.0;
            }
            __syncwarp();
            QpView Vc = V;
            Vc.col0 = t;
            Vc.nv = r;
            GiResult g = gi_solve(Vc, bt, tolv, rin, 1.0, Q, R, ldq, S, S.pinw, dx, act, u, max_iter);
            out.iters += g.iters;
            if (!g.status) out.status = 0;
            const double v = __shfl_sync(DWBC_FULL, dx, lane >= t ? lane - t : 0);
            if (lane >= t && lane < nv) xj = v;
        }
    }
    if (!out.status) xj = 0.0;
    // active rows at the solution: slack within 1e-9 (1 + |b|), rows with a zero normal excluded (the oracle's rule)
    if (lane < nv) S.xs[lane] = xj;
    __syncwarp();
    for (int i0 = 0; i0 < mc; i0 += 32) {
        const int i = i0 + lane;
        bool on = false;
        if (i < mc && out.status) {
            double s = -b[i], nn2 = 0.0;
            for (int j = 0; j < nv; j++) {
                const double a = V.a(i, j);
                s = fma(a, S.xs[j], s);
                nn2 = fma(a, a, nn2);
            }
            on = nn2 > 0.0 && s > -1e-9 * (1.0 + fabs(b[i]));
        }
        const unsigned bal = __ballot_sync(DWBC_FULL, on);
        if (lane == 0) S.pinw[i0 >> 5] = bal;
    }
    __syncwarp();
    if (active_out && lane < ACTIVE_WORDS) active_out[lane] = lane < (mc + 31) / 32 ? S.pinw[lane] : 0u;
    x_out = xj;
    __syncwarp();
    return out;
}

// One CTA (team) per state: QP assembly (U, Atemp U, right-hand sides) and the torque updates use every thread; the
// active-set iterations of one QP are inherently narrow (one lane per variable) and run on warp 0 while the CTA's other
// warps wait at the barrier -- the SM's other resident CTAs fill the issue slots meanwhile.
__global__ void __launch_bounds__(TEAM, 5) hqp_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, Work w, int hqp, int redist) {
    DWBC_DYN_SMEM(smem);
    const int tid = t_id(), nt = t_n(), warp = tid >> 5, lane = lane_id();
    const long long s = blockIdx.x;
    const int n = w.n, m = w.m, nb = w.nb;
    const HqpSmem L(n, m);
    double *sm = (double *)smem;
    double *U = sm + L.U, *AU = sm + L.AU, *At = sm + L.At, *Q = sm + L.Q, *R = sm + L.R, *b = sm + L.b, *tolv = sm + L.tol, *rin = sm + L.rin;
    const GiScratch S(sm + L.gi, sm + L.bt);
    double *vec = sm + L.vec;
    const int vs = n + 2;
    double *tprev = vec, *base = vec + vs, *fs = vec + 2 * vs, *bA0 = vec + 3 * vs, *xs = vec + 4 * vs, *lf = vec + 5 * vs;
    int *s_lo = (int *)(vec + 6 * vs);  // status, iterations of the QP just solved by warp 0
    const int ldm = L.ldm, ldc = L.ldc, ldq = L.ldq;
    int st = w.status[s];
    const bool contact_ok = st & 1;
    const bool task_ok = (st & 8) != 0;  // set by task_space_kernel when hqp = 1
    st &= ~8;
    const int c = contact_ok ? w.cdof[s] : 0;
    const int r = c > 6 ? c - 6 : 0;
    const int kc = c / 6;
    const int nc = 10 * kc;
    const int mt = cfg->tlim_set ? 2 * m : 0;
    int total_iters = 0;
    double *tau_task = w.tau_task + s * m, *tau_contact = w.tau_contact + s * m;
    if (!contact_ok) {  // uniform over the CTA
        if (tid == 0) w.status[s] = st, w.qp_iters[s] = 0;
        return;
    }
    // ---- Atemp = A_const A_rot Jbar^T[:, 6:]  (nc x m)  and  bA0 = A_const A_rot P_C ----
    {
        const double *Jb = w.Jbar + s * (long long)CMAX * n;
        const double *pose = w.pose + s * (long long)nb * POSE_STRIDE;
        for (int e = tid; e < kc * (m + 1); e += nt) {
            const int k = e / (m + 1), j = e - k * (m + 1);
            const ContactCfg &cc = cfg->contacts[w.cidx[s * MAX_ACTIVE_CONTACTS + k]];
            const double *Rl = pose + cc.link * POSE_STRIDE;
            double o[10];
            if (j < m) {
                const double *col = Jb + 6 * k + (long long)(6 + j) * CMAX;
                contact_rows(cc, Rl, ld3(col), ld3(col + 3), o);
                for (int a = 0; a < 10; a++) At[10 * k + a + j * ldc] = o[a];
            } else {
                const double *pc = w.PC + s * CMAX + 6 * k;
                contact_rows(cc, Rl, ld3(pc), ld3(pc + 3), o);
                for (int a = 0; a < 10; a++) bA0[10 * k + a] = o[a];
            }
        }
    }
    for (int i = tid; i < m; i += nt) tprev[i] = w.tau_grav[s * m + i];
    tsync();
    const double *NwJw = w.NwJw + s * (long long)m * RMAX;
    if (hqp) {
        const int nl = cfg->nlevels;
        for (int i = tid; i < m; i += nt) tau_task[i] = 0.0, tau_contact[i] = 0.0;
        for (int l = 0; l < nl; l++) {
            const int t = cfg->tasks[l].dof;
            for (int i = tid; i < t; i += nt) w.fstar_qp[l][s * t + i] = 0.0;
            for (int i = tid; i < RMAX; i += nt) w.contact_qp[l][s * RMAX + i] = 0.0;
            if (tid < ACTIVE_WORDS) w.active[l][s * ACTIVE_WORDS + tid] = 0u;
        }
        bool all_ok = task_ok;
        for (int l = 0; l < nl && all_ok; l++) {
            const int t = cfg->tasks[l].dof;
            const int nv = t + r;
            tsync();  // the previous level is done with U / AU / b / fs
            // U = [Ntorque | NwJw]
            tcopy(m, t, w.Ntq[l] + s * (long long)m * t, m, U, ldm);
            tcopy(m, r, NwJw, m, U + t * ldm, ldm);
            for (int i = tid; i < t; i += nt) fs[i] = w.fstar[l][s * t + i];
            tsync();
            // base = tau_prev + Ntorque f*
            tgemv<false>(m, t, U, ldm, fs, [&](int i, double v) { base[i] = tprev[i] + v; });
            // AU = Atemp U
            tgemm<2, 2, false, false>(nc, nv, m, At, ldc, U, ldm, [&](int i, int j, double v) { AU[i + j * ldc] = v; });
            tsync();
            if (mt)
                for (int i = tid; i < m; i += nt) {
                    b[i] = cfg->tlim[i] - base[i];
                    b[m + i] = cfg->tlim[i] + base[i];
                }
            tgemv<false>(nc, m, At, ldc, base, [&](int i, double v) { b[mt + i] = -(bA0[i] - v); });
            tsync();
            if (warp == 0) {
                double xj;
                LevelOut lo = solve_level(t, r, m, mt, nc, U, AU, ldm, ldc, b, tolv, rin, Q, R, ldq, S, xj, w.active[l] + s * ACTIVE_WORDS);
                if (lane == 0) s_lo[0] = lo.status, s_lo[1] = lo.iters;
                if (lo.status) {
                    if (lane < t) {
                        w.fstar_qp[l][s * t + lane] = xj;
                        fs[lane] += xj;
                    }
                    if (lane >= t && lane < nv) {
                        w.contact_qp[l][s * RMAX + lane - t] = xj;
                        xs[lane - t] = xj;
                    }
                }
            }
            tsync();
            total_iters += s_lo[1];
            if (!s_lo[0]) {
                all_ok = false;  // reference: f_star_qp_ = contact_qp_ = 0 and return 0 (src/dwbc.cpp:1117-1125, 836-837)
                break;
            }
            // torque_task += Null_{l-1} J_kt Lambda (f* + f*_qp) = Ntorque (f* + x_t);  tau_prev = tau_grav + torque_task
            tgemv<false>(m, t, U, ldm, fs, [&](int i, double v) {
                const double tt = tau_task[i] + v;
                tau_task[i] = tt;
                tprev[i] = w.tau_grav[s * m + i] + tt;
            });
            // torque_h = J_kt Lambda_t (f* + f*_qp)
            {
                const double *Lt = w.Lamt[l] + s * (long long)t * t, *Jk = w.Jkt[l] + s * (long long)m * t;
                tgemv<false>(t, t, Lt, t, fs, [&](int i, double v) { lf[i] = v; });
                tsync();
                tgemv<false>(m, t, Jk, m, lf, [&](int i, double v) { w.torque_h[l][s * m + i] = v; });
            }
            // torque_contact = NwJw contact_qp  (zero without an internal-force space: quirk 1 of SURVEY Appendix A)
            if (r > 0) tgemv<false>(m, r, NwJw, m, xs, [&](int i, double v) { tau_contact[i] = v; });
        }
        if (all_ok) st |= 2;
    }
    if (redist) {
        if (tid < ACTIVE_WORDS) w.active[MAX_LEVELS][s * ACTIVE_WORDS + tid] = 0u;
        for (int i = tid; i < RMAX; i += nt) w.cf_redis[s * RMAX + i] = 0.0;
        tsync();  // tau_task / tau_contact of the level loop are visible to every thread
        if (r <= 0) {
            for (int i = tid; i < m; i += nt) tau_contact[i] = 0.0;  // reference src/dwbc.cpp:1562-1566
            st |= 4;
        } else {
            for (int i = tid; i < m; i += nt) base[i] = w.tau_grav[s * m + i] + tau_task[i] + tau_contact[i];  // tau_in
            tcopy(m, r, NwJw, m, U, ldm);
            tsync();
            tgemm<2, 1, false, false>(nc, r, m, At, ldc, U, ldm, [&](int i, int j, double v) { AU[i + j * ldc] = v; });
            if (mt)
                for (int i = tid; i < m; i += nt) {
                    b[i] = cfg->tlim[i] - base[i];
                    b[m + i] = cfg->tlim[i] + base[i];
                }
            tgemv<false>(nc, m, At, ldc, base, [&](int i, double v) { b[mt + i] = -(bA0[i] - v); });
            tsync();
            if (warp == 0) {
                double xj;
                LevelOut lo = solve_level(0, r, m, mt, nc, U, AU, ldm, ldc, b, tolv, rin, Q, R, ldq, S, xj, w.active[MAX_LEVELS] + s * ACTIVE_WORDS);
                if (lane == 0) s_lo[0] = lo.status, s_lo[1] = lo.iters;
                if (lo.status && lane < r) {
                    w.cf_redis[s * RMAX + lane] = xj;
                    xs[lane] = xj;
                }
            }
            tsync();
            total_iters += s_lo[1];
            if (s_lo[0]) {
                tgemv<false>(m, r, NwJw, m, xs, [&](int i, double v) { tau_contact[i] += v; });
                st |= 4;
            } else {
                for (int i = tid; i < m; i += nt) tau_contact[i] = 0.0;  // reference src/dwbc.cpp:1553-1559
            }
        }
    }
    if (tid == 0) {
        w.status[s] = st;
        w.qp_iters[s] = total_iters;
    }
}

}  // namespace dwbc
