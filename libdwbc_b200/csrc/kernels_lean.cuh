// The whole control cycle of one robot state in ONE kernel, specialised at compile time for a model size and a task hierarchy:
// stages a1-a5 (rigid-body dynamics, A^{-1}, contact Jacobians, contact-consistent dynamics, gravity compensation, task-space
// quantities) plus the x = 0 test of every level QP and of the redistribution QP (a6 / a7).  A state whose QPs are all inactive is
// finished here; the others are appended to a compact list and finished by hqp_kernel (kernels_hqp.cuh) launched over that list.
//
// Reference: RobotData::UpdateKinematics (src/dwbc.cpp:279-371), SetContact / ContactConstraint::Update (include/dwbc.h:432-474,
// src/contact_constraint.cpp:51-77), CalculateContactConstraint + PinvCODWB (src/wbd.cpp:108-143, 32-53),
// CalculateGravityCompensation (src/wbd.cpp:186-192), CalcTaskSpace / CalculateJKT / CalculateTaskNullSpace (src/dwbc.cpp:685-816,
// src/wbd.cpp:207-213, 257-261), CalcTaskControlTorque (src/dwbc.cpp:818-873), the quick exit of CalcSingleTaskTorqueWithQP and
// CalcContactRedistribute (src/dwbc.cpp:941-1127, 1372-1568).
//
// Why one kernel (profiles/r01_traffic.json): the six-kernel pipeline moved 12.5 GB of HBM per 65 536-state step
// (every public RobotData matrix of every state materialised and re-read cold by the next kernel) and spent 90 % of its issue
// slots on index arithmetic of run-time dimensions.  Here nothing but the inputs (q, qdot, f*, flags), the three torque vectors
// and, for QP-active states only, the QP operands touches HBM; all matrix dimensions are template parameters.
//
// Algebra (same results up to rounding; torque-only outputs):
//   A^{-1} = L^{-T} L^{-1}                                     blocked Cholesky / triangular inverse / SYRK (lean_linalg.cuh)
//   X = A^{-1} Jc^T, Lambda_c = (Jc X)^{-1}, Jbar^T = Lambda_c X^T, P = A^{-1} N_c = A^{-1} - X Jbar^T
//   W = P[6:,6:],  W^+ = (W + V V^T)^{-1} - V V^T               V = closed-form orthonormal basis of null(W)
//   T = J P (all levels stacked), Tm = T[:, 6:],  Y = W^+ Tm^T,  M_l = Tm_l Y_l
//   J_kt Lambda_t = W^+ Q^T (Q W^+ Q^T)^{-1} Lambda_t  with  Q = Lambda_t Tm  ==>  J_kt Lambda_t = Y_l M_l^{-1}     (Lambda_t cancels:
//   the torque path never needs Lambda_t = (T J^T)^{-1}; the public fields Lambda_task_ / J_kt_ are produced by the generic kernels)
//   Ntorque_l = Null_{l-1} J_kt Lambda_t = JL_l - Nta[:, 0:r_l] (Tm[0:r_l, :] JL_l)
#pragma once
#include "batch.h"
#include "kernels_contact.cuh"
#include "kernels_hqp.cuh"
#include "kernels_rbd.cuh"
#include "lean_linalg.cuh"

namespace dwbc {

__host__ __device__ constexpr int ld4(int x) {  // smallest leading dimension >= x that is 4 (mod 8): conflict-free DMMA fragment loads
    int v = x;
    while ((v & 7) != 4) v++;
    return v;
}
__host__ __device__ constexpr int cmax2(int a, int b) { return a > b ? a : b; }

template <int... T>
struct LevelPack {
    static constexpr int NL = sizeof...(T);
    static constexpr int TS = (T + ... + 0);
    __host__ __device__ static constexpr int dof(int l) {
        constexpr int d[NL ? NL : 1] = {T...};
        return d[l];
    }
    __host__ __device__ static constexpr int r0(int l) {
        int s = 0;
        for (int k = 0; k < l; k++) s += dof(k);
        return s;
    }
};

template <int NBODY, class LV, int LDNP>
struct LeanSpec {
    using Lv = LV;
    static constexpr int NB = NBODY, N = NBODY + 5, M = NBODY - 1;
    static constexpr int LDN = LDNP, LDM = ld4(M), TS = LV::TS, LDT = ld4(TS > 0 ? TS : 1), LDC = 12;
    static_assert(CMAX == 12 && RMAX == 6, "two 6D contacts");
};

// kernel arguments beyond the workspace
struct LeanArgs {
    const double *q, *qd;
    const uint8_t *flags;
    int hqp;     // 1: CalcTaskControlTorque(hqp = true) -- classify the level QPs; 0: closed form
    int redist;  // 1: CalcContactRedistribute(hqp = true) follows
    int *list;   // [B] compact list of the states hqp_kernel still has to finish
    int *count;  // number of entries of list
    long long *timing;  // diagnostics (may be null): [B][32] clock64() stamps of the phases of every state, see LEAN_STAMP
    const int *plist;   // [<= B] the states this launch owns (those of one contact pattern class), CTA i runs state plist[i]
    const int *pcount;  // number of entries of plist; CTAs beyond it exit
};

// Contact-pattern classes: the kernel is compiled once per contact dof (the two variants are ~20 k SASS instructions each and share
// nothing from the contact stage on).  A batch that mixes double and single support used to run both variants side by side on every
// SM and the instruction fetch collapsed (ncu: no_instruction 1.0 -> 7.6 warps per issue, 10 -> 25 ms per 65 536 states); now the
// states are binned by class first and each class gets its own launch, so an SM streams one variant at a time.
// class 0: exactly two active 6D contacts (c = 12); class 1: everything else (c = 6, or no / too many contacts -> status 0).
__global__ void lean_classify_kernel(const BatchCfg *__restrict__ cfg, const uint8_t *__restrict__ flags, int B, int *lists, int *counts) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= B) return;
    int k = 0;
    for (int i = 0; i < cfg->ncontact; i++) k += flags[(long long)s * cfg->ncontact + i] ? 1 : 0;
    const int cls = k == 2 ? 0 : 1;
    const int slot = atomicAdd(counts + cls, 1);
    lists[(long long)cls * B + slot] = s;
}
// phase stamps (thread 0 of the CTA): 0 start, 1 sweeps, 2 A, 3 A^-1, 4 Jc, 5 X, 6 Lambda_c, 7 Jbar + P, 8 V, 9 J_task, 10 T, 11 W + VV^T,
// 12 (W + VV^T)^-1, 13 W^+ / NwJw / tau_grav, 14 Y, 15 M_l, 16 M_l^-1, 17 JL, 18 Ntorque, 19 QP tests, 20 end
#define LEAN_STAMP(tm, k)                            \
    do {                                             \
        if ((tm) && t_id() == 0) (tm)[k] = clock64(); \
    } while (0)

// states per CTA of lean_cycle_kernel (see the kernel)
#ifndef LEAN_GANG
#ifdef DWBC_CPU_EMUL
#define LEAN_GANG 1
#else
#define LEAN_GANG 1
#endif
#endif

// shared-memory arena of one state (doubles)
template <class S>
struct LeanSmem {
    using LV = typename S::Lv;
    static constexpr int N = S::N, M = S::M, LDN = S::LDN, LDM = S::LDM, LDT = S::LDT, LDC = S::LDC, TS = S::TS;
    // R1: A -> L -> A^{-1} -> P (its [6:,6:] block: W -> W + V V^T -> W^+) -> task-level blocks
    static constexpr int R1 = 0, R1_SZ = LDN * N;
    // R2: motion / force columns of the mass matrix -> L^{-1} -> {Jbar, Jc, X, Lambda_c, scratch} -> J_task -> scratch of the W inverse -> Y -> JL -> Ntorque
    static constexpr int oJbar = 0, oJc = LDC * N, oX = 2 * LDC * N, oLam = oX + LDN * CMAX, oLw = oLam + LDC * CMAX, GRP = oLw + LDC * CMAX;
    static constexpr int R2 = R1 + R1_SZ;
    static constexpr int R2_SZ = cmax2(cmax2(cmax2(LDN * N, GRP), cmax2(LDT * N, LDM * M)), cmax2(LDM * (TS > 0 ? TS : 1), 12 * N));
    // R3: per-body records of the dynamics sweeps (alive until the Jacobians are built) -> T
    static constexpr int R3 = R2 + R2_SZ;
    static constexpr int R3_SZ = cmax2(S::NB * RB_STRIDE + 16, LDT * N);
    // R4: small persistent items
    static constexpr int V = R3 + R3_SZ;             // M x 6, LDM
    static constexpr int Aug = V + LDM * 6;          // Gauss-Jordan scratch 13 x 13
    static constexpr int Jb6 = Aug + 13 * 13 + 1;    // Jbar^T[:, 6:]  (c x M, LDC)
    static constexpr int A06 = Jb6 + LDC * M;        // rows 0..5 of A (6 x N, ld 6)
    static constexpr int NwJw = A06 + 6 * N;         // M x 6, LDM
    static constexpr int Rc = NwJw + LDM * 6;        // rotation of the active contact links, 2 x 9
    static constexpr int vec = Rc + 18;
    static constexpr int rd = vec, Gs = rd + N + 1, u = Gs + N + 1, fs = u + M + 1, tg = fs + (TS | 1) + 1, tt = tg + M + 1, base = tt + M + 1,
                         PC = base + M + 1, lf = PC + CMAX, com = lf + CMAX, misc = com + 16;
    // misc: [0] pivot slot, [1] validity flag (int), [2..5] relative pivots of the task levels, [6] contact dof (int), [7] scratch flag (int), [8] active contact indices (2 int)
    static constexpr int anc = misc + 16;  // ancestor bit masks of the bodies (uint64 each): read in inner loops, fetched from the model once
    static constexpr int total = anc + S::NB + 2;
    // task-level blocks inside R1 (W^+ is dead by then)
    __host__ __device__ static constexpr int ldb(int l) { return ld4(LV::dof(l)); }
    __host__ __device__ static constexpr int mo(int l) {  // offset of M_l; each level keeps 2 blocks (matrix + inverse scratch)
        int o = 0;
        for (int k = 0; k < l; k++) o += 2 * ldb(k) * LV::dof(k);
        return o;
    }
    static constexpr int Cs = mo(LV::NL);  // (r_l x t_l) coupling block of the null-space recursion
    static constexpr int CS_SZ = LV::NL > 1 ? ld4(LV::r0(LV::NL - 1)) * LV::dof(LV::NL - 1) : 0;
    static_assert(Cs + CS_SZ <= R1_SZ, "task blocks do not fit the A region");
};

// 6 x n point Jacobian (rows [linear; angular]) of a point fixed in `body` from the sweep records rb (positions relative to the
// pelvis origin), written to dst (leading dimension ldd) starting at row row0; one column per thread.
__device__ __forceinline__ void point_jacobian_rb(const uint64_t *anc_s, const double *rb, int n, int body, Vec3 pl, double *dst, int ldd,
                                                  int row0, int rows_first, int rows_count, int tid, int nt) {
    const double *pb = rb + body * RB_STRIDE;
    const Vec3 pw = ld3(pb + RB_P) + mv(pb + RB_R, pl);
    const double *R0 = rb + RB_R;
    const uint64_t anc = anc_s[body];
    for (int j = tid; j < n; j += nt) {
        Vec3 lin = v3(0, 0, 0), ang = v3(0, 0, 0);
        if (j < 3) {
            lin = v3(j == 0, j == 1, j == 2);
        } else if (j < 6) {
            const Vec3 a = v3(R0[j - 3], R0[3 + j - 3], R0[6 + j - 3]);
            ang = a;
            lin = cross(a, pw);
        } else if ((anc >> (j - 5)) & 1ull) {
            const double *pj = rb + (j - 5) * RB_STRIDE;
            const Vec3 a = ld3(pj + RB_AX);
            ang = a;
            lin = cross(a, pw - ld3(pj + RB_P));
        }
        const double v[6] = {lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
        for (int r = 0; r < rows_count; r++) dst[row0 + r + j * ldd] = v[rows_first + r];
    }
}


// Forward kinematics, common-frame inertias and composite inertias of one state for the torque path (no velocities, bias
// accelerations or forces: B_ is not an output of DWBC_OUT_TORQUES).  The trigonometry and the joint-local rotations E_T R(q) of ALL
// joints are formed by the team in parallel; only the two dependent chains -- parent-to-child pose, child-to-parent composite
// inertia -- run level by level on one warp (`serial_warp`), 27 + 9 FMAs per body.  Same records as rbd_sweeps (RB_R, RB_P, RB_AX,
// RB_IC).  Synchronises the team.
__device__ __forceinline__ void rbd_sweeps_lean(const DevModel *__restrict__ md, const double *__restrict__ q, double *rb, int serial_warp, uint64_t *anc_s) {
    const int tid = t_id(), nt = TEAM, lane = lane_id(), warp = lean::warp_id();
    const int nb = md->nb, n = md->n;
    for (int b = tid; b < nb; b += nt) {
        double *bb = rb + b * RB_STRIDE;
        anc_s[b] = md->anc_mask[b];
        if (b == 0) {
            const double x = q[3], y = q[4], z = q[5], w = q[n];
            double *R = bb + RB_R;
            R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
            R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
            R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
            bb[RB_P] = bb[RB_P + 1] = bb[RB_P + 2] = 0.0;  // positions are kept relative to the pelvis origin
            bb[RB_AX] = bb[RB_AX + 1] = bb[RB_AX + 2] = 0.0;
        } else {
            const Vec3 a = ld3(md->axis[b]);
            double sn, cs;
            sincos(q[5 + b], &sn, &cs);
            const double c1 = 1.0 - cs;
            const double Rq[9] = {cs + c1 * a.x * a.x,       c1 * a.x * a.y - sn * a.z, c1 * a.x * a.z + sn * a.y,
                                  c1 * a.x * a.y + sn * a.z, cs + c1 * a.y * a.y,       c1 * a.y * a.z - sn * a.x,
                                  c1 * a.x * a.z - sn * a.y, c1 * a.y * a.z + sn * a.x, cs + c1 * a.z * a.z};
            double Rj[9];
            mm3(md->E_T[b], Rq, Rj);
            DWBC_UNROLL(9)
            for (int k = 0; k < 9; k++) bb[RB_R + k] = Rj[k];  // joint-local rotation; becomes the world rotation in the chain below
        }
    }
    // The two dependent chains run on one warp, lane k owning entries k and k + 32 of the level order.  Everything the chains need from
    // the model (parent, depth, joint offset and axis, children) is fetched into registers BEFORE the chains start, so that their
    // steps only touch shared memory: with the loads inside the loops every level paid two or three dependent global-memory latencies.
    int cb_[2] = {-1, -1}, cp_[2] = {0, 0}, cd_[2] = {-1, -1}, cch_[2][4] = {{-1, -1, -1, -1}, {-1, -1, -1, -1}}, cnch_[2] = {0, 0};
    Vec3 crt_[2] = {v3(0, 0, 0), v3(0, 0, 0)}, cax_[2] = {v3(0, 0, 0), v3(0, 0, 0)};
    const int max_depth = md->max_depth;
    if (warp == serial_warp) {
        DWBC_UNROLL(2)
        for (int k = 0; k < 2; k++) {
            const int e = lane + 32 * k;
            if (e < nb) {
                const int b = md->order[e];
                cb_[k] = b;
                int d = 0;
                for (int dd = 1; dd <= max_depth; dd++) d = e >= md->level_start[dd] ? dd : d;
                cd_[k] = d;
                cp_[k] = b > 0 ? md->parent[b] : 0;
                crt_[k] = ld3(md->r_T[b]);
                cax_[k] = ld3(md->axis[b]);
                const int c0 = md->child_start[b], c1 = md->child_start[b + 1];
                cnch_[k] = c1 - c0;
                DWBC_UNROLL(4)
                for (int j = 0; j < 4; j++) cch_[k][j] = c0 + j < c1 ? md->child_list[c0 + j] : -1;
            }
        }
    }
    tsync();
    if (warp == serial_warp) {
        for (int d = 1; d <= max_depth; d++) {
            DWBC_UNROLL(2)
            for (int k = 0; k < 2; k++) {
                if (cd_[k] != d) continue;
                const double *pb = rb + cp_[k] * RB_STRIDE;
                double *bb = rb + cb_[k] * RB_STRIDE;
                double Rw[9];
                mm3(pb + RB_R, bb + RB_R, Rw);
                DWBC_UNROLL(9)
                for (int j = 0; j < 9; j++) bb[RB_R + j] = Rw[j];
                st3(bb + RB_P, ld3(pb + RB_P) + mv(pb + RB_R, crt_[k]));
                st3(bb + RB_AX, mv(Rw, cax_[k]));
            }
            __syncwarp();
        }
    }
    tsync();
    for (int b = tid; b < nb; b += nt) {  // inertia of every body in the common frame (mass, first moment, inertia about the pelvis origin)
        double *bb = rb + b * RB_STRIDE;
        const double *R = bb + RB_R;
        const double mass = md->mass[b];
        const Vec3 c = ld3(bb + RB_P) + mv(R, ld3(md->com[b]));
        double T[9], Iw[9], Rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
        mm3(R, md->inertia[b], T);
        mm3(T, Rt, Iw);
        const double cc = dot(c, c);
        double *ic = bb + RB_IC;
        ic[0] = mass;
        st3(ic + 1, mass * c);
        ic[4] = Iw[0] + mass * (cc - c.x * c.x);
        ic[5] = Iw[1] - mass * c.x * c.y;
        ic[6] = Iw[2] - mass * c.x * c.z;
        ic[7] = Iw[4] + mass * (cc - c.y * c.y);
        ic[8] = Iw[5] - mass * c.y * c.z;
        ic[9] = Iw[8] + mass * (cc - c.z * c.z);
    }
    tsync();
    if (warp == serial_warp) {
        for (int d = max_depth - 1; d >= 0; d--) {  // parents gather the composite inertia of their children
            DWBC_UNROLL(2)
            for (int k = 0; k < 2; k++) {
                if (cd_[k] != d || cnch_[k] == 0) continue;
                double *bb = rb + cb_[k] * RB_STRIDE;
                if (cnch_[k] <= 4) {
                    DWBC_UNROLL(4)
                    for (int j = 0; j < 4; j++) {
                        if (cch_[k][j] < 0) continue;
                        const double *cb = rb + cch_[k][j] * RB_STRIDE;
                        DWBC_UNROLL(10)
                        for (int i = 0; i < 10; i++) bb[RB_IC + i] += cb[RB_IC + i];
                    }
                } else {  // more than four children: read the list from the model
                    const int b = cb_[k];
                    for (int ci = md->child_start[b]; ci < md->child_start[b + 1]; ci++) {
                        const double *cb = rb + md->child_list[ci] * RB_STRIDE;
                        DWBC_UNROLL(10)
                        for (int i = 0; i < 10; i++) bb[RB_IC + i] += cb[RB_IC + i];
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---- contact-consistent dynamics for a compile-time contact dof C (6 or 12): everything from Jc to W^+, tau_grav, NwJw ----
// On return: P (R1) holds A^{-1}N_c with W^+ in its [6:,6:] block, T (R3) = J P, and the small items of R4 are filled.
// Returns false (uniformly) if a factorisation failed.
template <class S, int C>
__device__ __forceinline__ bool lean_contact_task_T(const DevModel *__restrict__ md, const BatchCfg &cfg, const Work &w, long long s,
                                                    double *sm, const int *cix, long long *tm) {
    using L = LeanSmem<S>;
    constexpr int N = S::N, M = S::M, LDN = S::LDN, LDM = S::LDM, LDT = S::LDT, LDC = S::LDC, TS = S::TS, R = C - 6;
    const int tid = t_id(), nt = TEAM, warp = lean::warp_id(), lane = lane_id();
    double *P = sm + L::R1, *R2 = sm + L::R2, *rb = sm + L::R3, *T = sm + L::R3;
    double *Jbar = R2 + L::oJbar, *Jc = R2 + L::oJc, *X = R2 + L::oX, *Lam = R2 + L::oLam, *Lw = R2 + L::oLw;
    double *V = sm + L::V, *Aug = sm + L::Aug, *Jb6 = sm + L::Jb6, *NwJw = sm + L::NwJw, *Rc = sm + L::Rc;
    double *rd = sm + L::rd, *Gs = sm + L::Gs, *u = sm + L::u, *PC = sm + L::PC, *tgv = sm + L::tg, *misc = sm + L::misc;
    int &ok = *(int *)(misc + 1);
    // Jc (C x N): stacked 6D point Jacobians of the active contacts; rotation of their links for the wrench-cone rows
    for (int k = 0; k < C / 6; k++) {
        const ContactCfg &c = cfg.contacts[cix[k]];
        point_jacobian_rb((const uint64_t *)(sm + L::anc), rb, N, c.link, v3(c.point[0], c.point[1], c.point[2]), Jc, LDC, 6 * k, 0, 6, tid, nt);
        if (tid < 9) Rc[9 * k + tid] = rb[c.link * RB_STRIDE + RB_R + tid];
    }
    tsync();
    LEAN_STAMP(tm, 4);
    // X = A^{-1} Jc^T  (N x C)
    lean::mma_gemm<N, C, N, LDN, LDC, false, true, (C + 7) / 8>(P, Jc, [&](int i, int j, double v) { X[i + j * LDN] = v; });
    tsync();
    LEAN_STAMP(tm, 5);
    // Jc X (C x C, SPD) -> Lambda_c
    lean::mma_gemm<C, C, N, LDC, LDN, false, false, 1>(Jc, X, [&](int i, int j, double v) { Lam[i + j * LDC] = v; });
    // While one warp factors the first panel of Jc X, another prepares the null-space basis of W (below): T = B_1^{-T} [B_2^T ...],
    // B_i = Jc[6i:6i+6, 0:6], by Gauss-Jordan on [B_1^T | B_2^T ...].  It only needs Jc and its own scratch.
    constexpr int lda = 13;
    auto null_basis_prep = [&](int busy_warp) {
        if constexpr (R > 0) {
            if (warp == (busy_warp + 1) % lean::NW) {
                for (int e = lane; e < 6 * (6 + R); e += 32) {
                    const int col = e / 6, row = e - col * 6;
                    Aug[row + col * lda] = Jc[col + row * LDC];
                }
                __syncwarp();
                const double cond1 = wgauss_jordan(6, R, Aug, lda);
                if (!(cond1 > 1e-14) && lane == 0) ok = 0;
            }
        }
    };
    lean::spd_inverse<C, LDC, LDC>(Lam, Lw, rd, misc, 1, null_basis_prep);
    if (!(misc[0] > 0.0)) return false;
    LEAN_STAMP(tm, 6);
    // Jbar^T = Lambda_c X^T  (C x N)
    lean::mma_gemm<C, N, C, LDC, LDN, false, true, 3>(Lam, X, [&](int i, int j, double v) { Jbar[i + j * LDC] = v; });
    tsync();
    // P = A^{-1} N_c = A^{-1} - X Jbar^T  (symmetric rank-C update)
    lean::mma_gemm_sym<N, C, LDN, LDC, false, false>(X, Jbar, [&](int i, int j, double v) {
        P[i + j * LDN] -= v;
        if (i != j) P[j + i * LDN] -= v;
    });
    // Jbar^T[:, 6:] and P_C = Jbar^T G are what the QPs need later
    for (int e = tid; e < C * M; e += nt) {
        const int j = e / C, i = e - j * C;
        Jb6[i + j * LDC] = Jbar[i + (6 + j) * LDC];
    }
    lean::gemv<C, N, LDC, false, 4>(Jbar, Gs, [&](int i, double v) { PC[i] = v; });
    LEAN_STAMP(tm, 7);
    if constexpr (R > 0) {
        // ---- basis V (M x R) of null(W): {S Jc^T f : (Jc^T f)[0:6] = 0}, from the T prepared during the inverse of Jc X above ----
        for (int e = tid; e < M * R; e += nt) {  // V_raw(a, b) = Jc[6+b][6+a] - sum_e Jc[e][6+a] T(e, b)
            const int b = e / M, a = e - b * M;
            double v = Jc[(6 + b) + (6 + a) * LDC];
            DWBC_UNROLL(6)
            for (int k = 0; k < 6; k++) v = fma(-Jc[k + (6 + a) * LDC], Aug[k + (6 + b) * lda], v);
            V[a + b * LDM] = v;
        }
    }
    tsync();
    // Two independent strands from here to the next barrier.  Warp 3 alone (warp-level synchronisation only) turns V_raw into the
    // scaled basis V, G = V^T V, G^{-1} and V G^{-2}; the other three warps form u and the task Jacobians of all levels.  The strand of
    // warp 3 is a chain of seven short dependent steps that used to keep the whole team at seven barriers.
    const bool vwarp = R > 0 && warp == 3;
    const int jtid = tid, jnt = R > 0 ? nt - 32 : nt;  // thread index / count of the Jacobian strand
    if constexpr (R > 0) {
        if (vwarp) {
        // V does not have to be orthonormal: for ANY basis V of null(W), with G = V^T V,
        //     (W + V V^T)^{-1} = W^+ + V G^{-2} V^T          (W symmetric, range(W) orthogonal to span V)
        // and NwJw = V (Jbar^T[0:r, 6:] V)^{-1} does not depend on the basis.  Only the columns are normalised (conditioning of
        // W + V V^T); G^{-1} is a 6 x 6 SPD inverse on one warp.  VG = V G^{-2} is kept for the correction after the big inverse.
        // (The orthonormal V2 of the reference is a public field of DWBC_OUT_ALL only.)
        double *Gm = Aug, *G2 = Aug + 6 * lda;  // two 6 x 6 blocks inside the 13 x 13 scratch
        if (lane < R) {
            double d = 0.0;
            for (int a = 0; a < M; a++) d = fma(V[a + lane * LDM], V[a + lane * LDM], d);
            Gm[12 * lda + lane] = rsqrt(d);  // column scales (row 12 of the scratch is free)
        }
        __syncwarp();
        for (int e = lane; e < M * R; e += 32) V[(e % M) + (e / M) * LDM] *= Gm[12 * lda + e / M];
        __syncwarp();
        for (int e = lane; e < R * R; e += 32) {
            const int j = e / R, i = e - j * R;
            double d = 0.0;
            for (int a = 0; a < M; a++) d = fma(V[a + i * LDM], V[a + j * LDM], d);
            Gm[i + j * lda] = d;
        }
        __syncwarp();
        {
            const double cg = lean::warp_spd_inverse_small<R, lda>(Gm, G2);
            if (!(cg > 1e-13) && lane == 0) ok = 0;
        }
        __syncwarp();
        for (int e = lane; e < R * R; e += 32) {  // G2 = G^{-1} G^{-1}
            const int j = e / R, i = e - j * R;
            double d = 0.0;
            DWBC_UNROLL(R)
            for (int k = 0; k < R; k++) d = fma(Gm[i + k * lda], Gm[k + j * lda], d);
            G2[i + j * lda] = d;
        }
        __syncwarp();
        double *VG = sm + L::NwJw;  // V G^{-2} (M x R); NwJw itself is formed last, in the same place
        for (int e = lane; e < M * R; e += 32) {
            const int j = e / M, i = e - j * M;
            double d = 0.0;
            DWBC_UNROLL(R)
            for (int k = 0; k < R; k++) d = fma(V[i + k * LDM], G2[k + j * lda], d);
            VG[i + j * LDM] = d;
        }
        __syncwarp();
        }
    }
    LEAN_STAMP(tm, 8);
    if (!vwarp) {
    // u = (A^{-1}N_c)[6:, :] G
    lean::gemv<M, N, LDN, false, 2>(P + 6, Gs, [&](int i, double v) { u[i] = v; }, jtid, jnt);
    // ---- J_task of all levels (stacked, R2: Jc / X / Lambda_c are dead) ----
    if constexpr (TS > 0) {
        double *Js = R2;
        using LV = typename S::Lv;
        const double *A06 = sm + L::A06, *co = sm + L::com;
        lean::static_for<0, LV::NL>([&](auto l_) {
            constexpr int l = decltype(l_)::value, t = LV::dof(l), r0 = LV::r0(l);
            const TaskCfg &tc = cfg.tasks[l];
            if (tc.custom) {
                const double *src = w.jcustom[l] + (w.jcustom_per_state[l] ? s * (long long)t * N : 0);
                for (int e = jtid; e < t * N; e += jnt) {
                    const int j = e / t, i = e - j * t;
                    Js[r0 + i + j * LDT] = src[e];
                }
            } else {
                int row = r0;
                for (int k = 0; k < tc.link_count; k++) {
                    const TaskLinkCfg &tl = cfg.tlinks[tc.link_start + k];
                    const int first = (tl.mode == DWBC_TASK_LINK_ROTATION || tl.mode == DWBC_TASK_LINK_ROTATION_CUSTOM_FRAME) ? 3 : 0;
                    if (tl.link == S::NB) {
                        // COM pseudo link: jac_ = SI_body^{-1} CMM_ (reference src/dwbc.cpp:331-356), see build_task_jacobian
                        const Vec3 c = ld3(co + 3);
                        const double Mt = co[15];
                        const double *I = co + 6;
                        const double c00 = I[4] * I[8] - I[5] * I[7], c01 = I[5] * I[6] - I[3] * I[8], c02 = I[3] * I[7] - I[4] * I[6];
                        const double det = I[0] * c00 + I[1] * c01 + I[2] * c02, rdet = 1.0 / det;
                        const double c11 = I[0] * I[8] - I[2] * I[6], c12 = I[1] * I[6] - I[0] * I[7], c22 = I[0] * I[4] - I[1] * I[3];
                        const double inv[9] = {c00 * rdet, c01 * rdet, c02 * rdet, c01 * rdet, c11 * rdet, c12 * rdet, c02 * rdet, c12 * rdet, c22 * rdet};
                        for (int j = jtid; j < N; j += jnt) {
                            const Vec3 al = v3(A06[0 + 6 * j], A06[1 + 6 * j], A06[2 + 6 * j]), aa = v3(A06[3 + 6 * j], A06[4 + 6 * j], A06[5 + 6 * j]);
                            const Vec3 lin = (1.0 / Mt) * al;
                            const Vec3 h = cross(al, c) + mv(rb + RB_R, aa);
                            const Vec3 ang = mv(inv, h);
                            const double v[6] = {lin.x, lin.y, lin.z, ang.x, ang.y, ang.z};
                            for (int r = 0; r < tl.dof; r++) Js[row + r + j * LDT] = v[first + r];
                        }
                    } else {
                        Vec3 pl = v3(0, 0, 0);
                        if (tl.mode == DWBC_TASK_LINK_6D_COM_FRAME || tl.mode == DWBC_TASK_LINK_POSITION_COM_FRAME)
                            pl = ld3(md->com[tl.link]);
                        else if (tl.mode == DWBC_TASK_LINK_6D_CUSTOM_FRAME || tl.mode == DWBC_TASK_LINK_POSITION_CUSTOM_FRAME)
                            pl = v3(tl.point[0], tl.point[1], tl.point[2]);
                        point_jacobian_rb((const uint64_t *)(sm + L::anc), rb, N, tl.link, pl, Js, LDT, row, first, tl.dof, jtid, jnt);
                    }
                    row += tl.dof;
                }
            }
        });
    }
    }
    // ---- T = J P (R3: the sweep records are dead after this) ----
    if constexpr (TS > 0) {
        double *Js = R2;
        tsync();
        LEAN_STAMP(tm, 9);
        lean::mma_gemm<TS, N, N, LDT, LDN, false, false, (N + 7) / 8>(Js, P, [&](int i, int j, double v) { T[i + j * LDT] = v; });
    }
    tsync();
    LEAN_STAMP(tm, 10);
    // ---- W^+ = (W + V V^T)^{-1} - V V^T in place in P[6:, 6:] ----
    double *Wb = P + 6 + 6 * LDN;
    if constexpr (R > 0) {
        lean::mma_gemm_sym<M, R, LDM, LDM, false, true>(V, V, [&](int i, int j, double v) {
            Wb[i + j * LDN] += v;
            if (i != j) Wb[j + i * LDN] += v;
        });
    }
    LEAN_STAMP(tm, 11);
    // While one warp factors the first panel, another prepares NwJw = V Mx^{-1}: Mx = Jbar^T[0:R, 6:] V (R x R) with the identity beside it,
    // inverted by Gauss-Jordan in the 13 x 13 scratch (free again: G and G^{-2} went into V G^{-2}).  Independent of the inverse.
    auto mx_prep = [&](int busy_warp) {
        if constexpr (R > 0) {
            if (warp == (busy_warp + 1) % lean::NW) {
                for (int e = lane; e < R * R; e += 32) {
                    const int j = e / R, i = e - j * R;
                    double d = 0.0;
                    for (int a = 0; a < M; a++) d = fma(Jb6[i + a * LDC], V[a + j * LDM], d);
                    Aug[i + j * lda] = d;
                    Aug[i + (R + j) * lda] = (i == j) ? 1.0 : 0.0;
                }
                __syncwarp();
                const double cond2 = wgauss_jordan(R, R, Aug, lda);
                if (!(cond2 > 1e-14) && lane == 0) ok = 0;
            }
        }
    };
    lean::spd_inverse<M, LDN, LDM>(Wb, R2, rd, misc, 2, mx_prep);
    if (!(misc[0] > 0.0)) return false;
    LEAN_STAMP(tm, 12);
    if constexpr (R > 0) {
        lean::mma_gemm_sym<M, R, LDM, LDM, false, true>(sm + L::NwJw, V, [&](int i, int j, double v) {  // W^+ = (W + V V^T)^{-1} - (V G^{-2}) V^T
            Wb[i + j * LDN] -= v;
            if (i != j) Wb[j + i * LDN] -= v;
        });
        tsync();
        lean::gemv<M, M, LDN, false, 2>(Wb, u, [&](int i, double v) { tgv[i] = v; });  // tau_grav = W^+ u
        for (int e = tid; e < M * R; e += nt) {
            const int j = e / M, i = e - j * M;
            double d = 0.0;
            DWBC_UNROLL(R)
            for (int k = 0; k < R; k++) d = fma(V[i + k * LDM], Aug[k + (R + j) * lda], d);
            NwJw[i + j * LDM] = d;
        }
    } else {
        tsync();
        lean::gemv<M, M, LDN, false, 2>(Wb, u, [&](int i, double v) { tgv[i] = v; });
    }
    tsync();
    LEAN_STAMP(tm, 13);
    return true;
}

// ---- task levels: Y = W^+ Tm^T, M_l = Tm_l Y_l, JL_l = Y_l M_l^{-1}, null-space recursion -> Ntorque (stacked, M x TS, in R2) ----
template <class LV>
__host__ __device__ constexpr int lean_count_big() {
    int c = 0;
    for (int l = 0; l < LV::NL; l++) c += LV::dof(l) > 8 ? 1 : 0;
    return c;
}
template <class LV>
__host__ __device__ constexpr int lean_small_index(int l) {  // number of small levels before level l
    int c = 0;
    for (int k = 0; k < l; k++) c += LV::dof(k) <= 8 ? 1 : 0;
    return c;
}
template <class LV>
__host__ __device__ constexpr int lean_first_big() {
    for (int l = 0; l < LV::NL; l++)
        if (LV::dof(l) > 8) return l;
    return -1;
}
template <class S>
__device__ __forceinline__ bool lean_task_levels(double *sm, long long *tm) {
    using L = LeanSmem<S>;
    using LV = typename S::Lv;
    constexpr int M = S::M, LDN = S::LDN, LDM = S::LDM, LDT = S::LDT, TS = S::TS;
    const int tid = t_id(), warp = lean::warp_id(), lane = lane_id();
    double *P = sm + L::R1, *Y = sm + L::R2, *T = sm + L::R3, *misc = sm + L::misc, *rd = sm + L::rd;
    double *Wb = P + 6 + 6 * LDN, *Tm = T + 6 * LDT;
    double *cnd = misc + 2;  // per level: relative pivot of M_l
    // Y = W^+ Tm^T   (M x TS)
    lean::mma_gemm<M, TS, M, LDN, LDT, false, true, 2>(Wb, Tm, [&](int i, int j, double v) { Y[i + j * LDM] = v; });
    tsync();
    LEAN_STAMP(tm, 14);
    // M_l = Tm_l Y_l (t x t, SPD) into the task blocks of R1 (W^+ is dead: tau_grav and Y are formed)
    lean::static_for<0, LV::NL>([&](auto l_) {
        constexpr int l = decltype(l_)::value, t = LV::dof(l), r0 = LV::r0(l), ldb = L::ldb(l);
        double *Ml = P + L::mo(l);
        lean::mma_gemm<t, t, M, LDT, LDM, false, false, 1>(Tm + r0, Y + r0 * LDM, [&](int i, int j, double v) { Ml[i + j * ldb] = v; }, l);
    });
    tsync();
    LEAN_STAMP(tm, 15);
    // inverses: levels with t <= 8 by one warp each; a larger one by the team, and while its first Cholesky panel occupies one warp the
    // small levels are inverted by the other warps (they used to run before it, with the team waiting at a barrier)
    constexpr int NBIG = lean_count_big<LV>(), LBIG = lean_first_big<LV>();
    auto small_levels = [&](int busy_warp) {
        lean::static_for<0, LV::NL>([&](auto l_) {
            constexpr int l = decltype(l_)::value, t = LV::dof(l), ldb = L::ldb(l);
            if constexpr (t <= 8) {
                constexpr int idx = lean_small_index<LV>(l);  // k-th small level -> k-th warp that is not the panel warp
                const int wl = (idx + ((busy_warp >= 0 && idx >= busy_warp) ? 1 : 0)) % lean::NW;
                if (warp == wl) {
                    const double c = lean::warp_spd_inverse_small<t, ldb>(P + L::mo(l), P + L::mo(l) + ldb * t);
                    if (lane == 0) cnd[l] = c;
                }
            }
        });
    };
    if constexpr (NBIG == 0) small_levels(-1);
    lean::static_for<0, LV::NL>([&](auto l_) {
        constexpr int l = decltype(l_)::value, t = LV::dof(l), ldb = L::ldb(l);
        if constexpr (t > 8) {
            double *Ml = P + L::mo(l);
            double dmax = 0.0;
            for (int i = 0; i < t; i++) dmax = fmax(dmax, Ml[i + i * ldb]);
            if constexpr (l == LBIG) lean::spd_inverse<t, ldb, ldb>(Ml, Ml + ldb * t, rd, misc, l, small_levels);
            else lean::spd_inverse<t, ldb, ldb>(Ml, Ml + ldb * t, rd, misc, l);
            if (tid == 0) cnd[l] = dmax > 0.0 ? misc[0] / dmax : 0.0;
        }
    });
    tsync();
    LEAN_STAMP(tm, 16);
    bool okl = true;
    for (int l = 0; l < LV::NL; l++) okl = okl && (cnd[l] > 1e-14);
    // JL_l = Y_l M_l^{-1} in place (each work item owns one 8-row band of Y_l)
    lean::static_for<0, LV::NL>([&](auto l_) {
        constexpr int l = decltype(l_)::value, t = LV::dof(l), r0 = LV::r0(l), ldb = L::ldb(l);
        double *Yl = Y + r0 * LDM;
        lean::mma_gemm<M, t, t, LDM, ldb, false, false, (t + 7) / 8>(Yl, P + L::mo(l), [&](int i, int j, double v) { Yl[i + j * LDM] = v; }, l);
    });
    tsync();
    LEAN_STAMP(tm, 17);
    // Ntorque_l = JL_l - Nta[:, 0:r_l] (Tm[0:r_l, :] JL_l), in place, levels in order
    lean::static_for<1, LV::NL>([&](auto l_) {
        constexpr int l = decltype(l_)::value, t = LV::dof(l), rl = LV::r0(l), ldc = ld4(rl);
        double *Cs = P + L::Cs, *Yl = Y + rl * LDM;
        lean::mma_gemm<rl, t, M, LDT, LDM, false, false, 1>(Tm, Yl, [&](int i, int j, double v) { Cs[i + j * ldc] = v; });
        tsync();
        lean::mma_gemm<M, t, rl, LDM, ldc, false, false, (t + 7) / 8>(Y, Cs, [&](int i, int j, double v) { Yl[i + j * LDM] -= v; });
        tsync();
    });
    LEAN_STAMP(tm, 18);
    return okl;
}

// The x = 0 test of one QP (reference quick exit, same margins as hqp_kernel's quick_feasible): torque rows tlim -/+ base > 0 and
// contact rows -(A_const A_rot (P_C - Jbar^T[:,6:] base)) > 0, strictly.  Uniform result.  Synchronises.
template <class S>
__device__ __noinline__ bool lean_quick_feasible(const BatchCfg *__restrict__ cfg, double *sm, int c, const int *cix, const double *base) {
    using L = LeanSmem<S>;
    constexpr int M = S::M, LDC = S::LDC;
    const int tid = t_id(), nt = TEAM;
    double *Jb6 = sm + L::Jb6, *PC = sm + L::PC, *lf = sm + L::lf, *Rc = sm + L::Rc, *misc = sm + L::misc;
    int &flag = *(int *)(misc + 7);
    if (tid == 0) flag = 1;
    // the limit of this thread's joint and the parameters of this thread's contact are fetched before the product: their global-memory
    // latency then hides behind it instead of following the barrier
    static_assert(M <= TEAM, "one torque row pair per thread");
    const bool has_tl = cfg->tlim_set && tid < M;
    const double tl = has_tl ? cfg->tlim[tid] : 0.0;
    ContactCfg cc{};
    if (tid < c / 6) cc = cfg->contacts[cix[tid]];
    if (c == 12) lean::gemv<12, M, LDC, false, 4>(Jb6, base, [&](int i, double v) { lf[i] = PC[i] - v; });
    else lean::gemv<6, M, LDC, false, 4>(Jb6, base, [&](int i, double v) { lf[i] = PC[i] - v; });  // (one out-of-line copy serves both classes)
    tsync();
    bool okq = true;
    if (has_tl) {
        const double lo = tl - base[tid], hi = tl + base[tid];
        okq = okq && lo > 1e-8 * (1.0 + fabs(lo)) && hi > 1e-8 * (1.0 + fabs(hi));
    }
    if (tid < c / 6) {
        double o[10];
        contact_rows(cc, Rc + 9 * tid, ld3(lf + 6 * tid), ld3(lf + 6 * tid + 3), o);
        for (int a = 0; a < 10; a++) okq = okq && (-o[a] > 1e-8 * (1.0 + fabs(o[a])));
    }
    if (!okq) flag = 0;
    tsync();
    const bool res = flag != 0;
    tsync();
    return res;
}

template <class S, int CSEL>
__global__ void __launch_bounds__(TEAM * LEAN_GANG, 1) lean_cycle_kernel(const DevModel *__restrict__ md, const BatchCfg *__restrict__ cfg, const BatchCfg cv, Work w, LeanArgs a) {
    static_assert(CSEL == 12 || CSEL == 6, "contact dof of the class");
    static_assert(lean::NW == 4, "the warp roles of the strands and of the side work assume a team of four warps");
    // LEAN_GANG teams per CTA, one state each, on private slices of the shared-memory arena.  tsync() is the CTA-wide barrier, so
    // the teams of an SM walk through the phases of the cycle together and fetch each stretch of this (large, mostly straight-line)
    // kernel once per LEAN_GANG states: the instruction cache holds 32 KB, and independent one-state CTAs drifting apart had it
    // thrashing (profiles/r02_*: icc hit rate 66 %).  A team whose slot lies beyond the list leaves at once; exited warps are not
    // waited for by the barrier.
    // cv: the configuration BY VALUE (kernel-parameter constant bank): the contact / task-link descriptions steer the Jacobian phases
    // with uniform reads, and from global memory (cfg) every one of them was a dependent L2 latency.  cfg stays for the per-thread reads.
    const int slot = (int)blockIdx.x * LEAN_GANG + (int)(threadIdx.x / TEAM);
    if (slot >= *a.pcount) return;
    using L = LeanSmem<S>;
    using LV = typename S::Lv;
    constexpr int N = S::N, M = S::M, NB = S::NB, LDN = S::LDN, LDM = S::LDM, TS = S::TS;
    DWBC_DYN_SMEM(smem);
    double *sm = (double *)smem + (threadIdx.x / TEAM) * L::total;
    const int tid = t_id(), nt = TEAM, warp = lean::warp_id(), lane = lane_id();
    const long long s = a.plist[slot];
    double *P = sm + L::R1, *R2 = sm + L::R2, *rb = sm + L::R3;
    double *Gs = sm + L::Gs, *fs = sm + L::fs, *tgv = sm + L::tg, *tt = sm + L::tt, *base = sm + L::base, *com = sm + L::com, *misc = sm + L::misc;
    double *A06 = sm + L::A06, *rd = sm + L::rd;
    int &ok = *(int *)(misc + 1);
    int *cix = (int *)(misc + 8);  // registered index of the active contacts (2 ints)
    int &cdof = *(int *)(misc + 6);
    double *tau_grav = w.tau_grav + s * M, *tau_task = w.tau_task + s * M, *tau_contact = w.tau_contact + s * M;
    // ---- a1: sweeps by one warp (rotating with the block index: warp k of every CTA shares an SM sub-partition) ----
    long long *tm = a.timing ? a.timing + s * 32 : nullptr;
    LEAN_STAMP(tm, 0);
    const int serial_warp = slot % lean::NW;
    uint64_t *anc_s = (uint64_t *)(sm + L::anc);
    rbd_sweeps_lean(md, a.q + s * (N + 1), rb, serial_warp, anc_s);
    if (warp == serial_warp) {
        if (lane == 0) {
            // COM quantities of the COM pseudo link (reference src/dwbc.cpp:320-346)
            const double Mtot = rb[RB_IC];
            const Vec3 cfp = (1.0 / Mtot) * ld3(rb + RB_IC + 1);
            st3(com + 3, cfp);
            const double *io = rb + RB_IC + 4;
            const double cc = dot(cfp, cfp);
            com[6] = io[0] - Mtot * (cc - cfp.x * cfp.x);
            com[7] = com[9] = io[1] + Mtot * cfp.x * cfp.y;
            com[8] = com[12] = io[2] + Mtot * cfp.x * cfp.z;
            com[10] = io[3] - Mtot * (cc - cfp.y * cfp.y);
            com[11] = com[13] = io[4] + Mtot * cfp.y * cfp.z;
            com[14] = io[5] - Mtot * (cc - cfp.z * cfp.z);
            com[15] = Mtot;
            // active contacts of this state
            int k = 0, over = 0;
            for (int i = 0; i < cv.ncontact; i++) {
                if (!a.flags[s * cv.ncontact + i]) continue;
                if (k >= MAX_ACTIVE_CONTACTS) {
                    over = 1;
                    break;
                }
                cix[k++] = i;
            }
            cdof = over ? -1 : 6 * k;
            ok = 1;
        }
    }
    for (int i = tid; i < TS; i += nt) {
        int l = 0, r = i;
        lean::static_for<0, LV::NL>([&](auto l_) {
            constexpr int ll = decltype(l_)::value;
            if (i >= LV::r0(ll)) l = ll, r = i - LV::r0(ll);
        });
        fs[i] = w.fstar[l][s * LV::dof(l) + r];
    }
    tsync();
    LEAN_STAMP(tm, 1);
    // motion column s_j = [w; v] and force column F_j = Ic_body(j) s_j of every velocity coordinate (R2: 12 x N)
    double *Sc = R2, *Fc = R2 + 6 * N;
    for (int j = tid; j < N; j += nt) {
        Vec3 wj, vj, nj, fj;
        motion_col(rb, j, wj, vj);
        const int bj = j < 6 ? 0 : j - 5;
        inertia_apply(rb + bj * RB_STRIDE + RB_IC, wj, vj, nj, fj);
        st3(Sc + 6 * j, wj);
        st3(Sc + 6 * j + 3, vj);
        st3(Fc + 6 * j, nj);
        st3(Fc + 6 * j + 3, fj);
        Gs[j] = 9.81 * fj.z;  // G_ = 9.81 A(2, :)   (reference src/dwbc.cpp:358)
    }
    tsync();
    // mass matrix, lower triangle mirrored: A(i,j) = s_j . F_i when body(j) is body(i) or one of its ancestors (j <= i)
    {
        constexpr int wf = N + 1, slots = ((N + 1) / 2) * wf;
        for (int e = tid; e < slots; e += nt) {
            const int fa = e / wf, fb = e - fa * wf;
            int i, j;
            if (fb < N - fa) {
                j = fa;
                i = fa + fb;
            } else {
                j = N - 1 - fa;
                i = j + (fb - (N - fa));
                if (j == fa) continue;
            }
            const int bi = i < 6 ? 0 : i - 5, bj = j < 6 ? 0 : j - 5;
            double val = 0.0;
            if (bj == 0 || bj == bi || ((anc_s[bi] >> bj) & 1ull))
                val = dot(ld3(Sc + 6 * j), ld3(Fc + 6 * i)) + dot(ld3(Sc + 6 * j + 3), ld3(Fc + 6 * i + 3));
            P[i + j * LDN] = val;
            P[j + i * LDN] = val;
            if (j < 6) A06[j + 6 * i] = val;
            if (i < 6) A06[i + 6 * j] = val;
        }
    }
    const int c = cdof;
    LEAN_STAMP(tm, 2);
    // ---- a1: A^{-1} ----
    lean::spd_inverse<N, LDN, LDN>(P, R2, rd, misc, slot % lean::NW);
    LEAN_STAMP(tm, 3);
    bool good = misc[0] > 0.0 && c == CSEL;
    // ---- a2-a5 ----
    if (good) {
        good = lean_contact_task_T<S, CSEL>(md, cv, w, s, sm, cix, tm);
        good = good && ok != 0;
    }
    if (!good) {  // uniform: no contact / too many contacts / a factorisation failed -> status 0, zero torques
        for (int i = tid; i < M; i += nt) tau_grav[i] = 0.0, tau_task[i] = 0.0, tau_contact[i] = 0.0;
        if (tid == 0) w.status[s] = 0, w.cdof[s] = c, w.qp_iters[s] = 0;
        return;
    }
    for (int i = tid; i < M; i += nt) tau_grav[i] = tgv[i], tau_contact[i] = 0.0;
    int st = 1;
    bool task_ok = true;
    if constexpr (TS > 0) task_ok = lean_task_levels<S>(sm, tm);
    const double *Nta = R2;  // Ntorque of all levels, M x TS (LDM)
    bool active = false;
    if constexpr (TS > 0) {
        if (!a.hqp) {
            // closed form (reference src/dwbc.cpp:856-871): torque_task = sum_l Ntorque_l f*_l; always returns 1, a singular task space is reported
            lean::gemv<M, TS, LDM, false, 2>(Nta, fs, [&](int i, double v) { tt[i] = v; });
            tsync();
            if (task_ok) st |= 2;
        } else if (task_ok) {
            for (int i = tid; i < M; i += nt) tt[i] = 0.0;
            tsync();
            lean::static_for<0, LV::NL>([&](auto l_) {
                constexpr int l = decltype(l_)::value, t = LV::dof(l), r0 = LV::r0(l);
                if (!active) {
                    // base = tau_prev + Ntorque_l f*_l with tau_prev = tau_grav + torque_task so far
                    lean::gemv<M, t, LDM, false, 2>(Nta + r0 * LDM, fs + r0, [&](int i, double v) { base[i] = tgv[i] + tt[i] + v; });
                    tsync();
                    if (lean_quick_feasible<S>(cfg, sm, c, cix, base)) {
                        for (int i = tid; i < M; i += nt) tt[i] = base[i] - tgv[i];
                        tsync();
                    } else {
                        active = true;
                    }
                }
            });
            if (!active) st |= 2;
        }
    }
    if (a.redist && !active) {
        if (c <= 6) {
            st |= 4;  // no internal-force space: torque_contact_ = 0, returns 1 (reference src/dwbc.cpp:1562-1566)
        } else if (a.hqp && TS > 0 && !task_ok) {
            active = true;  // the redistribution QP still runs after a failed task stage: leave it to hqp_kernel
        } else {
            // tau_in = torque_grav_ + torque_task_ + torque_contact_ (= 0 here)
            for (int i = tid; i < M; i += nt) base[i] = tgv[i] + ((TS > 0 && (!a.hqp || task_ok)) ? tt[i] : 0.0);
            tsync();
            if (lean_quick_feasible<S>(cfg, sm, c, cix, base)) st |= 4;
            else active = true;
        }
    }
    LEAN_STAMP(tm, 19);
    // torque_task_ so far: final for an inactive state; for an active one hqp_kernel recomputes it (hqp) or reads it (closed form + redistribution QP)
    for (int i = tid; i < M; i += nt) tau_task[i] = (TS > 0 && (!a.hqp || task_ok)) ? tt[i] : 0.0;
    if (TS > 0 && a.hqp) {  // f_star_qp_ / contact_qp_ / active sets of a state whose QPs all end at x = 0 (hqp_kernel rewrites them for the others)
        lean::static_for<0, LV::NL>([&](auto l_) {
            constexpr int l = decltype(l_)::value, t = LV::dof(l);
            for (int i = tid; i < t; i += nt) w.fstar_qp[l][s * t + i] = 0.0;
            for (int i = tid; i < RMAX; i += nt) w.contact_qp[l][s * RMAX + i] = 0.0;
            if (tid < ACTIVE_WORDS) w.active[l][s * ACTIVE_WORDS + tid] = 0u;
        });
    }
    if (a.redist) {
        if (tid < ACTIVE_WORDS) w.active[MAX_LEVELS][s * ACTIVE_WORDS + tid] = 0u;
        for (int i = tid; i < RMAX; i += nt) w.cf_redis[s * RMAX + i] = 0.0;
    }
    if (!active) {
        if (tid == 0) w.status[s] = st, w.cdof[s] = c, w.qp_iters[s] = 0;
        LEAN_STAMP(tm, 20);
        return;
    }
    // ---- QP-active state: hand the operands of the QPs to hqp_kernel (generic workspace layout) and enlist the state ----
    {
        const double *Jb6 = sm + L::Jb6, *PC = sm + L::PC, *NwJw = sm + L::NwJw, *Rc = sm + L::Rc;
        double *gJ = w.Jbar + s * (long long)CMAX * N, *gN = w.NwJw + s * (long long)M * RMAX, *gP = w.PC + s * CMAX;
        for (int e = tid; e < CMAX * M; e += nt) {
            const int j = e / CMAX, i = e - j * CMAX;
            gJ[i + (6 + j) * CMAX] = i < c ? Jb6[i + j * S::LDC] : 0.0;
        }
        for (int e = tid; e < M * RMAX; e += nt) {
            const int j = e / M, i = e - j * M;
            gN[e] = c > 6 ? NwJw[i + j * LDM] : 0.0;
        }
        if (tid < CMAX) gP[tid] = tid < c ? PC[tid] : 0.0;
        for (int k = 0; k < c / 6; k++) {
            const int link = cv.contacts[cix[k]].link;
            if (tid < 9) w.pose[(s * NB + link) * POSE_STRIDE + tid] = Rc[9 * k + tid];
        }
        if (tid < MAX_ACTIVE_CONTACTS) w.cidx[s * MAX_ACTIVE_CONTACTS + tid] = tid < c / 6 ? cix[tid] : -1;
        if constexpr (TS > 0) {
            lean::static_for<0, LV::NL>([&](auto l_) {
                constexpr int l = decltype(l_)::value, t = LV::dof(l), r0 = LV::r0(l);
                double *g = w.Ntq[l] + s * (long long)M * t;
                for (int e = tid; e < M * t; e += nt) {
                    const int j = e / M, i = e - j * M;
                    g[e] = Nta[i + (r0 + j) * LDM];
                }
            });
        }
        if (tid == 0) {
            w.status[s] = 1 | ((a.hqp && task_ok) ? 8 : 0);
            w.cdof[s] = c;
            const int slot = atomicAdd(a.count, 1);
            a.list[slot] = (int)s;
        }
    }
    LEAN_STAMP(tm, 20);
}

}  // namespace dwbc
