"""Algorithmic FP64 work per solve: the reference's dense, as-written operation counts.

This is the single formula list of SURVEY.md section 8(d) ("ALGORITHMIC FLOPs per solve") evaluated literally
(2*m*k*n per product, n^3/3 Cholesky, 2n^3 LU inverse / solve against identity, (4/3)n^3 QR), split by the CUDA
kernel that performs the corresponding reference stage.  bench.py divides these counts by measured kernel time to
get `roofline.achieved`; because the kernels restructure the arithmetic (rank-c update instead of the 39^3 product,
closed-form null space instead of a COD, ...) they EXECUTE fewer flops than this -- savings must show up as more
solves/s at the same algorithmic count, never as an inflated rate.

Reference anchors: UpdateKinematics src/dwbc.cpp:279-371; CalculateContactConstraint src/wbd.cpp:108-143;
CalculateGravityCompensation src/wbd.cpp:186-192; CalculateJKT / CalculateTaskNullSpace src/wbd.cpp:207-213,257-261;
CalcSingleTaskTorqueWithQP src/dwbc.cpp:941-1127; CalcContactRedistribute src/dwbc.cpp:1372-1568.
"""
from __future__ import annotations

N_SYS, M_DOF = 39, 33


def rbd_flops(n=N_SYS):
    # FK 11.6k + CRBA 20.6k + link loop & RNEA 13.2k + CMM/COM/G 6.0k  (TOCABI figures of SURVEY 8(d))
    return 11.6e3 + 20.6e3 + 13.2e3 + 6.0e3


def ainv_flops(n=N_SYS):
    return n ** 3 / 3.0 + 2.0 * n ** 3  # llt() + solve(I)


def contact_jacobian_flops(c):
    return 0.7e3 * (c / 6.0)


def contact_dynamics_flops(c, n=N_SYS, m=M_DOF):
    r = max(c - 6, 0)
    f = 2 * c * n * n            # Jc A^-1
    f += 2 * c * c * n           # Jc A^-1 Jc^T
    f += 2 * c ** 3              # inverse
    f += 2 * c * c * n           # Jbar^T
    f += 2 * c * n * n           # N_c
    f += 2 * n ** 3              # A^-1 N_c
    f += (4.0 / 3 + (2 if r > 0 else 0) + 2 + 1) * m ** 3  # COD pseudo-inverse of W (+ V2)
    f += 4 * r * r * m + 2 * r ** 3                          # NwJw
    f += 2 * n * n + 2 * m * n + 2 * m * m + 2 * c * n       # gravity compensation + P_C
    return f


def task_level_flops(t, level, last, n=N_SYS, m=M_DOF):
    f = 4 * t * n * n + 2 * t * t * n + 2 * t ** 3              # Lambda_t
    f += 2 * t * t * n + 4 * t * n * n                          # Q
    f += 2 * t * m * m + 2 * t * t * m + (13.0 / 3) * t ** 3    # (Q W^+ Q^T)^+
    f += 2 * m * m * t + 2 * m * t * t                          # J_kt
    if not last:
        f += 2 * m * t * t + 4 * m * m * t + (2 * m ** 3 if level > 0 else 0)  # Null_i
    return f


def qp_build_flops(t, level, c, n=N_SYS, m=M_DOF):
    r = max(c - 6, 0)
    k = c // 6
    f = (2 * m * m * t if level > 0 else 0) + 2 * m * t * t
    f += 2 * 10 * k * (c * c + c * m) + 2 * 10 * k * m * (t + r)
    return f


def qp_solve_flops(rows, nv, iters):
    return iters * (2 * rows * nv + 6 * nv * nv) + nv ** 3


def torque_flops(t, level, c, m=M_DOF):
    r = max(c - 6, 0)
    return 2 * m * t + 2 * m * t * t + (2 * m * m if level > 0 else 0) + 2 * m * r


def per_kernel_flops(task_dofs, c=12, hqp=True, redist=True, tlim=True, qp_iters_total=None, n=N_SYS, m=M_DOF):
    """Algorithmic flops of ONE solve split by CUDA kernel.  qp_iters_total: measured mean active-set iterations per
    state summed over all QPs of the cycle (None: the 8-per-QP placeholder of SURVEY 8(d))."""
    r = max(c - 6, 0)
    rows = (2 * m if tlim else 0) + 10 * (c // 6)
    nl = len(task_dofs)
    nqp = (nl if hqp else 0) + (1 if (redist and r > 0) else 0)
    iters_each = 8.0 if qp_iters_total is None else (qp_iters_total / max(nqp, 1))
    out = {
        "rbd_kernel": rbd_flops(n),
        "ainv_kernel": ainv_flops(n),
        "contact_jacobian_kernel": contact_jacobian_flops(c),
        "contact_dynamics_kernel": contact_dynamics_flops(c, n, m),
        "task_space_kernel": 0.0,
        "hqp_kernel": 0.0,
    }
    for l, t in enumerate(task_dofs):
        out["task_space_kernel"] += task_level_flops(t, l, l == nl - 1, n, m)
        if hqp:
            out["hqp_kernel"] += qp_build_flops(t, l, c, n, m) + qp_solve_flops(rows, t + r, iters_each) + torque_flops(t, l, c, m)
        else:
            out["task_space_kernel"] += torque_flops(t, l, c, m)
    if redist and r > 0:
        out["hqp_kernel"] += 2 * 10 * (c // 6) * (c * c + c * m) + 2 * 10 * (c // 6) * m * r + qp_solve_flops(rows, r, iters_each) + 2 * m * r
    return out


def flops_per_solve(task_dofs, **kw):
    return sum(per_kernel_flops(task_dofs, **kw).values())


# ---- variants (BASELINE.json configs[3]); same as-written convention, coarser because the reference code is longer ----
def reduced_flops(task_dofs, c=12, hqp=True, redist=True, qp_iters_total=None, n=N_SYS, m=M_DOF):
    """Reduced-dimension pipeline (reference src/dwbc.cpp:2752-3446): full-model rbd + A^-1, the nc-chain CRBA / lumped body,
    J_R A^-1 J_R^T and its LLT inverse, then the contact / task / QP pipeline at n_R = 6 + c + 6, m_R = c + 6."""
    nr, mr, nc = 6 + c + 6, c + 6, m - c
    assemble = 8.0e3 + 2 * 6 * 6 * nc + 2 * nr * n * n + 2 * nr * nr * n + nr ** 3 / 3.0 + 2.0 * nr ** 3 + 2 * nr * n + 4 * nc * nc * 6
    pk = per_kernel_flops(task_dofs, c=c, hqp=hqp, redist=redist, tlim=True, qp_iters_total=qp_iters_total, n=nr, m=mr)
    red = assemble + pk["contact_dynamics_kernel"] + pk["task_space_kernel"] + pk["hqp_kernel"] + 2 * m * nr
    return {"rbd_kernel": rbd_flops(n), "ainv_kernel": ainv_flops(n), "contact_jacobian_kernel": contact_jacobian_flops(c), "reduced_kernels": red}


def lqp_level_flops(nv, k_prev, eq_rows, ineq_rows, carried_rows, iters=8.0):
    """One level of the null-space hierarchy (reference src/dwbc_hqp.cpp:23-58, 291-395): B Z, COD null space, Z update,
    H = (BZ)^T BZ, A Z for this level's and the carried inequality rows, then the QP in [u; v]."""
    k = max(k_prev - eq_rows, 0)
    f = 2 * eq_rows * nv * k_prev + (4.0 / 3) * k_prev ** 3 + 2 * nv * k_prev * k
    f += 2 * eq_rows * k_prev * k_prev + 2 * (ineq_rows + carried_rows) * nv * k_prev
    f += qp_solve_flops(ineq_rows + carried_rows, k_prev + ineq_rows, iters)
    return f, k


def lqp_flops(task_dofs, c=12, n=N_SYS, m=M_DOF, iters=8.0):
    """ConfigureLQP + CalcControlTorqueLQP (reference src/dwbc.cpp:4304-4454): y = [qdd; Fc], level 0 dynamics (6 eq, 2m torque
    rows carried), level 1 contact (c eq, 10 c/6 + 2m rows), then one level per task."""
    nv = n + c
    f = 2 * n * n  # y0 = -A^-1 B
    l0, k = lqp_level_flops(nv, nv, 6, 0, 2 * m, 0.0)
    f += l0
    carried = 2 * m
    l1, k = lqp_level_flops(nv, k, c, 10 * (c // 6) + 2 * m, carried, iters)
    f += l1
    carried += 10 * (c // 6) + 2 * m
    for t in task_dofs:
        lt, k = lqp_level_flops(nv, k, t, 0, carried, iters)
        f += lt
    return f + 2 * m * nv  # torque from y


# Compulsory HBM traffic of one solve: q, qdot, f*, flags in; three torque vectors + status out.
def io_bytes_per_solve(task_dofs, ncontact=2, n=N_SYS, m=M_DOF):
    return 8 * (n + 1) + 8 * n + 8 * sum(task_dofs) + ncontact, 3 * 8 * m + 4


if __name__ == "__main__":
    for name, dofs, kw in (("config0", (), dict(hqp=False)), ("config1", (3, 3), dict(hqp=False, redist=False, tlim=False)),
                           ("config2 DS", (3, 6, 21), {}), ("config2 SS", (3, 6, 21), dict(c=6))):
        pk = per_kernel_flops(dofs, **kw)
        print(name, {k: round(v / 1e3, 1) for k, v in pk.items()}, "total MFLOP", round(sum(pk.values()) / 1e6, 3))
