"""Unit tests of the compile-time-specialised team linear algebra (libdwbc_b200/csrc/lean_linalg.cuh) through the CUDA
emulation (test infrastructure, tests/emul): blocked Cholesky / triangular inverse / SYRK on the emulated DMMA, the
one-warp small inverse, the DMMA products and the matrix-vector product, against numpy."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "libdwbc_b200", "csrc")
EMUL = os.path.join(HERE, "emul")
OUT = os.path.join(EMUL, "build", "liblean_unit.so")
_dp = C.POINTER(C.c_double)


@pytest.fixture(scope="module")
def lib():
    srcs = [os.path.join(EMUL, "lean_unit.cpp"), os.path.join(EMUL, "cuda_emul.h"), os.path.join(CSRC, "lean_linalg.cuh"),
            os.path.join(CSRC, "team_linalg.cuh"), os.path.join(CSRC, "warp_linalg.cuh")]
    if not os.path.exists(OUT) or any(os.path.getmtime(s) > os.path.getmtime(OUT) for s in srcs):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++20", "-fPIC", "-shared", "-DDWBC_CPU_EMUL", "-I", EMUL, "-I", CSRC,
                               srcs[0], "-o", OUT, "-lpthread", "-Wno-unused-variable", "-Wno-unused-but-set-variable"])
    return C.CDLL(OUT)


def p(a):
    return a.ctypes.data_as(_dp)


def spd(n, rng, cond=1e4):
    q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    return (q * np.geomspace(1.0, cond, n)) @ q.T


@pytest.mark.parametrize("n", [39, 40, 33, 21, 12, 24, 18, 6, 9])
def test_spd_inverse(lib, n):
    rng = np.random.default_rng(n)
    A = spd(n, rng)
    A = 0.5 * (A + A.T)
    out = np.zeros((n, n))
    piv = np.zeros(1)
    assert lib.lean_spd_inverse(n, p(np.asfortranarray(A).T.copy()), p(out), p(piv)) == 0
    ref = np.linalg.inv(A)
    assert np.abs(out.T - ref).max() / np.abs(ref).max() < 1e-11
    assert np.array_equal(out, out.T)
    assert piv[0] > 0


@pytest.mark.parametrize("n", [1, 3, 6, 8])
def test_small_inverse(lib, n):
    rng = np.random.default_rng(100 + n)
    A = spd(n, rng, 1e3)
    A = 0.5 * (A + A.T)
    out = np.zeros((n, n))
    cond = np.zeros(1)
    assert lib.lean_small_inverse(n, p(A.T.copy()), p(out), p(cond)) == 0
    ref = np.linalg.inv(A)
    assert np.abs(out.T - ref).max() / np.abs(ref).max() < 1e-11
    assert cond[0] > 0


GEMMS = {0: (30, 39, 39), 1: (39, 12, 39), 2: (12, 39, 12), 3: (33, 30, 33), 4: (3, 6, 33), 5: (33, 21, 9), 6: (20, 27, 33)}


@pytest.mark.parametrize("which", sorted(GEMMS))
def test_gemm(lib, which):
    M, N, K = GEMMS[which]
    rng = np.random.default_rng(which)
    a, b = rng.standard_normal((M, K)), rng.standard_normal((K, N))
    c = np.zeros((N, M))
    assert lib.lean_gemm(which, p(a.T.copy()), p(b.T.copy()), p(c)) == 0
    assert np.abs(c.T - a @ b).max() < 1e-12


@pytest.mark.parametrize("which,shape", [(0, (33, 6)), (1, (39, 12)), (2, (21, 33))])
def test_gemm_sym(lib, which, shape):
    rng = np.random.default_rng(which)
    a = rng.standard_normal(shape)
    c = np.zeros((shape[0], shape[0]))
    assert lib.lean_gemm_sym(which, p(a.T.copy()), p(c)) == 0
    assert np.abs(c.T - a @ a.T).max() < 1e-12


@pytest.mark.parametrize("which,shape", [(0, (33, 39)), (1, (12, 39)), (2, (33, 21)), (3, (70, 7))])
def test_gemv(lib, which, shape):
    rng = np.random.default_rng(which)
    a, x = rng.standard_normal(shape), rng.standard_normal(shape[1])
    y = np.zeros(shape[0])
    assert lib.lean_gemv(which, p(a.T.copy()), p(x), p(y)) == 0
    assert np.abs(y - a @ x).max() < 1e-12


# ---- the run-time-dimension twins (lean::rt): one compiled copy serves every size ----
@pytest.mark.parametrize("n,ld,ldx", [(39, 44, 44), (40, 44, 44), (33, 44, 36), (21, 28, 28), (12, 12, 12), (6, 12, 12), (3, 4, 4), (9, 12, 12), (1, 4, 4)])
def test_rt_spd_inverse(lib, n, ld, ldx):
    rng = np.random.default_rng(200 + n)
    A = spd(n, rng)
    A = 0.5 * (A + A.T)
    out = np.zeros((n, n))
    piv = np.zeros(1)
    assert lib.lean_rt_spd_inverse(n, ld, ldx, p(A.T.copy()), p(out), p(piv)) == 0
    ref = np.linalg.inv(A)
    assert np.abs(out.T - ref).max() / np.abs(ref).max() < 1e-11
    assert np.array_equal(out, out.T) and piv[0] > 0


@pytest.mark.parametrize("n", [1, 3, 6, 8])
def test_rt_small_inverse(lib, n):
    rng = np.random.default_rng(300 + n)
    A = spd(n, rng, 1e3)
    A = 0.5 * (A + A.T)
    out = np.zeros((n, n))
    cond = np.zeros(1)
    assert lib.lean_rt_small_inverse(n, 12, p(A.T.copy()), p(out), p(cond)) == 0
    ref = np.linalg.inv(A)
    assert np.abs(out.T - ref).max() / np.abs(ref).max() < 1e-11 and cond[0] > 0


@pytest.mark.parametrize("M,N,K,ta,tb,nt,alpha,beta", [(30, 39, 39, 0, 0, 5, 1.0, 0), (39, 12, 39, 0, 1, 2, 1.0, 0), (12, 39, 12, 1, 0, 3, 1.0, 0),
                                                     (39, 39, 12, 0, 0, 5, -1.0, 1), (33, 30, 33, 1, 1, 2, 1.0, 0), (3, 6, 33, 0, 0, 1, 1.0, 0),
                                                     (33, 21, 9, 0, 0, 3, -1.0, 1), (33, 33, 6, 0, 1, 5, 1.0, 1), (21, 21, 33, 0, 0, 1, 2.0, 0)])
def test_rt_gemm(lib, M, N, K, ta, tb, nt, alpha, beta):
    lib.lean_rt_gemm.argtypes = [C.c_int] * 6 + [C.c_double, C.c_int, _dp, _dp, _dp]
    rng = np.random.default_rng(M * 100 + N)
    a, b, c0 = rng.standard_normal((M, K)), rng.standard_normal((K, N)), rng.standard_normal((M, N))
    c = c0.T.copy()
    assert lib.lean_rt_gemm(M, N, K, ta, tb, nt, alpha, beta, p(a.T.copy()), p(b.T.copy()), p(c)) == 0
    assert np.abs(c.T - (alpha * (a @ b) + (c0 if beta else 0))).max() < 1e-12


@pytest.mark.parametrize("M,K,ta,alpha,usez", [(33, 39, 0, 1.0, 0), (12, 39, 0, -1.0, 1), (33, 21, 1, 1.0, 1), (70, 7, 0, 2.0, 0), (6, 33, 0, -1.0, 1)])
def test_rt_gemv(lib, M, K, ta, alpha, usez):
    lib.lean_rt_gemv.argtypes = [C.c_int] * 3 + [C.c_double, C.c_int, _dp, _dp, _dp]
    rng = np.random.default_rng(M + K)
    a, x, z = rng.standard_normal((M, K)), rng.standard_normal(K), rng.standard_normal(M)
    y = z.copy()
    assert lib.lean_rt_gemv(M, K, ta, alpha, usez, p(a.T.copy()), p(x), p(y)) == 0
    assert np.abs(y - (alpha * (a @ x) + (z if usez else 0))).max() < 1e-12
