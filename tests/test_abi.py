"""The C-ABI shared library loads and exports every symbol include/dwbc_b200.h declares (no compute calls:
this runs in the GPU-less container).  Also checks the setup-time host logic that needs no device."""
import ctypes as C
import os
import re

import pytest

from tests.common import ROOT, URDF


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dwbc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dwbc_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import __graft_entry__ as g

    path = g.build_cuda()
    lib = C.CDLL(path)
    syms = header_symbols()
    assert len(syms) >= 28
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/dwbc_b200.h but not exported by {path}"


def test_python_prototypes_cover_header():
    from libdwbc_b200._lib import EXPORTS

    assert sorted(EXPORTS) == header_symbols()


def test_model_host_logic():
    """URDF reader + link lookup run on the host (reference src/dwbc.cpp:102-236, 393-404)."""
    from libdwbc_b200._lib import load_library

    L = load_library()
    m = L.dwbc_model_load_urdf(URDF.encode(), 1)
    assert m
    n, mm, nb, mass = C.c_int(), C.c_int(), C.c_int(), C.c_double()
    L.dwbc_model_info(m, C.byref(n), C.byref(mm), C.byref(nb), C.byref(mass))
    assert (n.value, mm.value, nb.value) == (39, 33, 34)
    assert abs(mass.value - 96.211282) < 1e-6
    assert L.dwbc_model_link_id(m, b"pelvis_link") == 0
    assert L.dwbc_model_link_id(m, b"UPPERBODY_LINK") == 15  # case-insensitive
    assert L.dwbc_model_link_id(m, b"COM") == 34
    assert L.dwbc_model_link_id(m, b"nope") == -1
    assert L.dwbc_model_link_name(m, 6).decode().lower() == "l_ankleroll_link"
    L.dwbc_model_destroy(m)
    assert not L.dwbc_model_load_urdf(b"/nonexistent.urdf", 1)
    assert L.dwbc_last_error()


def test_no_device_fails_loudly():
    """Without a GPU the product refuses to run (no CPU fallback)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from libdwbc_b200 import DwbcError, RobotDataBatch

    with pytest.raises(DwbcError):
        RobotDataBatch(URDF, 4, 0)
