"""TEST INFRASTRUCTURE: runs libdwbc_b200.RobotDataBatch against tests/emul/build/libdwbc_emul.so, i.e. the
product's kernel sources compiled for the host (see cuda_emul.h).  "Device" buffers are numpy arrays.
Only tests may use this; the product's own backend is CUDA-only."""
import ctypes as C

import numpy as np

from libdwbc_b200._lib import declare
from tests.emul.build_emul import build

_lib = None


class EmulBackend:
    name = "emul"
    index = 0

    def lib(self):
        global _lib
        if _lib is None:
            _lib = declare(C.CDLL(build()))
        return _lib

    def empty(self, shape, dtype="f8"):
        return np.zeros(shape, dtype={"f8": np.float64, "f4": np.float32, "u1": np.uint8, "i4": np.int32, "u4": np.uint32}[dtype])

    def check(self, x, dtype="f8"):
        dt = {"f8": np.float64, "u1": np.uint8, "i4": np.int32}[dtype]
        assert isinstance(x, np.ndarray) and x.dtype == dt and x.flags["C_CONTIGUOUS"], (type(x), getattr(x, "dtype", None))
        return _NpView(x)

    def ptr(self, x):
        a = x.a if isinstance(x, _NpView) else x
        return C.c_void_p(a.ctypes.data)

    def stream(self):
        return None

    def to_numpy(self, x):
        return x.a if isinstance(x, _NpView) else x

    def from_numpy(self, a):
        return np.ascontiguousarray(a)


class _NpView:
    """numpy array with the few torch-like methods RobotDataBatch uses."""

    def __init__(self, a):
        self.a = a
        self.shape = a.shape
        self.ndim = a.ndim

    def t(self):
        return _NpView(self.a.T)

    def transpose(self, i, j):
        return _NpView(np.swapaxes(self.a, i, j))

    def contiguous(self):
        return _NpView(np.ascontiguousarray(self.a))
