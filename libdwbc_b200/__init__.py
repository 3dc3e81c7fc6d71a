"""libdwbc_b200 -- B200-native batched implementation of libdwbc's per-control-cycle whole-body-control path.

The product is the CUDA library `libdwbc_b200.so` (C ABI in include/dwbc_b200.h).  This package is the thin
Python host over it: `RobotDataBatch` mirrors the call sequence of the reference's `DWBC::RobotData`
(reference include/dwbc.h:59-430) for B states at once; `RlWBCBridgeBatch` mirrors the pybind bridge
(reference src/pybind/rl_wbc_bridge.cpp).  There is no CPU fallback: importing works anywhere, but every
compute call needs the CUDA library and a GPU and raises otherwise.
"""
from ._lib import DwbcError, library_path, load_library  # noqa: F401
from .batch import (  # noqa: F401
    CONTACT_6D,
    OUT_ALL,
    OUT_AUTO,
    OUT_TORQUES,
    SOLVE_GRAV,
    SOLVE_HQP,
    SOLVE_REDIST,
    SOLVE_REDIST_ANALYTIC,
    SOLVE_TASK,
    ST_CONTACT_OK,
    ST_QP_SLOW,
    ST_REDIST_OK,
    ST_TASK_OK,
    TASK_LINK_6D,
    TASK_LINK_6D_COM_FRAME,
    TASK_LINK_6D_CUSTOM_FRAME,
    TASK_LINK_POSITION,
    TASK_LINK_POSITION_COM_FRAME,
    TASK_LINK_POSITION_CUSTOM_FRAME,
    TASK_LINK_ROTATION,
    TASK_LINK_ROTATION_CUSTOM_FRAME,
    RobotDataBatch,
)
from .bridge import RlWBCBridge, RlWBCBridgeBatch  # noqa: F401,E402
from .multi import RobotDataMultiBatch  # noqa: F401,E402
