"""Loader of the CUDA library.  Fails loudly: no fallback of any kind."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBNAME = "libdwbc_b200.so"


class DwbcError(RuntimeError):
    pass


def library_path() -> str:
    # DWBC_LIB: another build of the same library (kernel experiments of tools/, e.g. a different unrolling); default: the in-tree one
    return os.environ.get("DWBC_LIB") or os.path.join(_HERE, _LIBNAME)


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def declare(L):
    """ctypes prototypes of every symbol include/dwbc_b200.h declares."""
    vp = C.c_void_p
    L.dwbc_last_error.restype = C.c_char_p
    L.dwbc_launch_count.restype = C.c_longlong
    L.dwbc_model_load_urdf.restype = vp
    L.dwbc_model_load_urdf.argtypes = [C.c_char_p, C.c_int]
    L.dwbc_model_destroy.argtypes = [vp]
    L.dwbc_model_info.argtypes = [vp, _ip, _ip, _ip, _dp]
    L.dwbc_model_link_id.argtypes = [vp, C.c_char_p]
    L.dwbc_model_link_name.restype = C.c_char_p
    L.dwbc_model_link_name.argtypes = [vp, C.c_int]
    L.dwbc_batch_create.restype = vp
    L.dwbc_batch_create.argtypes = [vp, C.c_int, C.c_int]
    L.dwbc_batch_destroy.argtypes = [vp]
    L.dwbc_batch_size.argtypes = [vp]
    L.dwbc_batch_limits.argtypes = [vp, _ip, _ip, _ip, _ip]
    L.dwbc_batch_add_contact.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double]
    L.dwbc_batch_set_friction.argtypes = [vp, C.c_int, C.c_double, C.c_double]
    L.dwbc_batch_add_task.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp]
    L.dwbc_batch_add_custom_task.argtypes = [vp, C.c_int, C.c_int]
    L.dwbc_batch_clear_tasks.argtypes = [vp]
    L.dwbc_batch_set_torque_limit.argtypes = [vp, _dp]
    L.dwbc_batch_update_kinematics.argtypes = [vp, vp, vp, vp]
    L.dwbc_batch_set_contact.argtypes = [vp, vp, vp]
    L.dwbc_batch_calc_contact_constraint.argtypes = [vp, vp]
    L.dwbc_batch_set_task.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
    L.dwbc_batch_solve.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.dwbc_batch_get_matrix.argtypes = [vp, C.c_int, C.c_int, vp, _ip, _ip, vp]
    L.dwbc_batch_get_active_set.argtypes = [vp, C.c_int, vp, vp]
    L.dwbc_batch_get_contact_force.argtypes = [vp, vp, vp, vp]
    L.dwbc_batch_get_zmp.argtypes = [vp, vp, vp, vp]
    L.dwbc_batch_cycle_host.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.dwbc_batch_set_host_slices.argtypes = [vp, C.c_int]
    L.dwbc_probe_fp64.argtypes = [C.c_int, C.c_double, _dp]
    L.dwbc_batch_get_qp_iterations.argtypes = [vp, vp, vp]
    L.dwbc_batch_set_profiling.argtypes = [vp, C.c_int]
    L.dwbc_batch_get_kernel_times.argtypes = [vp, _dp, C.POINTER(C.c_longlong), C.c_int]
    L.dwbc_kernel_name.restype = C.c_char_p
    L.dwbc_kernel_name.argtypes = [C.c_int]
    L.dwbc_batch_reduced_dynamics.argtypes = [vp, vp, vp]
    L.dwbc_batch_reduced_calc_contact_constraint.argtypes = [vp, vp]
    L.dwbc_batch_reduced_solve.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    L.dwbc_batch_set_reduced_torque_limit.argtypes = [vp, _dp, C.c_int]
    L.dwbc_batch_reduced_dims.argtypes = [vp, _ip, _ip, _ip, _ip, _ip]
    L.dwbc_batch_get_reduced_matrix.argtypes = [vp, C.c_int, C.c_int, vp, _ip, _ip, vp]
    L.dwbc_batch_calc_control_torque_lqp.argtypes = [vp, vp, vp, vp, vp]
    L.dwbc_batch_calc_control_torque_lqp_r.argtypes = [vp, vp, vp, vp, vp]
    L.dwbc_batch_calc_control_torque_lqp_r_nc.argtypes = [vp, vp, vp, vp, vp]
    L.dwbc_batch_get_lqp.argtypes = [vp, C.c_int, C.c_int, vp, _ip, _ip, vp]
    L.dwbc_batch_get_lqp_iterations.argtypes = [vp, vp, vp]
    L.dwbc_batch_set_task_gain.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.dwbc_batch_set_trajectory_quintic.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp]
    L.dwbc_batch_set_trajectory_rotation.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.dwbc_batch_calc_fstar.argtypes = [vp, C.c_int, vp, vp, vp]
    L.dwbc_batch_update_kinematics_mujoco.argtypes = [vp, vp, vp, vp]
    L.dwbc_batch_get_torque_command_f32.argtypes = [vp, vp, vp]
    L.dwbc_batch_set_outputs.argtypes = [vp, C.c_int]
    L.dwbc_batch_last_path.argtypes = [vp]
    L.dwbc_batch_workspace_bytes.argtypes = [vp]
    L.dwbc_batch_workspace_bytes.restype = C.c_longlong
    L.dwbc_batch_set_lean_timing.argtypes = [vp, vp]
    return L


EXPORTS = [
    "dwbc_last_error", "dwbc_launch_count", "dwbc_model_load_urdf", "dwbc_model_destroy", "dwbc_model_info", "dwbc_model_link_id",
    "dwbc_model_link_name", "dwbc_batch_create", "dwbc_batch_destroy", "dwbc_batch_size", "dwbc_batch_limits", "dwbc_batch_add_contact",
    "dwbc_batch_set_friction", "dwbc_batch_add_task", "dwbc_batch_add_custom_task", "dwbc_batch_clear_tasks", "dwbc_batch_set_torque_limit",
    "dwbc_batch_update_kinematics", "dwbc_batch_set_contact", "dwbc_batch_calc_contact_constraint", "dwbc_batch_set_task", "dwbc_batch_solve",
    "dwbc_batch_get_matrix", "dwbc_batch_get_active_set", "dwbc_batch_cycle_host", "dwbc_probe_fp64", "dwbc_batch_get_qp_iterations",
    "dwbc_batch_set_profiling", "dwbc_batch_get_kernel_times", "dwbc_kernel_name", "dwbc_batch_get_contact_force", "dwbc_batch_get_zmp",
    "dwbc_batch_reduced_dynamics", "dwbc_batch_reduced_calc_contact_constraint", "dwbc_batch_reduced_solve", "dwbc_batch_reduced_dims", "dwbc_batch_set_reduced_torque_limit",
    "dwbc_batch_get_reduced_matrix", "dwbc_batch_calc_control_torque_lqp", "dwbc_batch_calc_control_torque_lqp_r", "dwbc_batch_calc_control_torque_lqp_r_nc", "dwbc_batch_get_lqp", "dwbc_batch_get_lqp_iterations",
    "dwbc_batch_set_task_gain", "dwbc_batch_set_trajectory_quintic", "dwbc_batch_set_trajectory_rotation", "dwbc_batch_calc_fstar",
    "dwbc_batch_update_kinematics_mujoco", "dwbc_batch_get_torque_command_f32", "dwbc_batch_set_host_slices",
    "dwbc_batch_set_outputs", "dwbc_batch_last_path", "dwbc_batch_workspace_bytes", "dwbc_batch_set_lean_timing", "dwbc_batch_cycle_host_levels",
    "dwbc_multi_create", "dwbc_multi_destroy", "dwbc_multi_gpus", "dwbc_multi_shard", "dwbc_multi_add_contact", "dwbc_multi_add_task",
    "dwbc_multi_add_custom_task", "dwbc_multi_set_torque_limit", "dwbc_multi_set_outputs", "dwbc_multi_cycle_host", "dwbc_multi_last_shard_ms",
    "dwbc_host_alloc", "dwbc_host_free", "dwbc_device_alloc", "dwbc_device_free", "dwbc_memcpy_h2d", "dwbc_memcpy_d2h", "dwbc_device_count",
]

_lib = None


def load_library():
    """Loads libdwbc_b200.so (built in-tree by __graft_entry__.build()).  Raises DwbcError if it is missing."""
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise DwbcError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). There is no CPU fallback.")
        try:
            _lib = declare(C.CDLL(path))
        except OSError as e:
            raise DwbcError(f"cannot load {path}: {e}") from e
    return _lib
