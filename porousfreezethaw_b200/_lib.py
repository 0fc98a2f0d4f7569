"""ctypes binding of the C ABI declared in include/pft_b200.h (libpft_b200.so, nvcc sm_100a).

There is no CPU path: if the CUDA library has not been built, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libpft_b200.so")
RK_LIB_PATH = os.path.join(HERE, "csrc", "librk_b200.so")
FAST_LIB_PATH = os.path.join(HERE, "csrc", "libpft_b200_fast.so")      # opt-in FMA variant (not bit-exact)

NPARAM = 30
_dp = C.POINTER(C.c_double)


class PftCase(C.Structure):
    """pft_case (include/pft_b200.h)"""
    _fields_ = [("n1", C.c_int), ("n2", C.c_int), ("n3", C.c_int), ("total_n3", C.c_int),
                ("first_row", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("calc_mode", C.c_int),
                ("L1", C.c_double), ("L2", C.c_double), ("L3", C.c_double),
                ("param", C.c_double * NPARAM), ("max_n3", C.c_int)]


SERVICE_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_double, C.c_double, C.c_long, C.c_long, C.c_void_p)


class PftSystem(C.Structure):
    """pft_system (include/pft_b200.h)"""
    _fields_ = [("t", C.c_double), ("h", C.c_double), ("h_min", C.c_double), ("delta", C.c_double),
                ("delta_local", C.c_int), ("steps", C.c_long), ("steps_total", C.c_long),
                ("service", SERVICE_FN), ("service_ctx", C.c_void_p),
                ("handle_nan", C.c_int), ("last_nan", C.c_int),
                ("max_attempts", C.c_long), ("rhs_evals", C.c_long)]


class PftStats(C.Structure):
    """pft_stats (include/pft_b200.h)"""
    _fields_ = [("kernel_launches", C.c_long), ("rhs_evals", C.c_long), ("attempts", C.c_long), ("accepted", C.c_long),
                ("device_ms_total", C.c_double), ("stage_ms", C.c_double * 5), ("stage_launches", C.c_long * 5),
                ("tile_x", C.c_int), ("tile_y", C.c_int), ("z_chunk", C.c_int), ("threads", C.c_int), ("grid_blocks", C.c_int),
                ("bytes_per_field", C.c_long), ("host_wait_ms", C.c_double), ("host_launch_ms", C.c_double),
                ("device_pow", C.c_int), ("runahead", C.c_int), ("noop_launches", C.c_long),
                ("kernel_generation", C.c_int), ("rows_per_thread", C.c_int),
                ("speculative_stage1", C.c_int), ("reserved_", C.c_int)]


# every symbol include/pft_b200.h declares
EXPORTS = [
    "pft_last_error", "pft_version", "pft_create", "pft_destroy", "pft_upload", "pft_download",
    "pft_upload_interior", "pft_download_interior", "pft_rasterise", "pft_rhs_host", "pft_solve",
    "pft_ice_fraction", "pft_comm_unique_id", "pft_comm_init", "pft_get_stats", "pft_reset_stats",
    "pft_debug_fetch", "pft_debug_pow", "pft_upload_fields", "pft_download_fields", "pft_host_register", "pft_host_unregister",
    "pft_group_create", "pft_group_destroy", "pft_group_size", "pft_group_member", "pft_group_upload", "pft_group_download",
    "pft_group_solve", "pft_group_create_planes", "pft_group_balance", "pft_group_planes",
]

_cache: dict[str, C.CDLL] = {}


def load(path: str | None = None) -> C.CDLL:
    """Load the CUDA library (default: the in-tree nvcc build).  Raises if it is missing."""
    path = path or LIB_PATH
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} not found: the sm_100a CUDA library has not been built "
            f"(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C porousfreezethaw_b200/csrc`). "
            f"porousfreezethaw_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.pft_last_error.restype = C.c_char_p
    lib.pft_version.restype = C.c_char_p
    lib.pft_create.argtypes = [C.POINTER(PftCase), C.c_int, C.c_uint, C.POINTER(vp)]
    lib.pft_destroy.argtypes = [vp]
    lib.pft_upload.argtypes = [vp, _dp]
    lib.pft_download.argtypes = [vp, _dp]
    lib.pft_upload_interior.argtypes = [vp, _dp, _dp, _dp]
    lib.pft_download_interior.argtypes = [vp, _dp, _dp, _dp]
    lib.pft_rasterise.argtypes = [vp, _dp, C.c_int]
    lib.pft_rhs_host.argtypes = [vp, C.c_double, _dp, _dp]
    lib.pft_solve.argtypes = [vp, C.POINTER(PftSystem), C.c_double]
    lib.pft_ice_fraction.argtypes = [vp, _dp]
    lib.pft_comm_unique_id.argtypes = [C.c_void_p]
    lib.pft_comm_init.argtypes = [vp, C.c_void_p]
    lib.pft_get_stats.argtypes = [vp, C.POINTER(PftStats)]
    lib.pft_reset_stats.argtypes = [vp]
    lib.pft_debug_fetch.argtypes = [vp, C.c_int, _dp, _dp]
    lib.pft_debug_pow.argtypes = [vp, _dp, C.c_double, _dp, C.c_int]
    lib.pft_upload_fields.argtypes = [vp, _dp, C.c_uint]
    lib.pft_download_fields.argtypes = [vp, _dp, C.c_uint]
    lib.pft_host_register.argtypes = [C.c_void_p, C.c_ulonglong]
    lib.pft_host_unregister.argtypes = [C.c_void_p]
    lib.pft_group_create.argtypes = [C.POINTER(PftCase), C.c_int, C.POINTER(C.c_int), C.c_uint, C.POINTER(vp)]
    lib.pft_group_destroy.argtypes = [vp]
    lib.pft_group_size.argtypes = [vp]
    lib.pft_group_member.argtypes = [vp, C.c_int]
    lib.pft_group_member.restype = vp
    lib.pft_group_upload.argtypes = [vp, _dp, C.c_uint]
    lib.pft_group_download.argtypes = [vp, _dp, C.c_uint]
    lib.pft_group_solve.argtypes = [vp, C.POINTER(PftSystem), C.c_double]
    lib.pft_group_create_planes.argtypes = [C.POINTER(PftCase), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_uint, C.POINTER(vp)]
    lib.pft_group_balance.argtypes = [vp, C.POINTER(PftSystem), C.c_double, C.c_int]
    lib.pft_group_planes.argtypes = [vp, C.POINTER(C.c_int)]
    for name in EXPORTS:
        getattr(lib, name)
    _cache[path] = lib
    return lib


_rk: dict[str, C.CDLL] = {}


def load_rk(path: str | None = None) -> C.CDLL:
    """librk_b200.so: the drop-in host objects (solver module + host side of the model plug-in)."""
    path = path or RK_LIB_PATH
    if path not in _rk:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found (make -C porousfreezethaw_b200/csrc)")
        if path == RK_LIB_PATH:
            load()                  # libpft_b200.so first: librk links against it
        lib = C.CDLL(path)
        lib.pft_equation_rasterise_host.argtypes = [C.POINTER(PftCase), _dp, C.c_int, _dp]
        _rk[path] = lib
    return _rk[path]


def rasterise_host(case: PftCase, spheres_raw, solution_padded, rk_path: str | None = None) -> None:
    """the plug-in's host rasteriser (libm tanh, bit-identical to equation.c:474-529): gl of a padded array"""
    import numpy as np
    s = np.ascontiguousarray(spheres_raw, dtype=np.float64).reshape(-1, 3)
    rc = load_rk(rk_path).pft_equation_rasterise_host(C.byref(case), s.ctypes.data_as(_dp), s.shape[0], solution_padded.ctypes.data_as(_dp))
    if rc:
        raise RuntimeError(f"pft_equation_rasterise_host failed ({rc})")
