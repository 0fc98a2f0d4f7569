"""The reference-facing seam driven from Python: what `intertrack.c` does around one `RK_MPI_SA_solve` call
(apps/intertrack-hybrid-S-freezing/intertrack.c:2131-2283), through librk_b200.so -- the six `RK_MPI_SA_*`
entry points of include/RK_MPI_SAsolver.h and the host side of the model plug-in (pft_equation_*).

Used by bench.py (`e2e_dropin`: the end-to-end figure through the solver-module ABI with a host-resident `x`,
on one GPU or -- PFT_NGPUS -- on several GPUs driven by this one process) and by the tests.  The layout of the
structures is the reference's (tests/test_abi.py checks sizeof/offsetof against the reference header).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .cases import CaseSpec

BT = 2
DELTA_LOCAL, DELTA_GLOBAL = 0, 1
_dp = C.POINTER(C.c_double)


class RKMemDist(C.Structure):
    """RK_MEM_DIST (include/RK_MPI_SAsolver.h:184-194)"""
    _fields_ = [("n_chunks", C.c_int), ("chunk_start", C.POINTER(C.c_int)), ("chunk_size", C.POINTER(C.c_int)),
                ("chunk_eps_mult", _dp)]


META_F = C.CFUNCTYPE(C.c_void_p)
SERVICE_CB = C.CFUNCTYPE(C.c_int, C.c_double, C.c_void_p)


class RKSolution(C.Structure):
    """RK_MPI_S_SOLUTION (include/RK_MPI_SAsolver.h:196-289)"""
    _fields_ = [("n", C.POINTER(RKMemDist)), ("t", C.c_double), ("x", _dp), ("meta_f", META_F), ("h", C.c_double),
                ("h_min", C.c_double), ("delta", C.c_double), ("delta_mode", C.c_int), ("DDLBF_Rearrange", C.c_void_p),
                ("Service_Callback", SERVICE_CB), ("steps", C.c_long), ("steps_total", C.c_long)]


class DropInRun:
    """One single-rank application run on the drop-in: configure the plug-in for the whole grid, build the
    chunk table of intertrack.c:2149-2156 (one chunk per x-row of every variable), hand the padded host
    array to RK_MPI_SA_solve."""

    def __init__(self, spec: CaseSpec, ngpus: int = 1, rk_path: str | None = None):
        self.rk_path = rk_path
        self.rk = _lib.load_rk(rk_path)
        rk = self.rk
        rk.pft_equation_configure.argtypes = [C.POINTER(_lib.PftCase)]
        rk.pft_equation_release.restype = None
        rk.RK_MPI_SA_init.argtypes = [C.c_int, C.c_int, C.c_int]
        rk.RK_MPI_SA_solve.argtypes = [C.c_double, C.POINTER(RKSolution)]
        self.spec = spec
        n1, n2, n3 = spec.n1, spec.n2, spec.n3
        self.N1, self.N2, self.N3 = n1 + 2 * BT, n2 + 2 * BT, n3 + 2 * BT
        self.SIZE = self.N1 * self.N2 * self.N3
        cc = _lib.PftCase(n1, n2, n3, n3, 0, 0, 1, spec.calc_mode, spec.L1, spec.L2, spec.L3)
        par = spec.param_array()
        for i in range(_lib.NPARAM):
            cc.param[i] = float(par[i])
        self.case = cc
        old = os.environ.get("PFT_NGPUS")
        os.environ["PFT_NGPUS"] = str(ngpus)
        try:
            rc = rk.pft_equation_configure(C.byref(cc))
        finally:
            if old is None:
                os.environ.pop("PFT_NGPUS", None)
            else:
                os.environ["PFT_NGPUS"] = old
        if rc:
            raise RuntimeError(f"pft_equation_configure failed ({rc})")
        rc = rk.RK_MPI_SA_init(3 * self.SIZE, 0, 0)
        if rc:
            rk.pft_equation_release()
            raise RuntimeError(f"RK_MPI_SA_init failed ({rc})")
        # intertrack.c:2149-2156
        rowsize = self.N1 * self.N2
        q, k, j = np.meshgrid(np.arange(3), np.arange(n3), np.arange(n2), indexing="ij")
        self._start = np.ascontiguousarray((q * self.SIZE + (k + BT) * rowsize + (j + BT) * self.N1 + BT).reshape(-1), dtype=np.int32)
        self._size = np.full(self._start.size, n1, dtype=np.int32)
        self._mult = np.ones(self._start.size)
        self._dist = RKMemDist(self._start.size, self._start.ctypes.data_as(C.POINTER(C.c_int)),
                               self._size.ctypes.data_as(C.POINTER(C.c_int)), self._mult.ctypes.data_as(_dp))
        fn = rk.pft_f_device_model2 if spec.calc_mode == 2 else rk.pft_f_device_model01
        addr = C.cast(fn, C.c_void_p).value
        self._meta = META_F(lambda: addr)
        self.system = RKSolution()
        self.system.n = C.pointer(self._dist)
        self.system.meta_f = self._meta
        self.system.t, self.system.h = 0.0, spec.tau
        self.system.h_min, self.system.delta, self.system.delta_mode = spec.tau_min, spec.delta, DELTA_GLOBAL
        self.x = None

    def rasterise_host(self, x_padded: np.ndarray):
        _lib.rasterise_host(self.case, self.spec.spheres, x_padded, self.rk_path)

    def solve(self, x_padded: np.ndarray, final_time: float) -> int:
        """RK_MPI_SA_solve(final_time, &system) with system->x = the caller's padded host array (updated in place)"""
        assert x_padded.dtype == np.float64 and x_padded.size == 3 * self.SIZE and x_padded.flags["C_CONTIGUOUS"]
        self.x = x_padded
        self.system.x = x_padded.ctypes.data_as(_dp)
        return self.rk.RK_MPI_SA_solve(final_time, C.byref(self.system))

    def close(self):
        if getattr(self, "rk", None) is not None:
            self.rk.RK_MPI_SA_cleanup()
            self.rk.pft_equation_release()
            self.rk = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
