"""Host-side handle on one z-slab of a case living in one B200's HBM.

Thin Python over the C ABI (include/pft_b200.h); the names mirror the reference's solver-module
interface (include/RK_MPI_SAsolver.h: init / solve / check_NAN, RK_MPI_S_SOLUTION's t, h, h_min,
delta, steps, steps_total) so the parity tests read like calls into the reference.

Used by tests/ and bench.py.  The C drop-in for the reference application itself is
csrc/rk_mpi_sa_b200.c + csrc/equation_b200.c.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .cases import CaseSpec, partition

BT = 2
NVAR = 3

FLAG_NO_RUNAHEAD = 2
FLAG_RECOMPUTE_K1 = 4
FLAG_LEGACY_LOADS = 8
FLAG_HOST_POW = 32

FIELD_U, FIELD_P, FIELD_GL, FIELD_ALL = 1, 2, 4, 7

SOLVE_DONE, SOLVE_INTERRUPTED, SOLVE_BUDGET = 0, 1, 2

_dp = C.POINTER(C.c_double)


def _ptr(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "float64 C-contiguous array required"
    return a.ctypes.data_as(_dp)


class PftError(RuntimeError):
    pass


@dataclass
class SolveResult:
    code: int
    t: float
    h: float
    steps: int
    steps_total: int
    rhs_evals: int
    last_nan: int


class SlabSolver:
    """One block (z-slab) of the grid on one GPU: the device-resident counterpart of the
    reference's (solution array, RK_MPI_S_SOLUTION) pair."""

    def __init__(self, spec: CaseSpec, rank: int = 0, nranks: int = 1, device: int = 0, flags: int = 0,
                 lib_path: str | None = None, planes=None):
        """planes: planes per slab for all ranks (sum == spec.n3); None = the reference's rule (intertrack.c:1780-1787).
        Results do not depend on it; slower slabs can be given fewer planes (cases.balanced_planes)."""
        self.lib = _lib.load(lib_path)
        self.spec = spec
        self.rank, self.nranks = rank, nranks
        if planes is None:
            self.n3, self.first_row = partition(spec.n3, nranks, rank)
            deepest = 0
        else:
            assert len(planes) == nranks and sum(planes) == spec.n3 and min(planes) >= 1
            self.n3, self.first_row, deepest = int(planes[rank]), int(sum(planes[:rank])), int(max(planes))
        self.n1, self.n2 = spec.n1, spec.n2
        cc = _lib.PftCase(spec.n1, spec.n2, self.n3, spec.n3, self.first_row, rank, nranks, spec.calc_mode,
                          spec.L1, spec.L2, spec.L3)
        cc.max_n3 = deepest
        par = spec.param_array()
        for i in range(_lib.NPARAM):
            cc.param[i] = float(par[i])
        self._case = cc
        h = C.c_void_p()
        rc = self.lib.pft_create(C.byref(cc), device, flags, C.byref(h))
        if rc:
            raise PftError(f"pft_create failed ({rc}): {self.lib.pft_last_error().decode()}")
        self._h = h
        self.sys = _lib.PftSystem()
        self.sys.t, self.sys.h = 0.0, spec.tau
        self.sys.h_min, self.sys.delta = spec.tau_min, spec.delta
        self._service_keep = None

    # -- geometry of the reference's padded host layout (intertrack.c:1776-1800) -----------
    @property
    def N1(self): return self.n1 + 2 * BT
    @property
    def N2(self): return self.n2 + 2 * BT
    @property
    def N3(self): return self.n3 + 2 * BT
    @property
    def SIZE(self): return self.N1 * self.N2 * self.N3
    @property
    def cells(self): return self.n1 * self.n2 * self.n3

    def _check(self, rc, what):
        if rc:
            raise PftError(f"{what} failed ({rc}): {self.lib.pft_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pft_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data movement ---------------------------------------------------------------------
    def upload(self, x_padded: np.ndarray):
        assert x_padded.size == NVAR * self.SIZE
        self._check(self.lib.pft_upload(self._h, _ptr(x_padded)), "pft_upload")

    def download(self, x_padded: np.ndarray):
        assert x_padded.size == NVAR * self.SIZE
        self._check(self.lib.pft_download(self._h, _ptr(x_padded)), "pft_download")

    def upload_fields(self, x_padded: np.ndarray, mask: int):
        assert x_padded.size == NVAR * self.SIZE
        self._check(self.lib.pft_upload_fields(self._h, _ptr(x_padded), mask), "pft_upload_fields")

    def download_fields(self, x_padded: np.ndarray, mask: int):
        assert x_padded.size == NVAR * self.SIZE
        self._check(self.lib.pft_download_fields(self._h, _ptr(x_padded), mask), "pft_download_fields")

    def upload_interior(self, u=None, p=None, gl=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.float64) for a in (u, p, gl)]
        for a in arrs:
            assert a is None or a.size == self.cells
        self._check(self.lib.pft_upload_interior(self._h, *[_ptr(a) for a in arrs]), "pft_upload_interior")

    def download_interior(self):
        u, p, gl = (np.empty((self.n3, self.n2, self.n1)) for _ in range(3))
        self._check(self.lib.pft_download_interior(self._h, _ptr(u), _ptr(p), _ptr(gl)), "pft_download_interior")
        return u, p, gl

    def debug_fetch(self, which: int):
        u, p = (np.empty((self.n3, self.n2, self.n1)) for _ in range(2))
        self._check(self.lib.pft_debug_fetch(self._h, which, _ptr(u), _ptr(p)), "pft_debug_fetch")
        return u, p

    def debug_pow(self, x: np.ndarray, y: float = 0.2) -> np.ndarray:
        """x**y computed by the DEVICE instance of the step-size controller's pow twin (csrc/pft_pow.h)"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        self._check(self.lib.pft_debug_pow(self._h, _ptr(x), float(y), _ptr(out), x.size), "pft_debug_pow")
        return out

    def rasterise(self, spheres_raw: np.ndarray):
        s = np.ascontiguousarray(spheres_raw, dtype=np.float64).reshape(-1, 3)
        self._check(self.lib.pft_rasterise(self._h, _ptr(s), s.shape[0]), "pft_rasterise")

    def ice_fraction(self) -> float:
        out = C.c_double()
        self._check(self.lib.pft_ice_fraction(self._h, C.byref(out)), "pft_ice_fraction")
        return out.value

    # -- the hot path ----------------------------------------------------------------------
    def rhs_host(self, t: float, w_padded: np.ndarray) -> np.ndarray:
        """dw = f(t, w) with host buffers in the padded layout: RK_RightHandSide (RK_MPI_SAsolver.h:173)."""
        dw = np.zeros_like(w_padded)
        self._check(self.lib.pft_rhs_host(self._h, t, _ptr(w_padded), _ptr(dw)), "pft_rhs_host")
        return dw

    def set_service(self, fn):
        """fn(final_time, t, h, steps, steps_total) -> int; nonzero interrupts (Service_Callback)."""
        if fn is None:
            self.sys.service = _lib.SERVICE_FN()
            self._service_keep = None
            return
        cb = _lib.SERVICE_FN(lambda ft, t, h, s, st, ctx: int(fn(ft, t, h, s, st)))
        self._service_keep = cb
        self.sys.service = cb

    def solve(self, final_time: float, max_attempts: int = 0) -> SolveResult:
        """RK_MPI_SA_solve(final_time, &system) on the device-resident state."""
        self.sys.max_attempts = max_attempts
        rc = self.lib.pft_solve(self._h, C.byref(self.sys), final_time)
        if rc not in (0, 1, 2, -4):
            raise PftError(f"pft_solve failed ({rc}): {self.lib.pft_last_error().decode()}")
        s = self.sys
        return SolveResult(rc, s.t, s.h, s.steps, s.steps_total, s.rhs_evals, s.last_nan)

    # -- multi-GPU -------------------------------------------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._check(self.lib.pft_comm_unique_id(buf), "pft_comm_unique_id")
        return buf.raw

    def comm_init(self, uid: bytes):
        assert len(uid) == 128
        buf = C.create_string_buffer(uid, 128)
        self._check(self.lib.pft_comm_init(self._h, buf), "pft_comm_init")

    def stats(self) -> dict:
        st = _lib.PftStats()
        self._check(self.lib.pft_get_stats(self._h, C.byref(st)), "pft_get_stats")
        d = {n: getattr(st, n) for n, _ in st._fields_ if n not in ("stage_ms", "stage_launches")}
        d["stage_ms"] = list(st.stage_ms)
        d["stage_launches"] = list(st.stage_launches)
        return d

    def reset_stats(self):
        self._check(self.lib.pft_reset_stats(self._h), "pft_reset_stats")


class GroupSolver:
    """The whole block of a case on several GPUs of ONE host process (pft_group_*): what the reference does
    with M MPI ranks inside one application run.  Host data is the one padded array of the whole block."""

    def __init__(self, spec: CaseSpec, ngpus: int, devices=None, flags: int = 0, lib_path: str | None = None, planes=None):
        """planes: planes per slab (sum == spec.n3); None = the reference's rule, rebalanced once by the library
        after the first 30 attempts (pft_group_balance) unless PFT_BALANCE=0"""
        self.lib = _lib.load(lib_path)
        self.spec = spec
        self.ngpus = ngpus
        self.n1, self.n2, self.n3 = spec.n1, spec.n2, spec.n3
        cc = _lib.PftCase(spec.n1, spec.n2, spec.n3, spec.n3, 0, 0, 1, spec.calc_mode, spec.L1, spec.L2, spec.L3)
        par = spec.param_array()
        for i in range(_lib.NPARAM):
            cc.param[i] = float(par[i])
        dev = (C.c_int * ngpus)(*(devices if devices is not None else range(ngpus)))
        h = C.c_void_p()
        pl = None if planes is None else (C.c_int * ngpus)(*[int(v) for v in planes])
        rc = self.lib.pft_group_create_planes(C.byref(cc), ngpus, dev, pl, flags, C.byref(h))
        if rc:
            raise PftError(f"pft_group_create failed ({rc}): {self.lib.pft_last_error().decode()}")
        self._h = h
        self.sys = _lib.PftSystem()
        self.sys.t, self.sys.h = 0.0, spec.tau
        self.sys.h_min, self.sys.delta = spec.tau_min, spec.delta
        self._service_keep = None

    @property
    def SIZE(self): return (self.n1 + 2 * BT) * (self.n2 + 2 * BT) * (self.n3 + 2 * BT)

    def _check(self, rc, what):
        if rc:
            raise PftError(f"{what} failed ({rc}): {self.lib.pft_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self.lib.pft_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, x_padded: np.ndarray, mask: int = FIELD_ALL):
        assert x_padded.size == NVAR * self.SIZE
        self._check(self.lib.pft_group_upload(self._h, _ptr(x_padded), mask), "pft_group_upload")

    def download(self, x_padded: np.ndarray, mask: int = FIELD_ALL):
        assert x_padded.size == NVAR * self.SIZE
        self._check(self.lib.pft_group_download(self._h, _ptr(x_padded), mask), "pft_group_download")

    set_service = SlabSolver.set_service

    def solve(self, final_time: float, max_attempts: int = 0) -> SolveResult:
        self.sys.max_attempts = max_attempts
        rc = self.lib.pft_group_solve(self._h, C.byref(self.sys), final_time)
        if rc not in (0, 1, 2, -4):
            raise PftError(f"pft_group_solve failed ({rc}): {self.lib.pft_last_error().decode()}")
        s = self.sys
        return SolveResult(rc, s.t, s.h, s.steps, s.steps_total, s.rhs_evals, s.last_nan)

    def planes(self):
        out = (C.c_int * self.ngpus)()
        self.lib.pft_group_planes(self._h, out)
        return list(out)

    def balance(self, final_time: float, attempts: int = 30) -> int:
        """pft_group_balance: calibrate over `attempts` attempts and deal the planes out again; 1 if the partition changed"""
        rc = self.lib.pft_group_balance(self._h, C.byref(self.sys), final_time, attempts)
        if rc < 0:
            raise PftError(f"pft_group_balance failed ({rc}): {self.lib.pft_last_error().decode()}")
        return rc

    def member_stats(self, r: int) -> dict:
        st = _lib.PftStats()
        self._check(self.lib.pft_get_stats(self.lib.pft_group_member(self._h, r), C.byref(st)), "pft_get_stats")
        d = {n: getattr(st, n) for n, _ in st._fields_ if n not in ("stage_ms", "stage_launches")}
        d["stage_ms"] = list(st.stage_ms)
        d["stage_launches"] = list(st.stage_launches)
        return d
