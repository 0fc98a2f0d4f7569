"""ctypes bindings for the CHECKERS: oracle/libpft_oracle.so (our C restatement) and
oracle/_ref/libpft_ref*.so (the unmodified reference behind oracle/ref_harness.c).

TEST INFRASTRUCTURE ONLY.  Import this module only from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs.
The product package (porousfreezethaw_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BT = 2
NVAR = 3
NPARAM = 30

# parameter indices: model.c:44-59
PARAM_NAMES = [
    "u_star", "L", "xi", "a", "b", "alpha", "mu",
    "beads_scaling", "beads_offset_x", "beads_offset_y", "beads_offset_z",
    "xi_gl", "zeta", "p_eps0", "p_eps1", "gamma",
    "water_cp", "ice_cp", "glass_cp", "water_lambda", "ice_lambda", "glass_lambda",
    "water_rho", "ice_rho", "glass_rho", "top_temp1", "top_temp2", "phase_switch_time",
    "u_noise_amp", "ball_radius",
]
PIDX = {n: i for i, n in enumerate(PARAM_NAMES)}
assert len(PARAM_NAMES) == NPARAM

_dp = C.POINTER(C.c_double)


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class _CCase(C.Structure):
    _fields_ = [("n1", C.c_int), ("n2", C.c_int), ("n3", C.c_int), ("total_n3", C.c_int),
                ("first_row", C.c_int), ("rank", C.c_int), ("nranks", C.c_int), ("calc_mode", C.c_int),
                ("L1", C.c_double), ("L2", C.c_double), ("L3", C.c_double),
                ("param", C.c_double * NPARAM)]


HALO_FN = C.CFUNCTYPE(None, _dp, C.c_void_p)
ALLMAX_FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
SERVICE_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_double, C.c_double, C.c_long, C.c_long, C.c_void_p)


class _CSystem(C.Structure):
    _fields_ = [("t", C.c_double), ("h", C.c_double), ("h_min", C.c_double), ("delta", C.c_double),
                ("delta_local", C.c_int), ("steps", C.c_long), ("steps_total", C.c_long),
                ("halo", HALO_FN), ("halo_ctx", C.c_void_p),
                ("allmax", ALLMAX_FN), ("allmax_ctx", C.c_void_p),
                ("service", SERVICE_FN), ("service_ctx", C.c_void_p),
                ("handle_nan", C.c_int), ("last_nan", C.c_int),
                ("max_attempts", C.c_long), ("rhs_evals", C.c_long)]


def build(target: str = "all") -> None:
    """(Re)build the checkers with oracle/Makefile (building the checker is not using it)."""
    subprocess.run(["make", "-C", HERE, target], check=True, stdout=subprocess.DEVNULL)


_oracle_lib = None


def load_oracle():
    global _oracle_lib
    if _oracle_lib is None:
        path = os.path.join(HERE, "libpft_oracle.so")
        if not os.path.exists(path):
            build("oracle")
        lib = C.CDLL(path)
        lib.pfo_sizes.argtypes = [C.POINTER(_CCase), C.POINTER(C.c_int)]
        lib.pfo_partition.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.pfo_read_spheres.argtypes = [C.c_char_p, _dp, C.c_int]
        lib.pfo_read_spheres.restype = C.c_int
        lib.pfo_rasterise.argtypes = [C.POINTER(_CCase), _dp, C.c_int, _dp]
        lib.pfo_bcond.argtypes = [C.POINTER(_CCase), C.c_double, _dp]
        lib.pfo_rhs.argtypes = [C.POINTER(_CCase), C.c_double, _dp, _dp, HALO_FN, C.c_void_p]
        lib.pfo_solve.argtypes = [C.POINTER(_CCase), C.POINTER(_CSystem), _dp, _dp, C.c_double]
        lib.pfo_solve.restype = C.c_int
        lib.pfo_ice_fraction.argtypes = [C.POINTER(_CCase), _dp]
        lib.pfo_ice_fraction.restype = C.c_double
        lib.pfo_pow_array.argtypes = [_dp, C.c_double, _dp, C.c_long]
        _oracle_lib = lib
    return _oracle_lib


def libm_pow(x: np.ndarray, y: float) -> np.ndarray:
    """x**y with the C library's pow(), element by element (NOT numpy's vectorised power)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    load_oracle().pfo_pow_array(_ptr(x), float(y), _ptr(out), x.size)
    return out


def ref_path(variant: str = "hybrid2") -> str:
    name = {"hybrid2": "libpft_ref.so", "hybrid": "libpft_ref_hybrid.so", "fast": "libpft_ref_fast.so"}[variant]
    return os.path.join(HERE, "_ref", name)


def have_ref(variant: str = "hybrid2") -> bool:
    return os.path.exists(ref_path(variant))


def spheres_file() -> str | None:
    """The reference's 200-sphere pack (data/spheres_positions.txt), staged by oracle/Makefile
    into oracle/_ref/data/ (git-ignored; not product source)."""
    p = os.path.join(HERE, "_ref", "data", "spheres_positions.txt")
    return p if os.path.exists(p) else None


# --------------------------------------------------------------------------------------

@dataclass
class Case:
    """Geometry + parameters of one block (names follow intertrack.c)."""
    n1: int
    n2: int
    total_n3: int
    L1: float
    L2: float
    L3: float
    param: np.ndarray
    calc_mode: int = 0
    rank: int = 0
    nranks: int = 1
    n3: int = field(default=-1)
    first_row: int = field(default=-1)

    def __post_init__(self):
        self.param = np.ascontiguousarray(self.param, dtype=np.float64)
        assert self.param.shape == (NPARAM,)
        if self.n3 < 0:
            self.n3, self.first_row = partition(self.total_n3, self.nranks, self.rank)

    # reference layout helpers -----------------------------------------------------
    @property
    def N1(self): return self.n1 + 2 * BT
    @property
    def N2(self): return self.n2 + 2 * BT
    @property
    def N3(self): return self.n3 + 2 * BT
    @property
    def rowsize(self): return self.N1 * self.N2
    @property
    def SIZE(self): return self.rowsize * self.N3
    @property
    def cells(self): return self.n1 * self.n2 * self.n3

    def slab(self, rank: int, nranks: int) -> "Case":
        return Case(self.n1, self.n2, self.total_n3, self.L1, self.L2, self.L3, self.param.copy(),
                    self.calc_mode, rank, nranks)

    def c(self) -> _CCase:
        cc = _CCase(self.n1, self.n2, self.n3, self.total_n3, self.first_row, self.rank, self.nranks,
                    self.calc_mode, self.L1, self.L2, self.L3)
        for i in range(NPARAM):
            cc.param[i] = float(self.param[i])
        return cc

    def pack(self, u, p, gl) -> np.ndarray:
        """interior [n3][n2][n1] fields -> padded reference array (ghosts zero)."""
        w = np.zeros((NVAR, self.N3, self.N2, self.N1))
        for q, f in enumerate((u, p, gl)):
            w[q, BT:-BT, BT:-BT, BT:-BT] = np.asarray(f, dtype=np.float64).reshape(self.n3, self.n2, self.n1)
        return w.reshape(-1)

    def interior(self, w: np.ndarray) -> np.ndarray:
        """padded reference array -> [3][n3][n2][n1] copy of the interior."""
        return np.ascontiguousarray(w.reshape(NVAR, self.N3, self.N2, self.N1)[:, BT:-BT, BT:-BT, BT:-BT])


def partition(total_n3: int, nranks: int, rank: int):
    """intertrack.c:1780-1787 (pure-python twin of pfo_partition, used by host-side tests)."""
    q, r = divmod(total_n3, nranks)
    if rank < r:
        return q + 1, rank * q + rank
    return q, rank * q + r


def read_spheres(path: str, max_count: int = 1000) -> np.ndarray:
    lib = load_oracle()
    buf = np.zeros((max_count, 3))
    n = lib.pfo_read_spheres(path.encode(), _ptr(buf), max_count)
    if n < 0:
        raise FileNotFoundError(path)
    return buf[:n].copy()


class Oracle:
    """Our C restatement bound to one Case."""

    def __init__(self, case: Case):
        self.case = case
        self.lib = load_oracle()
        self.cc = case.c()
        self._keep = []

    def rasterise(self, spheres_raw: np.ndarray, w: np.ndarray) -> None:
        s = np.ascontiguousarray(spheres_raw, dtype=np.float64).reshape(-1, 3)
        self.lib.pfo_rasterise(C.byref(self.cc), _ptr(s), s.shape[0], _ptr(w))

    def bcond(self, t: float, w: np.ndarray) -> None:
        self.lib.pfo_bcond(C.byref(self.cc), t, _ptr(w))

    def rhs(self, t: float, w: np.ndarray, halo=None) -> np.ndarray:
        dw = np.zeros_like(w)
        cb = HALO_FN(halo) if halo else HALO_FN()
        self.lib.pfo_rhs(C.byref(self.cc), t, _ptr(w), _ptr(dw), cb, None)
        return dw

    def ice_fraction(self, w: np.ndarray) -> float:
        return self.lib.pfo_ice_fraction(C.byref(self.cc), _ptr(w))

    def new_system(self, t=0.0, h=1.0, h_min=1e-6, delta=1e-3, delta_local=False, handle_nan=False,
                   halo=None, allmax=None, service=None, max_attempts=0) -> _CSystem:
        s = _CSystem()
        s.t, s.h, s.h_min, s.delta = t, h, h_min, delta
        s.delta_local = int(delta_local)
        s.handle_nan = int(handle_nan)
        s.max_attempts = max_attempts
        if halo:
            s.halo = HALO_FN(halo)
        if allmax:
            s.allmax = ALLMAX_FN(allmax)
        if service:
            s.service = SERVICE_FN(service)
        self._keep.append(s)
        return s

    def solve(self, s: _CSystem, x: np.ndarray, final_time: float, work: np.ndarray | None = None) -> int:
        if work is None:
            work = np.zeros(5 * x.size)
        return self.lib.pfo_solve(C.byref(self.cc), C.byref(s), _ptr(x), _ptr(work), final_time)


class Reference:
    """The unmodified reference (single rank) behind oracle/ref_harness.c.  One instance at a
    time per variant: the reference keeps its state in file-static globals."""

    def __init__(self, variant: str = "hybrid2"):
        path = ref_path(variant)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} (build with `make -C oracle ref` where /root/reference exists)")
        lib = C.CDLL(path)
        lib.pftref_setup.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _dp, C.c_int,
                                     C.c_char_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
        lib.pftref_dims.argtypes = [C.POINTER(C.c_int)]
        lib.pftref_solution.restype = _dp
        lib.pftref_rhs.argtypes = [C.c_double, _dp, _dp]
        lib.pftref_bcond.argtypes = [C.c_double, _dp]
        lib.pftref_solve.argtypes = [C.c_double]
        lib.pftref_state.argtypes = [_dp]
        lib.pftref_set_state.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double]
        lib.pftref_handle_nan.argtypes = [C.c_int]
        lib.pftref_break_after.argtypes = [C.c_long]
        lib.pftref_record_count.restype = C.c_long
        lib.pftref_record_get.argtypes = [_dp, C.c_long]
        lib.pftref_record_get.restype = C.c_long
        self.lib = lib
        self.case = None

    def setup(self, case: Case, u, p, gl, spheres_path: str | None, t0=0.0, tau=1.0, tau_min=1e-6, delta=1e-3,
              use_callback=False) -> None:
        assert case.nranks == 1
        self.case = case
        u = np.ascontiguousarray(u, dtype=np.float64)
        p = np.ascontiguousarray(p, dtype=np.float64)
        gl = np.ascontiguousarray(gl, dtype=np.float64)
        par = np.ascontiguousarray(case.param)
        if spheres_path is None:
            # the reference aborts the process when the sphere file is missing (equation.c:491-499):
            # hand it an empty pack (gl is then taken as given)
            import tempfile
            f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
            f.close()
            spheres_path = f.name
        rc = self.lib.pftref_setup(case.n1, case.n2, case.total_n3, case.L1, case.L2, case.L3, _ptr(par), case.calc_mode,
                                   spheres_path.encode(), _ptr(u), _ptr(p), _ptr(gl),
                                   t0, tau, tau_min, delta, int(use_callback))
        if rc:
            raise RuntimeError(f"pftref_setup failed: {rc}")

    def solution(self) -> np.ndarray:
        """live view of the reference's padded solution array"""
        n = NVAR * self.case.SIZE
        return np.ctypeslib.as_array(self.lib.pftref_solution(), shape=(n,))

    def rhs(self, t: float, w: np.ndarray | None = None) -> np.ndarray:
        dw = np.zeros(NVAR * self.case.SIZE)
        self.lib.pftref_rhs(t, _ptr(w) if w is not None else None, _ptr(dw))
        return dw

    def bcond(self, t: float, w: np.ndarray) -> None:
        self.lib.pftref_bcond(t, _ptr(w))

    def solve(self, final_time: float) -> int:
        return self.lib.pftref_solve(final_time)

    def state(self):
        out = np.zeros(4)
        self.lib.pftref_state(_ptr(out))
        return float(out[0]), float(out[1]), int(out[2]), int(out[3])

    def set_state(self, t, h, h_min, delta):
        self.lib.pftref_set_state(t, h, h_min, delta)

    def record(self) -> np.ndarray:
        n = self.lib.pftref_record_count()
        out = np.zeros((max(n, 1), 2))
        n = self.lib.pftref_record_get(_ptr(out), n)
        return out[:n].copy()

    def close(self):
        self.lib.pftref_cleanup()
