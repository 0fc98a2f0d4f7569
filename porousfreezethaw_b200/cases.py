"""Case definitions: the parameter sets of the reference's shipped/archived `Params` case files
and the synthetic sphere-pack cubes of SURVEY.md section 8(d), as plain numbers.

Host-side mirror of what the reference driver derives from a case file before it calls the
solver (apps/intertrack-hybrid-S-freezing/Params:1-146, intertrack.c:1491-1575 parameter fetch,
:1776-1800 block geometry, :1934-1968 initial-condition node coordinates).  Expressions are
evaluated in float64 in the order the case file states them.

Names follow the reference (n1..n3, L1..L3, calc_mode, tau, tau_min, delta, param[] indices of
model.c:44-59).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# model.c:44-59
PARAM_NAMES = [
    "u_star", "L", "xi", "a", "b", "alpha", "mu",
    "beads_scaling", "beads_offset_x", "beads_offset_y", "beads_offset_z",
    "xi_gl", "zeta", "p_eps0", "p_eps1", "gamma",
    "water_cp", "ice_cp", "glass_cp", "water_lambda", "ice_lambda", "glass_lambda",
    "water_rho", "ice_rho", "glass_rho", "top_temp1", "top_temp2", "phase_switch_time",
    "u_noise_amp", "ball_radius",
]
PIDX = {n: i for i, n in enumerate(PARAM_NAMES)}
NPARAM = len(PARAM_NAMES)


@dataclass
class CaseSpec:
    """Everything the hot path consumes for one case (whole grid, before slab partition)."""
    name: str
    n1: int
    n2: int
    n3: int                      # total_n3
    L1: float
    L2: float
    L3: float
    calc_mode: int
    params: dict
    tau: float = 1.0
    tau_min: float = 1e-6
    delta: float = 1e-3
    final_time: float = 36000.0
    saved_files: int = 100
    spheres: np.ndarray | None = None          # raw (unit) centres, shape (count, 3)
    ic_cap: tuple = (0.052, 0.058)             # z-range of the initial ice cap (Params:11)
    ic_lid_z: float = 0.055                    # glass lid position (Params:18 / archived cases)
    ic_p_zero: bool = False                    # mode-2 cases start without ice (Cases-*/…-Temp/Params)
    notes: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def cells(self) -> int:
        return self.n1 * self.n2 * self.n3

    def param_array(self) -> np.ndarray:
        a = np.zeros(NPARAM)
        for k, v in self.params.items():
            a[PIDX[k]] = v
        return a

    def snapshot_time(self, snapshot: int, t0: float = 0.0, s0: int = 0) -> float:
        """intertrack.c:2271"""
        return t0 + ((self.final_time - t0) * (snapshot - s0)) / (self.saved_files - 1 - s0)

    def initial_fields(self, first_row: int = 0, n3_local: int | None = None):
        """Initial u, p, gl (before the sphere pack is max-ed into gl) on rows
        [first_row, first_row + n3_local) -- node coordinates as intertrack.c:1937-1951:
        x = L1*((0.5+i)/n1), z = L3*((0.5+k+first_row)/total_n3)."""
        n3l = self.n3 if n3_local is None else n3_local
        x = self.L1 * ((0.5 + np.arange(self.n1)) / self.n1)
        y = self.L2 * ((0.5 + np.arange(self.n2)) / self.n2)
        z = self.L3 * ((0.5 + np.arange(n3l) + first_row) / self.n3)
        Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
        u = np.full((n3l, self.n2, self.n1), 293.15)
        if self.ic_p_zero:
            p = np.zeros_like(u)
        else:
            z0, z1 = self.ic_cap
            p = ((Z > z0) & (Z < z1) & ((X - self.L1 / 2) ** 2 + (Y - self.L2 / 2) ** 2 < (self.L1 / 3) ** 2)).astype(np.float64)
        xi_gl = self.params["xi_gl"]
        gl = 0.5 * (1.0 + np.tanh(0.5 / xi_gl * (Z - self.ic_lid_z)))
        return u, p, gl


def _base_params(L1, L3, xi_div, xi_gl_div, zeta, alpha_mult, gamma=2.0):
    """Physical constants common to every archived case (Params:54-101), with the per-case
    overrides seen in results/*/Cases-*.tgz!*/Params."""
    hours = 60.0 * 60.0
    water_rho, water_cp = 997.0, 4.18e3
    beads_scaling = L1
    alpha = water_rho * water_cp
    if alpha_mult != 1:
        alpha = alpha * alpha_mult
    return dict(
        u_noise_amp=0.0,
        water_cp=water_cp, ice_cp=2.05e3, glass_cp=0.84e3,
        water_lambda=0.6, ice_lambda=2.22, glass_lambda=1.1,
        water_rho=water_rho, ice_rho=917.0, glass_rho=2500.0,
        u_star=273.15, L=3.34e5,
        xi_gl=L3 / xi_gl_div, zeta=zeta,
        beads_scaling=beads_scaling, ball_radius=0.1 * beads_scaling,
        beads_offset_x=0.0, beads_offset_y=0.0, beads_offset_z=0.0,
        xi=L3 / xi_div, a=2.0, b=1.0, alpha=alpha, mu=1e-4,
        p_eps0=0.05, p_eps1=0.2, gamma=gamma,
        top_temp1=273.15 - 25, top_temp2=273.15 + 20, phase_switch_time=5 * hours,
    )


def _grid(grid_nodes, L1=0.03, L2=0.03, L3=0.06):
    """Params:136-144: grid_nodes along the longest side."""
    mult = grid_nodes / max(L1, L2, L3)
    return int(round(L1 * mult)), int(round(L2 * mult)), int(round(L3 * mult))


def archived_case(name: str) -> CaseSpec:
    """The archived production cases named in BASELINE.json / SURVEY.md section 8(d).

    name: '100-GradP' | '100-Temp' | '200-GradP' | '200-SigmaP1-P' | '200-Temp' |
          '2025-200-GradP' | '2025-200-SigmaP1-P' | '2025-200-Temp-Gamma=10' |
          '400-SigmaP1-P-smallsigma' | '400-Temp'
    (sphere centres are NOT attached here; see `with_spheres`)."""
    L1, L2, L3 = 0.03, 0.03, 0.06
    hours = 3600.0
    table = {
        # name: (grid_nodes, calc_mode, xi_div, xi_gl_div, zeta, alpha_mult, gamma)
        "100-GradP": (100, 0, 100, 300, 2.0, 1, 2.0),          # Cases-LR.tgz!freeze-thaw-10h-GradP/Params
        "100-SigmaP1-P": (100, 1, 100, 300, 2.0, 1, 2.0),
        "100-Temp": (100, 2, 100, 300, 1.0, 1, 2.0),
        "200-GradP": (200, 0, 200, 300, 2.0, 1, 2.0),          # Cases-MR.tgz!freeze-thaw-10h-GradP/Params
        "200-SigmaP1-P": (200, 1, 200, 300, 2.0, 1, 2.0),
        "200-Temp": (200, 2, 200, 300, 1.0, 1, 2.0),
        "2025-200-GradP": (200, 0, 200, 1000, 2.0, 1, 2.0),    # …Computational_costs.tgz!…-1node/freeze-thaw-10h-GradP/Params
        "2025-200-SigmaP1-P": (200, 1, 200, 1000, 2.0, 1, 2.0),
        "2025-200-Temp-Gamma=10": (200, 2, 200, 1000, 1.0, 1, 10.0),
        "400-SigmaP1-P-smallsigma": (400, 1, 400, 300, 2.0, 3, 2.0),   # Cases-HR.tgz!…-SigmaP1-P-smallsigma/Params
        "400-Temp": (400, 2, 400, 300, 1.0, 1, 2.0),
    }
    if name not in table:
        raise KeyError(f"unknown case {name!r}; known: {sorted(table)}")
    g, mode, xi_div, xi_gl_div, zeta, amult, gamma = table[name]
    n1, n2, n3 = _grid(g)
    return CaseSpec(name=name, n1=n1, n2=n2, n3=n3, L1=L1, L2=L2, L3=L3, calc_mode=mode,
                    params=_base_params(L1, L3, xi_div, xi_gl_div, zeta, amult, gamma),
                    tau=1.0, tau_min=1e-6, delta=1e-3, final_time=10 * hours, saved_files=100,
                    ic_p_zero=(mode == 2),
                    notes="archived case parameters (values differ from the shipped Params only where the archived Params do)")


def synthetic_spheres(count: int, seed: int = 20251018, zmax: float = 1.0) -> np.ndarray:
    """Seeded uniform random unit-cube centres; overlaps allowed."""
    rng = np.random.default_rng(seed)
    s = rng.random((count, 3))
    s[:, 2] *= zmax
    return s


def cube_case(n: int, nz: int | None = None, calc_mode: int = 0, spheres: int = 4000) -> CaseSpec:
    """Synthetic random sphere-pack cube of SURVEY.md 8(d)-5: L = 0.06 m, n x n x nz nodes
    (nz defaults to n), radius 0.03*L, xi_gl = L/1000, the case-1 cap scaled to the cube."""
    L = 0.06
    nz = n if nz is None else nz
    Lz = L * nz / n
    par = _base_params(L, L, 200, 1000, 2.0, 1)
    par.update(beads_scaling=L, ball_radius=0.03 * L)
    cs = CaseSpec(name=f"cube-{n}x{n}x{nz}", n1=n, n2=n, n3=nz, L1=L, L2=L, L3=Lz, calc_mode=calc_mode,
                  params=par, tau=1.0, tau_min=1e-6, delta=1e-3, final_time=36000.0,
                  ic_cap=(0.87 * Lz, 0.97 * Lz), ic_lid_z=0.92 * Lz,
                  notes="synthetic sphere pack, numpy default_rng(20251018)")
    cs.spheres = synthetic_spheres(spheres, zmax=Lz / L)
    return cs


def small_case(n1=12, n2=10, n3=16, calc_mode=0, spheres: int = 6, seed: int = 7) -> CaseSpec:
    """Tiny case for parity tests: coarse grid with proportionally thick interfaces so every
    term of the RHS is exercised (ice cap, glass lid, a few beads)."""
    L1, L2, L3 = 0.03 * n1 / 12, 0.03 * n2 / 12, 0.06 * n3 / 24
    par = _base_params(L1, L3, 8, 10, 2.0 if calc_mode != 2 else 1.0, 1, gamma=2.0)
    par.update(beads_scaling=L1, ball_radius=0.22 * L1)
    cs = CaseSpec(name=f"small-{n1}x{n2}x{n3}-m{calc_mode}", n1=n1, n2=n2, n3=n3, L1=L1, L2=L2, L3=L3,
                  calc_mode=calc_mode, params=par, tau=1.0, tau_min=1e-6, delta=1e-3, final_time=600.0,
                  ic_cap=(0.55 * L3, 0.9 * L3), ic_lid_z=0.85 * L3, ic_p_zero=(calc_mode == 2))
    rng = np.random.default_rng(seed)
    s = rng.random((spheres, 3))
    s[:, 1] *= L2 / L1
    s[:, 2] *= 0.6 * L3 / L1
    cs.spheres = s
    return cs


def with_spheres(cs: CaseSpec, raw_centres: np.ndarray) -> CaseSpec:
    cs.spheres = np.ascontiguousarray(raw_centres, dtype=np.float64).reshape(-1, 3)
    return cs


def partition(total_n3: int, nranks: int, rank: int):
    """z-slab partition, intertrack.c:1780-1787: total_n3 // nranks rows each, the remainder one
    row each to the lowest ranks.  Returns (n3_local, first_row)."""
    q, r = divmod(total_n3, nranks)
    if rank < r:
        return q + 1, rank * q + rank
    return q, rank * q + r


def balanced_planes(planes, cost):
    """Planes per slab in inverse proportion to the measured cost per plane (`cost[r]` = device time per attempt of slab r
    with `planes[r]` planes), every slab keeps >= 2 planes, the sum is exact -- the rule of pft_group_balance
    (csrc/pft_b200.cu: balanced_planes), for callers that own one slab per process."""
    n, total = len(planes), int(sum(planes))
    w = [planes[r] / cost[r] for r in range(n)]
    ws = sum(w)
    out = [max(2, int(math.floor(total * w[r] / ws))) for r in range(n)]
    given = sum(out)
    while given != total:
        best, bestv = -1, 0.0
        for r in range(n):
            frac = total * w[r] / ws - out[r]
            v = frac if given < total else -frac
            if (given < total or out[r] > 2) and (best < 0 or v > bestv):
                best, bestv = r, v
        if best < 0:
            break
        out[best] += 1 if given < total else -1
        given += 1 if given < total else -1
    return out
