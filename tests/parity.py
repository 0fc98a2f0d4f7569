"""Parity checks shared by the GPU tests (product library on a B200) and the CPU emulation tests
(tests/cuda_emu: same sources executed by fibers).  Every check compares the CUDA path, called
through the C ABI, with the oracle (oracle/pft_oracle.c) on the same seeded inputs.

Bar: bit-exact (np.array_equal) for calc_mode 0, 1, 10, 11 -- strict IEEE association, no FMA,
host-side glibc pow in the controller.  calc_mode 2 uses cosh(), whose device implementation
differs from glibc in the last ulp, so on a GPU the mode-2 bar is rel. L_inf <= 1e-10 on the
fields and +-1% on step counts; under emulation mode 2 is bit-exact too.
"""
from __future__ import annotations

import numpy as np

from oracle import pyoracle as po
from porousfreezethaw_b200 import cases
from porousfreezethaw_b200.solver import SlabSolver

MODE2_RTOL = 1e-10


def make(spec: cases.CaseSpec):
    """oracle-side case + initial padded array (IC + rasterised sphere pack)."""
    case = po.Case(spec.n1, spec.n2, spec.n3, spec.L1, spec.L2, spec.L3, spec.param_array(), spec.calc_mode)
    u, p, gl = spec.initial_fields()
    w = case.pack(u, p, gl)
    orc = po.Oracle(case)
    if spec.spheres is not None and len(spec.spheres):
        orc.rasterise(spec.spheres, w)
    return case, orc, w


def rel_linf(a, b):
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def assert_fields(case, got, want, exact: bool, what: str):
    a, b = case.interior(got), case.interior(want)
    if exact:
        assert np.array_equal(a, b), f"{what}: not bit-identical, max|diff| u={np.abs(a[0]-b[0]).max():.3e} p={np.abs(a[1]-b[1]).max():.3e}"
    else:
        for q, name in ((0, "u"), (1, "p")):
            r = rel_linf(a[q], b[q])
            assert r <= MODE2_RTOL, f"{what}: {name} rel L_inf {r:.3e} > {MODE2_RTOL}"
        assert np.array_equal(a[2], b[2]), f"{what}: gl changed"


def check_rhs(spec, lib_path, exact=True, t=0.0, flags=0, perturb=True):
    """K = f(t, w) through the RK_RightHandSide-shaped entry point (pft_rhs_host)."""
    case, orc, w = make(spec)
    if perturb == "tiny":
        # the far tail of the phase field ahead of an interface: |p| from 1e-200 down through the
        # subnormals to exact zeros, both signs, under a (mostly) uniform temperature -- tiny and
        # subnormal numerators and quotients in every division of the right-hand side
        rng = np.random.default_rng(12)
        W = w.reshape(3, case.N3, case.N2, case.N1)
        shape = (case.n3, case.n2, case.n1)
        mag = 10.0 ** rng.uniform(-332, -200, size=shape)          # < 4.9e-324 underflows to 0
        W[1, 2:-2, 2:-2, 2:-2] = mag * rng.choice([-1.0, 1.0], size=shape)
        W[0, 2:-2, 2:-2, 2:-2] = 272.0
        W[0, 2:-2, 2:-2, 2:-2][rng.random(shape) < 0.02] += 1.0
        assert np.count_nonzero((np.abs(W[1]) > 0) & (np.abs(W[1]) < 2.3e-308)) > 0
    elif perturb:   # a smooth non-trivial temperature / phase state so that every term is live
        rng = np.random.default_rng(11)
        W = w.reshape(3, case.N3, case.N2, case.N1)
        W[0, 2:-2, 2:-2, 2:-2] += rng.uniform(-30, 5, size=(case.n3, case.n2, case.n1))
        W[1, 2:-2, 2:-2, 2:-2] = np.clip(W[1, 2:-2, 2:-2, 2:-2] * 0.8 + rng.uniform(0, 0.3, size=(case.n3, case.n2, case.n1)), 0, 1)
    want = orc.rhs(t, w.copy())
    sol = SlabSolver(spec, lib_path=lib_path, flags=flags)
    try:
        got = sol.rhs_host(t, w.copy())
    finally:
        sol.close()
    a, b = case.interior(got), case.interior(want)
    if exact:
        assert np.array_equal(a, b), f"RHS not bit-identical: max|diff| du={np.abs(a[0]-b[0]).max():.3e} dp={np.abs(a[1]-b[1]).max():.3e}"
    else:
        for q in (0, 1):
            assert rel_linf(a[q], b[q]) <= MODE2_RTOL
    assert not a[2].any()


def oracle_stage_arrays(case, orc, w, t, h):
    """K1, aux3, K3, K4, x_new, eps of ONE Merson attempt, formed with numpy in the reference's
    association (RK_MPI_SAsolver_hybrid2.c:373-453, 521, 667)."""
    h2, h3, h6, h8 = h / 2.0, h / 3.0, h / 6.0, h / 8.0
    I = np.zeros(w.size, dtype=bool).reshape(3, case.N3, case.N2, case.N1)
    I[:, 2:-2, 2:-2, 2:-2] = True
    I = I.reshape(-1)
    x = w.copy()
    K1 = orc.rhs(t, x)
    aux = x.copy(); aux[I] = K1[I] * h3 + x[I]
    K2 = orc.rhs(t + h3, aux)
    aux3 = x.copy(); aux3[I] = (K1[I] + K2[I]) * h6 + x[I]
    K3 = orc.rhs(t + h3, aux3.copy())
    aux = x.copy(); aux[I] = (K1[I] + 3.0 * K3[I]) * h8 + x[I]
    K4 = orc.rhs(t + h2, aux)
    aux = x.copy(); aux[I] = (0.5 * K1[I] - 1.5 * K3[I] + 2.0 * K4[I]) * h + x[I]
    K5 = orc.rhs(t + h, aux)
    e = np.abs(0.2 * K1[I] - 0.9 * K3[I] + 0.8 * K4[I] - 0.1 * K5[I])
    xn = x.copy(); xn[I] = x[I] + h3 * (0.5 * (K1[I] + K5[I]) + 2.0 * K4[I])
    return dict(K1=K1, aux3=aux3, K3=K3, K4=K4, xn=xn, eps=float(e.max()))


def check_fused_stages(spec, lib_path, exact=True, h=0.37, flags=0):
    """One forced-accept attempt (h_min = inf) and every intermediate array the fused stage
    kernels leave in HBM, against the unfused numpy/oracle formation."""
    case, orc, w = make(spec)
    want = oracle_stage_arrays(case, orc, w, 0.0, h)
    sol = SlabSolver(spec, lib_path=lib_path, flags=flags)
    try:
        sol.upload(w.copy())
        sol.sys.t, sol.sys.h, sol.sys.h_min = 0.0, h, float("inf")
        r = sol.solve(0.0 + h)
        assert r.code == 0 and r.steps == 1 and r.steps_total == 1
        # (K1 is not checked here: after an accepted step it already holds f(t+h, x_new), ready
        #  for the next attempt; check_rhs covers stage 1)
        names = {2: "aux3", 3: "K3", 4: "K4", 0: "xn"}
        for which, name in names.items():
            gu, gp = sol.debug_fetch(which)
            ref = case.interior(want[name])
            for q, g in ((0, gu), (1, gp)):
                if spec.calc_mode in (10, 11) and q == 0 and which != 0:
                    continue
                if exact:
                    assert np.array_equal(g, ref[q]), f"{name}[{'up'[q]}] differs: max|diff| {np.abs(g-ref[q]).max():.3e}"
                else:
                    assert rel_linf(g, ref[q]) <= MODE2_RTOL, f"{name}[{'up'[q]}]"
    finally:
        sol.close()


def check_fixed_steps(spec, lib_path, nsteps=5, H=0.25, exact=True, flags=0):
    """Fixed-step mode as SURVEY.md 8(d) defines it on the unmodified solver: h = H, h_min = +inf,
    solve(t + H) -> exactly one always-accepted Merson step with h_eff = fl(fl(t+H) - t)."""
    case, orc, w = make(spec)
    x = w.copy()
    s = orc.new_system(t=0.0, h=H, h_min=float("inf"), delta=spec.delta)
    work = np.zeros(5 * x.size)
    sol = SlabSolver(spec, lib_path=lib_path, flags=flags)
    try:
        sol.upload(w.copy())
        sol.sys.h_min = float("inf")
        for _ in range(nsteps):
            s.h = H
            assert orc.solve(s, x, s.t + H, work) == 0
            sol.sys.h = H
            r = sol.solve(sol.sys.t + H)
            assert r.code == 0
            assert r.t == s.t
        assert sol.sys.steps == nsteps == s.steps and sol.sys.steps_total == s.steps_total
        got = w.copy()
        sol.download(got)
        assert_fields(case, got, x, exact, f"{nsteps} fixed steps")
    finally:
        sol.close()


def check_adaptive(spec, lib_path, T, exact=True, flags=0, chunks=1, with_service=False):
    """Adaptive mode: identical accepted-step sequence (t_i, h_i), identical steps / steps_total,
    identical final fields and hand-over step size -- over `chunks` consecutive solve calls
    (the snapshot loop of intertrack.c:2265-2283)."""
    case, orc, w = make(spec)
    x = w.copy()
    rec_o, rec_g = [], []
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta,
                       service=(lambda ft, t, h, st, tot, ctx: rec_o.append((t, h)) or 0) if with_service else None)
    work = np.zeros(5 * x.size)
    sol = SlabSolver(spec, lib_path=lib_path, flags=flags)
    try:
        sol.upload(w.copy())
        if with_service:
            sol.set_service(lambda ft, t, h, st, tot: rec_g.append((t, h)) or 0)
        for c in range(1, chunks + 1):
            Tc = T * c / chunks
            assert orc.solve(s, x, Tc, work) == 0
            r = sol.solve(Tc)
            assert r.code == 0
            if exact:
                assert (r.t, r.h, r.steps, r.steps_total) == (s.t, s.h, s.steps, s.steps_total), \
                    f"chunk {c}: gpu {(r.t, r.h, r.steps, r.steps_total)} oracle {(s.t, s.h, s.steps, s.steps_total)}"
            else:
                assert r.t == s.t
                assert abs(r.steps_total - s.steps_total) <= max(1, 0.01 * s.steps_total)
        if exact and with_service:
            assert rec_g == rec_o
        got = w.copy()
        sol.download(got)
        assert_fields(case, got, x, exact, f"adaptive to T={T}")
        return r
    finally:
        sol.close()
