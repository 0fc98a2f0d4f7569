"""One host process driving several slabs (pft_group_*, the in-process form of the reference's multi-rank run:
intertrack.c:1776-1800 slab partition, equation.c:290-326 halo exchange, RK_MPI_SAsolver_hybrid2.c:572 error
maximum) on the CPU emulation of the CUDA sources: one worker thread per emulated device, so the peer-memory
protocol of csrc/pft_kernels.cuh -- boundary planes stored into the neighbour's ghost planes, halo flags and
tickets, the all-to-all of per-slab error maxima inside stage 5, stage 1 skipped after a rejected attempt --
is stepped for real and checked against the single-rank oracle, bit for bit."""
import numpy as np
import pytest

import parity
from porousfreezethaw_b200 import cases
from porousfreezethaw_b200.solver import GroupSolver, FLAG_HOST_POW, FLAG_NO_RUNAHEAD


def _run(spec, ngpus, emu_lib, T, chunks=2, flags=0, service=False, budget=0):
    case, orc, w = parity.make(spec)
    x = w.copy()
    rec_o, rec_g = [], []
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta,
                       service=(lambda ft, t, h, st, tot, ctx: rec_o.append((t, h)) or 0) if service else None)
    work = np.zeros(5 * x.size)
    g = GroupSolver(spec, ngpus, lib_path=emu_lib, flags=flags)
    try:
        g.upload(w.copy())
        if service:
            g.set_service(lambda ft, t, h, st, tot: rec_g.append((t, h)) or 0)
        for c in range(1, chunks + 1):
            Tc = T * c / chunks
            s.max_attempts = budget
            ro = orc.solve(s, x, Tc, work)
            r = g.solve(Tc, max_attempts=budget)
            assert r.code == ro
            assert (r.t, r.h, r.steps, r.steps_total) == (s.t, s.h, s.steps, s.steps_total), f"chunk {c}"
        assert rec_g == rec_o
        got = w.copy()
        g.download(got)
        parity.assert_fields(case, got, x, True, f"{ngpus} slabs")
        # ghost shell of the host array untouched by the download
        mask = np.ones(got.size, dtype=bool).reshape(3, case.N3, case.N2, case.N1)
        mask[:, 2:-2, 2:-2, 2:-2] = False
        assert np.array_equal(got.reshape(mask.shape)[mask], w.reshape(mask.shape)[mask])
        return r
    finally:
        g.close()


@pytest.mark.parametrize("ngpus,mode", [(2, 0), (3, 1), (4, 0), (2, 11), (5, 10)])
def test_group_adaptive_matches_oracle(emu_lib, ngpus, mode):
    r = _run(cases.small_case(12, 10, 19, calc_mode=mode), ngpus, emu_lib, T=10.0)
    if mode in (0, 1):
        assert r.steps_total > r.steps      # the run contains rejected attempts (stage 1 skipped, flags still advance)


def test_group_with_service_callback(emu_lib):
    _run(cases.small_case(12, 10, 16, calc_mode=0), 3, emu_lib, T=8.0, service=True)


@pytest.mark.parametrize("flags", [FLAG_HOST_POW, FLAG_NO_RUNAHEAD])
def test_group_variants(emu_lib, flags):
    _run(cases.small_case(12, 10, 16, calc_mode=1), 2, emu_lib, T=8.0, flags=flags)


@pytest.mark.parametrize("env", [{"PFT_SPEC": "0"}, {"PFT_DATAFLOW": "0"}])
def test_group_classic_launch_order(emu_lib, monkeypatch, env):
    """the peer-memory protocol in the classic order S1..S5 (halo sequence 5*aseq + stage, stage 1 skipped after a
    rejected attempt) -- the default is the speculative order S2..S5 S1', which every other test of this file runs"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = GroupSolver(cases.small_case(12, 10, 16, calc_mode=1), 2, lib_path=emu_lib)
    assert g.member_stats(0)["speculative_stage1"] == 0
    g.close()
    r = _run(cases.small_case(12, 10, 19, calc_mode=1), 3, emu_lib, T=10.0, service=True)
    assert r.steps_total > r.steps
    _run(cases.small_case(12, 10, 16, calc_mode=0), 2, emu_lib, T=400.0, chunks=3, budget=7)


def test_group_speculative_order_is_default(emu_lib):
    g = GroupSolver(cases.small_case(12, 10, 16, calc_mode=1), 3, lib_path=emu_lib)
    try:
        assert all(g.member_stats(r)["speculative_stage1"] == 1 for r in range(3))
    finally:
        g.close()


def test_group_budget_resume(emu_lib):
    _run(cases.small_case(12, 10, 16, calc_mode=0), 2, emu_lib, T=400.0, chunks=3, budget=7)


def test_group_break_by_callback(emu_lib):
    """slab 0's worker asks the callback, every slab stops after the same accepted step"""
    spec = cases.small_case(12, 10, 16, calc_mode=0)
    case, orc, w = parity.make(spec)
    x = w.copy()
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta, service=lambda ft, t, h, st, tot, ctx: int(st % 5 == 0))
    work = np.zeros(5 * x.size)
    g = GroupSolver(spec, 3, lib_path=emu_lib)
    try:
        g.upload(w.copy())
        g.set_service(lambda ft, t, h, st, tot: int(st % 5 == 0))
        for _ in range(4):
            ro = orc.solve(s, x, 30.0, work)
            r = g.solve(30.0)
            assert r.code == ro
            assert (r.t, r.h, r.steps, r.steps_total) == (s.t, s.h, s.steps, s.steps_total)
        got = w.copy()
        g.download(got)
        parity.assert_fields(case, got, x, True, "after breaks")
    finally:
        g.close()


@pytest.mark.parametrize("ngpus,mode", [(1, 0), (2, 1), (3, 0)])
def test_solver_module_abi_from_python(emu_lib, ngpus, mode):
    """RK_MPI_SA_init / RK_MPI_SA_solve of librk (the reference's solver-module ABI) driven the way intertrack.c
    drives it -- chunk table of x-rows, meta_f, host-resident padded x updated in place -- on 1..3 emulated
    devices (PFT_NGPUS): state and fields bit-identical to the oracle; the static glass field goes up once"""
    import os
    from porousfreezethaw_b200.dropin import DropInRun
    rk_emu = os.path.join(os.path.dirname(emu_lib), "librk_b200_emu.so")
    spec = cases.small_case(12, 10, 16, calc_mode=mode)
    case, orc, w = parity.make(spec)
    x = w.copy()
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta)
    work = np.zeros(5 * x.size)
    run = DropInRun(spec, ngpus=ngpus, rk_path=rk_emu)
    try:
        xh = w.copy()
        for T in (3.0, 7.0, 7.5):
            assert orc.solve(s, x, T, work) == 0
            assert run.solve(xh, T) == 0
            sy = run.system
            assert (sy.t, sy.h, sy.steps, sy.steps_total) == (s.t, s.h, s.steps, s.steps_total)
            assert np.array_equal(xh, x) or np.array_equal(case.interior(xh), case.interior(x))
        parity.assert_fields(case, xh, x, True, "through RK_MPI_SA_solve")
    finally:
        run.close()


def test_host_right_hand_side_is_refused(emu_lib):
    """A caller that hands the solver module a HOST right-hand side -- what the reference's DEM application does with
    this ABI (apps/sphere-collider/spheres_friction_angular.c:573,606) -- is refused with the reference's error code
    -2 (bad system); there is no CPU path.  Same for a DDLBF rearrangement hook and weighted chunks."""
    import ctypes as C
    import os
    from porousfreezethaw_b200.dropin import DropInRun, META_F
    rk_emu = os.path.join(os.path.dirname(emu_lib), "librk_b200_emu.so")
    spec = cases.small_case(12, 10, 16, calc_mode=0)
    case, orc, w = parity.make(spec)
    run = DropInRun(spec, rk_path=rk_emu)
    try:
        RHS = C.CFUNCTYPE(None, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double))
        host_rhs = RHS(lambda t, x, dx: None)
        host_addr = C.cast(host_rhs, C.c_void_p).value
        host_meta = META_F(lambda: host_addr)             # (kept alive in this frame: ctypes does not hold it for the struct)
        run.system.meta_f = host_meta
        xh = w.copy()
        assert run.solve(xh, 1.0) == -2
        assert np.array_equal(xh, w)                      # nothing was touched
        run.system.meta_f = run._meta
        run.system.DDLBF_Rearrange = 1
        assert run.solve(xh, 1.0) == -2
        run.system.DDLBF_Rearrange = None
        run._mult[3] = 2.0
        assert run.solve(xh, 1.0) == -2
        run._mult[3] = 1.0
        assert run.solve(xh, 1.0) == 0
    finally:
        run.close()


@pytest.mark.parametrize("planes", [[3, 9, 7], [10, 2, 2, 5], [2, 17]])
def test_group_uneven_partition_is_bit_exact(emu_lib, planes):
    """results do not depend on how the planes are dealt out to the slabs"""
    spec = cases.small_case(12, 10, sum(planes), calc_mode=0)
    case, orc, w = parity.make(spec)
    x = w.copy()
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta)
    work = np.zeros(5 * x.size)
    g = GroupSolver(spec, len(planes), lib_path=emu_lib, planes=planes)
    try:
        assert g.planes() == planes
        g.upload(w.copy())
        for T in (4.0, 9.0):
            assert orc.solve(s, x, T, work) == 0
            r = g.solve(T)
            assert (r.code, r.t, r.h, r.steps, r.steps_total) == (0, s.t, s.h, s.steps, s.steps_total)
        got = w.copy()
        g.download(got)
        parity.assert_fields(case, got, x, True, "uneven slabs")
    finally:
        g.close()


@pytest.mark.parametrize("service", [False, True])
def test_group_rebalances_itself(emu_lib, monkeypatch, service):
    """pft_group_solve calibrates over its first 30 attempts and deals the planes out again (here with given costs:
    the emulation has no device clock); the integration goes on from the migrated state, bit for bit, also across
    a caller's attempt budget and with a callback"""
    monkeypatch.setenv("PFT_BALANCE_COSTS", "1.0,1.6,1.2")
    spec = cases.small_case(12, 10, 21, calc_mode=1)
    case, orc, w = parity.make(spec)
    x = w.copy()
    rec_o, rec_g = [], []
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta,
                       service=(lambda ft, t, h, st, tot, ctx: rec_o.append((t, h)) or 0) if service else None)
    work = np.zeros(5 * x.size)
    g = GroupSolver(spec, 3, lib_path=emu_lib)
    try:
        assert g.planes() == [7, 7, 7]
        g.upload(w.copy())
        if service:
            g.set_service(lambda ft, t, h, st, tot: rec_g.append((t, h)) or 0)
        for T, budget in ((400.0, 12), (400.0, 25), (400.0, 0)):
            s.max_attempts = budget
            ro = orc.solve(s, x, T, work)
            r = g.solve(T, max_attempts=budget)
            assert r.code == ro
            assert (r.t, r.h, r.steps, r.steps_total) == (s.t, s.h, s.steps, s.steps_total)
        assert s.steps_total > 40
        assert g.planes() != [7, 7, 7] and sum(g.planes()) == 21 and g.planes()[1] < 7
        assert rec_g == rec_o
        got = w.copy()
        g.download(got)
        parity.assert_fields(case, got, x, True, "after rebalancing")
    finally:
        g.close()


def test_balanced_planes_rule():
    assert cases.balanced_planes([100, 100, 100, 100], [317, 357, 332, 312]) == [104, 92, 99, 105]
    assert sum(cases.balanced_planes([50] * 8, [1, 1.2, 1.1, 1.3, 1.0, 1.05, 1.15, 0.9])) == 400
    assert cases.balanced_planes([7, 7, 7], [1.0, 1.0, 1.0]) == [7, 7, 7]
    assert min(cases.balanced_planes([3, 3, 3], [1.0, 50.0, 1.0])) == 2
