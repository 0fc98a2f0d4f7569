"""GPU parity tests (run on a B200 with -m gpu): the nvcc-built sm_100a library, called through
the C ABI, against the oracle on identical seeded inputs.

Bit-exact for calc_mode 0/1/10/11; calc_mode 2 (device cosh) within rel. L_inf 1e-10 / +-1% steps.
Full-size cases (200x200x400 = BASELINE's "400" case) are checked through size-independent
properties: tiling/chunking invariance, flux-sharing on/off, K1 reuse on/off, interrupted vs
uninterrupted integration -- all bit-exact -- plus conservation-type sanity."""
import os

import numpy as np
import pytest

import parity
from oracle import pyoracle as po
from porousfreezethaw_b200 import cases
from porousfreezethaw_b200.solver import (SlabSolver, GroupSolver, FLAG_NO_RUNAHEAD, FLAG_RECOMPUTE_K1, FLAG_LEGACY_LOADS,
                                          FLAG_HOST_POW)

pytestmark = pytest.mark.gpu


def _real_spheres(spec):
    f = po.spheres_file()
    if f:
        spec.spheres = np.loadtxt(f).reshape(-1, 3)
    else:
        spec.spheres = cases.synthetic_spheres(200, zmax=1.6)
    return spec


@pytest.mark.parametrize("mode", [0, 1, 10, 11])
def test_rhs_bit_exact(gpu_lib, mode):
    parity.check_rhs(cases.small_case(24, 20, 32, calc_mode=mode), gpu_lib)


@pytest.mark.parametrize("mode", [0, 1, 11])
def test_rhs_tiny_and_subnormal_values_bit_exact(gpu_lib, mode):
    """the far tail of the phase field (1e-200 ... subnormal ... 0): tiny and subnormal numerators and
    quotients in every division of the right-hand side keep their bits (no flush-to-zero anywhere)"""
    parity.check_rhs(cases.small_case(24, 20, 32, calc_mode=mode), gpu_lib, perturb="tiny")


def test_rhs_mode2_tolerance(gpu_lib):
    parity.check_rhs(cases.small_case(24, 20, 32, calc_mode=2), gpu_lib, exact=False)


@pytest.mark.parametrize("mode,flags", [(0, 0), (0, FLAG_LEGACY_LOADS), (1, 0), (1, FLAG_LEGACY_LOADS), (10, 0), (11, 0)])
def test_fused_stages_bit_exact(gpu_lib, mode, flags):
    parity.check_fused_stages(cases.small_case(24, 20, 32, calc_mode=mode), gpu_lib, flags=flags)


def test_fused_stages_mode2(gpu_lib):
    parity.check_fused_stages(cases.small_case(24, 20, 32, calc_mode=2), gpu_lib, exact=False)


@pytest.mark.parametrize("tx,ty,zc", [(5, 4, 3), (32, 8, 2), (7, 9, 5), (64, 4, 1)])
def test_ragged_tiles_and_chunks(gpu_lib, monkeypatch, tx, ty, zc):
    """grids that are not multiples of the tile, many tiles, several z-chunks (edge cases)"""
    monkeypatch.setenv("PFT_TX", str(tx))
    monkeypatch.setenv("PFT_TY", str(ty))
    monkeypatch.setenv("PFT_ZCHUNKS", str(zc))
    spec = cases.small_case(37, 23, 19, calc_mode=0)
    parity.check_rhs(spec, gpu_lib)
    parity.check_fused_stages(spec, gpu_lib)


def test_minimal_grids(gpu_lib):
    """smallest sensible blocks: one row / one plane thick"""
    for n in ((8, 8, 1), (8, 1, 8), (1, 8, 8), (3, 3, 3)):
        spec = cases.small_case(*n, calc_mode=1, spheres=2)
        parity.check_rhs(spec, gpu_lib)
        parity.check_fixed_steps(spec, gpu_lib, nsteps=2)


@pytest.mark.parametrize("mode", [0, 1, 10, 11])
def test_fixed_step_mode(gpu_lib, mode):
    parity.check_fixed_steps(cases.small_case(24, 20, 32, calc_mode=mode), gpu_lib, nsteps=10)


@pytest.mark.parametrize("mode", [0, 1, 10, 11])
def test_adaptive_sequence_bit_exact(gpu_lib, mode):
    parity.check_adaptive(cases.small_case(24, 20, 32, calc_mode=mode), gpu_lib, T=60.0, chunks=3, with_service=True)


@pytest.mark.parametrize("flags", [FLAG_NO_RUNAHEAD, FLAG_HOST_POW, FLAG_RECOMPUTE_K1, FLAG_RECOMPUTE_K1 | FLAG_HOST_POW, FLAG_LEGACY_LOADS])
def test_adaptive_variants(gpu_lib, flags):
    """the device-resident controller with attempts enqueued ahead (default), one attempt at a time, with
    pow() taken from the host C library instead of its device twin, with K1 recomputed like the reference:
    all the same sequence, bit for bit"""
    parity.check_adaptive(cases.small_case(24, 20, 32, calc_mode=0), gpu_lib, T=40.0, chunks=2, flags=flags)


def test_device_controller_is_on(gpu_lib):
    """the default path must be the device controller (host pow is the fallback when the twin of the C
    library's pow cannot be verified at start-up)"""
    sol = SlabSolver(cases.small_case(12, 10, 16), lib_path=gpu_lib)
    st = sol.stats()
    sol.close()
    assert st["device_pow"] == 1 and st["runahead"] > 1


def test_device_pow_twin_matches_host_libm(gpu_lib):
    """pow(x, 0.2) of the DEVICE instance of csrc/pft_pow.h against the host C library's pow() (called from C:
    oracle/pft_oracle.c:pfo_pow_array), bit for bit, on 4e7 operands: every positive bit pattern, the
    controller's working range of delta/eps, the neighbourhood of 1, subnormals, 0 and inf"""
    sol = SlabSolver(cases.small_case(12, 10, 16), lib_path=gpu_lib)
    rng = np.random.default_rng(5)
    n = 10_000_000
    one = np.uint64(0x3ff0000000000000)
    xs = [rng.integers(0, 0x7ff0000000000000, size=n, dtype=np.uint64).view(np.float64),
          rng.uniform(0.75, 1.5, size=n), 10.0 ** rng.uniform(-3, 3.6, size=n),
          (one + rng.integers(0, 4096, size=n).astype(np.uint64) - np.uint64(2048)).view(np.float64),
          np.array([0.0, np.inf, 5e-324, 1e-310, 2.2250738585072014e-308, 1.7976931348623157e308, 1.0])]
    try:
        for x in xs:
            got = sol.debug_pow(x, 0.2)
            want = po.libm_pow(x, 0.2)
            bad = np.count_nonzero(got.view(np.uint64) != want.view(np.uint64))
            assert bad == 0, f"{bad} of {x.size} operands differ from the host pow"
    finally:
        sol.close()


def test_adaptive_mode2(gpu_lib):
    parity.check_adaptive(cases.small_case(24, 20, 32, calc_mode=2), gpu_lib, T=60.0, exact=False)


def test_case100_first_steps(gpu_lib):
    """configs[0]: the "100" low-resolution case (50x50x100) with the reference sphere pack:
    the first attempts of the real run (t = 0 -> 2 s), bit-exact sequence and fields."""
    spec = _real_spheres(cases.archived_case("100-GradP"))
    parity.check_adaptive(spec, gpu_lib, T=2.0, chunks=2, with_service=True)


def test_case200_rhs_and_steps(gpu_lib):
    """configs[1]: the "200" medium-resolution case (100x100x200)"""
    spec = _real_spheres(cases.archived_case("200-GradP"))
    parity.check_rhs(spec, gpu_lib, perturb=False)
    parity.check_fixed_steps(spec, gpu_lib, nsteps=3, H=0.01)


def test_dirichlet_switch(gpu_lib):
    """the z-rear Dirichlet temperature switches at phase_switch_time (equation.c:110): stage times
    straddling the switch must pick the value per stage"""
    spec = cases.small_case(16, 12, 20, calc_mode=0)
    spec.params["phase_switch_time"] = 3.0
    parity.check_adaptive(spec, gpu_lib, T=8.0, chunks=2)


def test_callback_break_and_resume(gpu_lib):
    """Service_Callback returning nonzero -> solve returns 1 with (t, h=new_h) saved
    (RK_MPI_SAsolver_hybrid2.c:697-705); resuming reaches the same final state as the oracle
    driven the same way."""
    spec = cases.small_case(24, 20, 32, calc_mode=0)
    case, orc, w = parity.make(spec)
    x = w.copy()
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta,
                       service=lambda ft, t, h, st, tot, ctx: int(st % 7 == 0))
    sol = SlabSolver(spec, lib_path=gpu_lib)
    sol.upload(w.copy())
    sol.set_service(lambda ft, t, h, st, tot: int(st % 7 == 0))
    codes_o, codes_g = [], []
    for _ in range(6):
        codes_o.append(orc.solve(s, x, 25.0))
        r = sol.solve(25.0)
        codes_g.append(r.code)
        assert (r.t, r.h, r.steps, r.steps_total) == (s.t, s.h, s.steps, s.steps_total)
    assert codes_o == codes_g and 1 in codes_g
    got = w.copy(); sol.download(got); sol.close()
    parity.assert_fields(case, got, x, True, "callback break/resume")


def test_attempt_budget_resume(gpu_lib):
    """max_attempts (bench extension): stopping every 5 attempts and resuming == one uninterrupted run"""
    spec = cases.small_case(24, 20, 32, calc_mode=1)
    case, orc, w = parity.make(spec)
    a = SlabSolver(spec, lib_path=gpu_lib); a.upload(w.copy())
    ra = a.solve(40.0)
    b = SlabSolver(spec, lib_path=gpu_lib); b.upload(w.copy())
    while True:
        rb = b.solve(40.0, max_attempts=5)
        if rb.code != 2:
            break
    assert (ra.t, ra.h, ra.steps, ra.steps_total) == (rb.t, rb.h, rb.steps, rb.steps_total)
    xa, xb = w.copy(), w.copy()
    a.download(xa); b.download(xb); a.close(); b.close()
    assert np.array_equal(xa, xb)


def test_nan_handling(gpu_lib):
    """RK_MPI_SA_handle_NAN semantics (hybrid2.c:464-504, 626-646): a step so long that the error
    estimate overflows is retried with h/10 and last_NAN is raised"""
    spec = cases.small_case(16, 12, 20, calc_mode=0)
    case, orc, w = parity.make(spec)
    x = w.copy()
    s = orc.new_system(t=0.0, h=1e200, h_min=0.0, delta=spec.delta, handle_nan=True)
    rc_o = orc.solve(s, x, 1e300)
    sol = SlabSolver(spec, lib_path=gpu_lib)
    sol.upload(w.copy())
    sol.sys.h, sol.sys.h_min, sol.sys.handle_nan = 1e200, 0.0, 1
    sol.sys.max_attempts = 0
    s2 = orc.new_system(t=0.0, h=1e200, h_min=0.0, delta=spec.delta, handle_nan=True, max_attempts=12)
    x2 = w.copy()
    rc_o = orc.solve(s2, x2, 1e300)
    r = sol.solve(1e300, max_attempts=12)
    sol.close()
    assert r.code == rc_o
    assert r.last_nan == s2.last_nan == 1
    assert (r.t, r.h, r.steps, r.steps_total) == (s2.t, s2.h, s2.steps, s2.steps_total)


# ---- full-size properties (BASELINE sizes; the oracle would take minutes here) ------------

@pytest.fixture(scope="module")
def hr_state(gpu_lib):
    """the "400" high-resolution case (200x200x400, calc_mode 1), 12 attempts in"""
    spec = _real_spheres(cases.archived_case("400-SigmaP1-P-smallsigma"))
    sol = SlabSolver(spec, lib_path=gpu_lib)
    u, p, gl = spec.initial_fields()
    sol.upload_interior(u, p, gl)
    sol.rasterise(spec.spheres)
    u, p, gl = sol.download_interior()
    r = sol.solve(spec.final_time, max_attempts=12)
    out = sol.download_interior()
    sol.close()
    return spec, (u, p, gl), r, out


@pytest.mark.parametrize("env,flags", [
    ({"PFT_TX": "40", "PFT_TY": "6", "PFT_ZCHUNKS": "7"}, 0),
    ({"PFT_TX": "50", "PFT_TY": "3", "PFT_ZCHUNKS": "1"}, 0),
    ({}, FLAG_LEGACY_LOADS),
    ({}, FLAG_RECOMPUTE_K1 | FLAG_NO_RUNAHEAD),
    ({}, FLAG_HOST_POW),
])
def test_fullsize_invariance(gpu_lib, hr_state, monkeypatch, env, flags):
    """results do not depend on tiling, z-chunking, load path, K1 reuse, run-ahead or where pow() is computed
    (bit-exact).  The comparison of this workload with the REFERENCE is tests/test_zz_interval_parity.py
    (partial_400-*.npz)."""
    spec, (u, p, gl), r0, out0 = hr_state
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sol = SlabSolver(spec, lib_path=gpu_lib, flags=flags)
    sol.upload_interior(u, p, gl)
    r = sol.solve(spec.final_time, max_attempts=12)
    out = sol.download_interior()
    sol.close()
    assert (r.t, r.h, r.steps, r.steps_total) == (r0.t, r0.h, r0.steps, r0.steps_total)
    for a, b in zip(out, out0):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("grid", [(200, 200, 50), (100, 100, 200), (50, 50, 100)])
def test_stage_data_flow_matches_grid_wait(gpu_lib, monkeypatch, grid):
    """The three ways an attempt's launches are ordered on the device, on short grids where the overlap is largest:
    (default) [S2 S3 S4 S5 S1'] with K1 of the next attempt evaluated speculatively from stage 5's candidate and
    stage s+1 starting on the CTAs of stage s around it (dep_wait); the classic [S1 .. S5] with that data flow
    (PFT_SPEC=0); and the same kernels waiting for the whole predecessor grid (PFT_DATAFLOW=0).  1500 attempts in
    three solve calls: state, step counts, rhs_evals and fields bit-identical."""
    n1, n2, n3 = grid
    spec = cases.cube_case(n1, nz=n3, calc_mode=1, spheres=200)
    spec.n2 = n2
    u, p, gl = spec.initial_fields()
    out = []
    for flow, spc in (("1", "1"), ("1", "0"), ("0", "0")):
        monkeypatch.setenv("PFT_DATAFLOW", flow)
        monkeypatch.setenv("PFT_SPEC", spc)
        sol = SlabSolver(spec, lib_path=gpu_lib)
        assert sol.stats()["speculative_stage1"] == int(spc)
        sol.upload_interior(u, p, gl)
        sol.rasterise(spec.spheres)
        trace = []
        for budget in (7, 993, 500):
            r = sol.solve(spec.final_time, max_attempts=budget)
            trace.append((r.t, r.h, r.steps, r.steps_total, r.rhs_evals))
        out.append((trace, sol.download_interior()))
        sol.close()
    for trace, fields in out[1:]:
        assert trace == out[0][0]
        for a, b in zip(fields, out[0][1]):
            assert np.array_equal(a, b)


def test_fullsize_sanity(hr_state):
    spec, (u, p, gl), r, (u1, p1, gl1) = hr_state
    assert np.array_equal(gl, gl1)                      # the glass field is static (equation.c:731)
    assert np.isfinite(u1).all() and np.isfinite(p1).all()
    assert r.steps >= 1 and r.steps_total == 12
    # latent heat (L*dp/dt) may lift u slightly above the initial 293.15 K near the growing ice cap
    assert 248.15 - 1e-9 <= u1.min() and u1.max() < 300.0
    assert -1e-6 <= p1.min() and p1.max() <= 1 + 1e-6


RASTER_ULP = 2      # include/pft_b200.h, SURVEY.md 8(f)-1: device tanh within 2 ulp (of 1.0) of the host libm result


def test_device_rasteriser_vs_host(gpu_lib):
    """pft_rasterise (device tanh, spheres binned) against the oracle's host rasteriser (libm tanh, every
    sphere tested): |difference| <= RASTER_ULP * 2^-52 -- the one bar stated in include/pft_b200.h"""
    spec = _real_spheres(cases.archived_case("100-GradP"))
    case, orc, w = parity.make(spec)
    sol = SlabSolver(spec, lib_path=gpu_lib)
    u, p, gl = spec.initial_fields()
    sol.upload_interior(u, p, gl)
    sol.rasterise(spec.spheres)
    _, _, gl_dev = sol.download_interior()
    sol.close()
    assert np.abs(gl_dev - case.interior(w)[2]).max() <= RASTER_ULP * np.finfo(float).eps


def test_device_rasteriser_binning_is_exact(gpu_lib, monkeypatch):
    """the binned rasteriser (a node visits the 27 bins around it) against the same kernel with ONE bin
    (every sphere tested): bit-identical, on a pack with many small spheres incl. centres outside the block"""
    spec = cases.cube_case(96, nz=64, spheres=1500)
    rng = np.random.default_rng(3)
    spec.spheres = rng.uniform(-0.05, 1.05, size=(1500, 3))          # some centres outside the block
    out = []
    for brute in ("0", "1"):
        monkeypatch.setenv("PFT_RASTER_ONE_BIN", brute)
        sol = SlabSolver(spec, lib_path=gpu_lib)
        u, p, gl = spec.initial_fields()
        sol.upload_interior(u, p, gl)
        sol.rasterise(spec.spheres)
        out.append(sol.download_interior()[2])
        sol.close()
    assert np.array_equal(out[0], out[1])
    assert out[0].max() > 0.99 and np.count_nonzero(out[0] > 0.5) > 1000


def test_ice_fraction(gpu_lib):
    spec = cases.small_case(24, 20, 32, calc_mode=0)
    case, orc, w = parity.make(spec)
    sol = SlabSolver(spec, lib_path=gpu_lib)
    sol.upload(w.copy())
    f = sol.ice_fraction()
    sol.close()
    assert f == orc.ice_fraction(w)


def test_fast_variant_within_tolerance():
    """the opt-in FMA build (libpft_b200_fast.so): fixed-step fields within rel. L_inf 1e-10 of the
    oracle (north-star tolerance); it makes no claim on the adaptive step sequence"""
    from porousfreezethaw_b200 import _lib
    if not os.path.exists(_lib.FAST_LIB_PATH):
        pytest.skip("fast variant not built")
    for mode in (0, 1):
        spec = cases.small_case(24, 20, 32, calc_mode=mode)
        case, orc, w = parity.make(spec)
        x = w.copy()
        s = orc.new_system(t=0.0, h=0.25, h_min=float("inf"), delta=spec.delta)
        sol = SlabSolver(spec, lib_path=_lib.FAST_LIB_PATH)
        sol.upload(w.copy()); sol.sys.h_min = float("inf")
        for _ in range(20):
            s.h = 0.25; assert orc.solve(s, x, s.t + 0.25) == 0
            sol.sys.h = 0.25; assert sol.solve(sol.sys.t + 0.25).code == 0
        got = w.copy(); sol.download(got); sol.close()
        a, b = case.interior(got), case.interior(x)
        for q in (0, 1):
            assert parity.rel_linf(a[q], b[q]) <= 1e-10
