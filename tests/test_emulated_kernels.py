"""CPU checks of the CUDA sources through tests/cuda_emu (fibers instead of a GPU): the same
parity functions the GPU tests run, on small grids.  These validate kernel indexing, the fused
stage algebra and the integrator's control flow; the GPU tests validate the real thing."""
import pytest

import parity
from porousfreezethaw_b200 import cases


@pytest.mark.parametrize("mode", [0, 1, 2, 10, 11])
def test_rhs(emu_lib, mode):
    parity.check_rhs(cases.small_case(12, 10, 16, calc_mode=mode), emu_lib)


@pytest.mark.parametrize("mode", [0, 1])
def test_rhs_tiny_and_subnormal_values(emu_lib, mode):
    parity.check_rhs(cases.small_case(12, 10, 16, calc_mode=mode), emu_lib, perturb="tiny")


@pytest.mark.parametrize("mode,flags", [(0, 0), (0, 16), (0, 8), (0, 24), (1, 0), (1, 16), (2, 0), (2, 16), (10, 0)])
def test_fused_stages(emu_lib, mode, flags):
    parity.check_fused_stages(cases.small_case(12, 10, 8, calc_mode=mode), emu_lib, flags=flags)


def test_ragged_tiles_and_chunks(emu_lib, monkeypatch):
    """grid not a multiple of the tile, several tiles in x and y, several z-chunks"""
    monkeypatch.setenv("PFT_TX", "5")
    monkeypatch.setenv("PFT_TY", "4")
    monkeypatch.setenv("PFT_ZCHUNKS", "3")
    spec = cases.small_case(13, 7, 11, calc_mode=0)
    parity.check_rhs(spec, emu_lib)
    parity.check_fused_stages(spec, emu_lib)


@pytest.mark.parametrize("mode", [0, 11])
def test_adaptive_sequence(emu_lib, mode):
    parity.check_adaptive(cases.small_case(12, 10, 8, calc_mode=mode), emu_lib, T=12.0, chunks=2, with_service=True)


def test_fixed_steps(emu_lib):
    parity.check_fixed_steps(cases.small_case(12, 10, 8, calc_mode=1), emu_lib, nsteps=3)


@pytest.mark.parametrize("rows,nobar", [("2", "1"), ("2", "0"), ("21212", "1"), ("1", "0")])
def test_second_generation_kernel_variants(emu_lib, monkeypatch, rows, nobar):
    """merson_stage_v2: R cells per thread (y-faces shared inside a thread, rows beyond the domain are mirror
    duplicates), stages 1 and 3 without CTA barrier (producer warp + empty mbarriers): ragged tiles in x and y,
    several z-chunks (remainder planes of the 3-way unrolled march), bit-exact RHS / fused stages / adaptive run"""
    monkeypatch.setenv("PFT_ROWS", rows)
    monkeypatch.setenv("PFT_NOBAR", nobar)
    monkeypatch.setenv("PFT_TX", "8")
    monkeypatch.setenv("PFT_TY", "4")
    monkeypatch.setenv("PFT_ZCHUNKS", "2")
    for mode in (0, 1):
        spec = cases.small_case(13, 7, 11, calc_mode=mode)
        parity.check_rhs(spec, emu_lib)
        parity.check_fused_stages(spec, emu_lib)
    parity.check_adaptive(cases.small_case(13, 7, 11, calc_mode=1), emu_lib, T=6.0, chunks=2)
    from porousfreezethaw_b200.solver import SlabSolver
    sol = SlabSolver(cases.small_case(13, 7, 11, calc_mode=1), lib_path=emu_lib)
    st = sol.stats()
    sol.close()
    assert st["kernel_generation"] == 2 and st["rows_per_thread"] == int(rows[-1])


def test_speculative_stage1_matches_classic_order(emu_lib, monkeypatch):
    """Attempts launched as [S2 S3 S4 S5 S1'] with two alternating control blocks and one K1 buffer per state pair
    (the default) against the classic [S1 .. S5] order (PFT_SPEC=0): several solve calls, attempt budgets that end a
    call after a rejected and after an accepted attempt, an upload in between (K1 invalid: stage 1 runs at solve
    entry) -- step sequence, counters (rhs_evals included) and fields bit-identical."""
    import numpy as np
    from porousfreezethaw_b200.solver import SlabSolver
    monkeypatch.setenv("PFT_TX", "8")
    monkeypatch.setenv("PFT_TY", "4")
    monkeypatch.setenv("PFT_ZCHUNKS", "3")
    spec = cases.small_case(13, 7, 11, calc_mode=1)
    u, p, gl = spec.initial_fields()
    runs = []
    for on in ("1", "0"):
        monkeypatch.setenv("PFT_SPEC", on)
        sol = SlabSolver(spec, lib_path=emu_lib)
        sol.upload_interior(u, p, gl)
        sol.rasterise(spec.spheres)
        trace = []
        for budget in (1, 2, 1, 3, 5, 4):
            r = sol.solve(6.0, max_attempts=budget)
            trace.append((r.code, r.t, r.h, r.steps, r.steps_total, r.rhs_evals))
        k1 = sol.debug_fetch(1)
        if on == "1":
            # K1 of the state at rest is what stage 1 computes from it (S1' ran on stage 5's candidate)
            x = np.zeros(3 * sol.SIZE)
            sol.download(x)
            probe = SlabSolver(spec, lib_path=emu_lib)
            probe.upload_interior(u, p, gl)
            probe.rasterise(spec.spheres)
            xx = np.zeros(3 * sol.SIZE)
            probe.download(xx)
            xx[:2 * sol.SIZE] = x[:2 * sol.SIZE]
            dw = probe.rhs_host(r.t, xx).reshape(3, sol.N3, sol.N2, sol.N1)[:, 2:-2, 2:-2, 2:-2]
            probe.close()
            assert np.array_equal(dw[0], k1[0]) and np.array_equal(dw[1], k1[1])
        uu, pp, gg = sol.download_interior()
        sol.upload_interior(uu, pp, None)             # same values, but K1 is no longer known to be valid
        r = sol.solve(6.0)
        trace.append((r.code, r.t, r.h, r.steps, r.steps_total, r.rhs_evals))
        st = sol.stats()
        assert st["speculative_stage1"] == int(on)
        runs.append((trace, k1, sol.download_interior()))
        sol.close()
    (t1, _, f1), (t0, _, f0) = runs
    assert t1 == t0
    assert t1[-1][0] == 0 and t1[-1][1] == 6.0
    for a, b in zip(f1, f0):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_split_barrier_variant_of_forming_stages(emu_lib, monkeypatch, mode):
    """PFT_SPLITBAR=1: stages 2, 4, 5 with an arrive / wait pair instead of the per-plane CTA barrier (warps drift up to
    one plane apart, the last warp to arrive requests the next raw plane).  Exists for the default 50x5 tile only, so
    the grid here is two such tiles, three z-chunks with remainder planes: fused stages and an adaptive run against
    the oracle, bit for bit (tolerance of the oracle comparison for mode 2 as everywhere)."""
    from porousfreezethaw_b200.solver import SlabSolver
    monkeypatch.setenv("PFT_SPLITBAR", "1")
    monkeypatch.setenv("PFT_ZCHUNKS", "3")
    spec = cases.small_case(50, 10, 11, calc_mode=mode)
    sol = SlabSolver(spec, lib_path=emu_lib)
    st = sol.stats()
    sol.close()
    assert (st["tile_x"], st["tile_y"], st["kernel_generation"]) == (50, 5, 2)
    parity.check_fused_stages(spec, emu_lib)
    parity.check_adaptive(spec, emu_lib, T=4.0, chunks=2)
