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
