"""Host-side logic that needs neither GPU nor reference: slab partition, case tables, the
reference-shaped solver-module entry points of the C drop-in (argument checking, return codes)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from porousfreezethaw_b200 import _lib, cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("total,nranks", [(400, 1), (400, 8), (100, 3), (101, 4), (7, 7), (50, 6)])
def test_partition_matches_reference_rule(total, nranks):
    """intertrack.c:1780-1787: contiguous slabs, remainder rows one each to the lowest ranks"""
    lib = po.load_oracle()
    row = 0
    for r in range(nranks):
        n3, first = cases.partition(total, nranks, r)
        a, b = C.c_int(), C.c_int()
        lib.pfo_partition(total, nranks, r, C.byref(a), C.byref(b))
        assert (n3, first) == (a.value, b.value) == po.partition(total, nranks, r)
        assert first == row and n3 in (total // nranks, total // nranks + 1)
        row += n3
    assert row == total


def test_archived_case_tables():
    for name, grid in (("100-GradP", (50, 50, 100)), ("200-GradP", (100, 100, 200)), ("400-SigmaP1-P-smallsigma", (200, 200, 400))):
        cs = cases.archived_case(name)
        assert (cs.n1, cs.n2, cs.n3) == grid and cs.cells == grid[0] * grid[1] * grid[2]
        p = cs.params
        assert p["alpha"] == 997.0 * 4.18e3 * (3 if "smallsigma" in name else 1)
        assert p["xi"] == 0.06 / (grid[2] / 1) * 1.0 if False else p["xi"] == 0.06 / grid[2]
        assert p["top_temp1"] == 273.15 - 25 and p["phase_switch_time"] == 5 * 3600.0
        assert cs.snapshot_time(1) == pytest.approx(36000.0 / 99)
    assert len(cases.PARAM_NAMES) == _lib.NPARAM == po.NPARAM and cases.PARAM_NAMES == po.PARAM_NAMES


def test_solver_module_return_codes():
    """RK_MPI_SA_init / cleanup / check_mem follow RK_MPI_SAsolver_hybrid2.c:72-212"""
    subprocess.run(["make", "-C", os.path.join(ROOT, "porousfreezethaw_b200", "csrc"), "all"], check=True, stdout=subprocess.DEVNULL)
    rk = C.CDLL(_lib.RK_LIB_PATH)

    class MemDist(C.Structure):
        _fields_ = [("n_chunks", C.c_int), ("chunk_start", C.POINTER(C.c_int)), ("chunk_size", C.POINTER(C.c_int)),
                    ("chunk_eps_mult", C.POINTER(C.c_double))]

    def dist(starts, sizes):
        n = len(starts)
        a, b, m = (C.c_int * n)(*starts), (C.c_int * n)(*sizes), (C.c_double * n)(*([1.0] * n))
        d = MemDist(n, a, b, m)
        d._keep = (a, b, m)
        return d

    assert rk.RK_MPI_SA_cleanup() == -3
    assert rk.RK_MPI_SA_check_mem(C.byref(dist([0], [4]))) == -3
    assert rk.RK_MPI_SA_init(0, 0, 0) == -2
    assert rk.RK_MPI_SA_init(100, 0, 0) == 0
    assert rk.RK_MPI_SA_init(100, 0, 0) == -3
    assert rk.RK_MPI_SA_check_mem(C.byref(dist([0, 10, 20], [5, 10, 30]))) == 0
    assert rk.RK_MPI_SA_check_mem(C.byref(dist([0, 10, 20], [5, 10, 81]))) == -5
    assert rk.RK_MPI_SA_check_mem(C.byref(dist([0, 4], [5, 10]))) == -6
    assert rk.RK_MPI_SA_check_mem(C.byref(dist([0, 10], [5, 0]))) == -6
    d0 = dist([0], [1]); d0.n_chunks = 0
    assert rk.RK_MPI_SA_check_mem(C.byref(d0)) == -7
    assert rk.RK_MPI_SA_solve(C.c_double(1.0), None) == -2
    rk.RK_MPI_SA_handle_NAN(5)
    assert rk.RK_MPI_SA_check_NAN() == 0
    assert rk.RK_MPI_SA_cleanup() == 0



def test_reference_arm_prints_one_contract_line():
    """bench.py --impl reference: exactly one JSON line on stdout (the reference's own chatter goes to
    stderr), carrying the keys the measurement contract names"""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                        "--workload", "100-GradP"], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mcell-RHS/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["value"] > 0
    assert d["config"]["workload"] == "100-GradP" and d["config"]["grid"] == [50, 50, 100]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.parametrize("grid", [(100, 100, 200), (200, 200, 50), (200, 200, 400), (50, 50, 100), (24, 20, 6), (400, 400, 50)])
def test_z_chunking_rule(emu_lib, grid):
    """choose_chunks (csrc/pft_b200.cu): about four waves of CTAs at three resident CTAs per SM, never
    fewer than seven planes per chunk (profiles/round1_tuning.md); a block thinner than that is one
    chunk.  (The emulated device reports 4 SMs; a B200 has 148.)"""
    from porousfreezethaw_b200.solver import SlabSolver
    n1, n2, n3 = grid
    spec = cases.small_case(n1, n2, n3, calc_mode=1)
    sol = SlabSolver(spec, lib_path=emu_lib)
    try:
        st = sol.stats()
    finally:
        sol.close()
    sms = 4
    tiles = -(-n1 // st["tile_x"]) * -(-n2 // st["tile_y"])
    nzc = max(1, min((4 * 3 * sms + tiles // 2) // tiles, max(1, n3 // 7)))
    zc = -(-n3 // nzc)
    assert st["z_chunk"] == zc, (st["z_chunk"], zc, tiles)
    assert st["z_chunk"] >= min(7, n3)
    assert st["grid_blocks"] == tiles * -(-n3 // st["z_chunk"])


@pytest.mark.parametrize("extra", [[], ["--fixed-step"]])
def test_bench_line_schema_dry_run(emu_lib, extra):
    """bench.py walked through end to end against the CPU emulation (PFT_BENCH_DRYRUN_LIB): one JSON line
    with every key of the measurement contract.  The values mean nothing (the line says so)."""
    import json
    import sys
    env = dict(os.environ, PFT_BENCH_DRYRUN_LIB=emu_lib)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--workload", "slab-24x16",
                        "--profile-steps", "2", "--cpu-steps", "2", *extra], capture_output=True, text=True, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert "DRY RUN" in d["data"]
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "Mcell-RHS/s" and d["dtype"] == "f64" and d["n_gpus"] == 1 and d["steps"] == 3 and d["vs_baseline"] is None
    assert d["config"]["workload"] and "model" not in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert len(d["roofline"]["per_stage"]) == 5 and [s["bytes_per_cell"] for s in d["roofline"]["per_stage"]] == [40, 56, 40, 72, 88]
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(d["cpu_baseline"])
    # 5 stage launches per attempt (stage 1 returns at once while K1 is still valid); the adaptive run may enqueue
    # attempts beyond the budgeted end only up to the budget itself
    assert d["gpu_launches"] == 5 * 3
    assert len(d["device_ms_regions"]) == 7 and "parity" in d and "also" in d
    assert ("fixed step" in d["config"]["mode"]) == bool(extra)
