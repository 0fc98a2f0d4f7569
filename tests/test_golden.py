"""The oracle (oracle/pft_oracle.c) and, on a GPU, the CUDA path against the committed golden
vectors in tests/golden/ -- generated from the UNMODIFIED reference by tests/golden/make_golden.py."""
import glob
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from porousfreezethaw_b200 import cases
from porousfreezethaw_b200.solver import SlabSolver

# (interval_*.npz / partial_*.npz are the step records of tests/test_zz_interval_parity.py, bench_*.npz of bench.py)
GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz"))
              if not os.path.basename(p).startswith(("interval_", "partial_", "bench_")))


def load(path):
    g = np.load(path)
    n = [int(v) for v in g["n"]]
    L = g["L"]
    case = po.Case(n[0], n[1], n[2], float(L[0]), float(L[1]), float(L[2]), g["param"], int(g["mode"]))
    x0 = g["x0"]
    w = case.pack(x0[0], x0[1], x0[2])
    return g, case, w


def spec_of(g, case):
    spec = cases.small_case(case.n1, case.n2, case.total_n3, calc_mode=case.calc_mode)
    assert np.array_equal(spec.param_array(), g["param"]), "small_case() drifted from the fixture's parameters"
    return spec


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_oracle_matches_reference_vectors(path):
    g, case, w = load(path)
    orc = po.Oracle(case)
    k1 = case.interior(orc.rhs(0.0, w.copy()))
    assert np.array_equal(k1[:2], g["k1"]) and not k1[2].any()
    rec = []
    s = orc.new_system(t=0.0, h=float(g["tau"]), h_min=float(g["tau_min"]), delta=float(g["delta"]),
                       service=lambda ft, t, h, st, tot, ctx: rec.append((t, h)) or 0)
    x = w.copy()
    assert orc.solve(s, x, float(g["T"])) == 0
    assert [s.t, s.h, s.steps, s.steps_total] == list(g["state"])
    assert np.array_equal(np.array(rec), g["record"])
    assert np.array_equal(case.interior(x)[:2], g["xT"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_matches_reference_vectors(path, gpu_lib):
    g, case, w = load(path)
    exact = case.calc_mode != 2
    sol = SlabSolver(spec_of(g, case), lib_path=gpu_lib)
    try:
        k1 = case.interior(sol.rhs_host(0.0, w.copy()))
        rec = []
        sol.upload(w.copy())
        sol.set_service(lambda ft, t, h, st, tot: rec.append((t, h)) or 0)
        r = sol.solve(float(g["T"]))
        got = w.copy()
        sol.download(got)
    finally:
        sol.close()
    xT = case.interior(got)[:2]
    if exact:
        assert np.array_equal(k1[:2], g["k1"])
        assert [r.t, r.h, r.steps, r.steps_total] == list(g["state"])
        assert np.array_equal(np.array(rec), g["record"])
        assert np.array_equal(xT, g["xT"])
    else:   # calc_mode 2: device cosh differs from glibc in the last ulp
        assert np.abs(k1[:2] - g["k1"]).max() <= 1e-10 * np.abs(g["k1"]).max()
        assert r.t == g["state"][0] and abs(r.steps_total - g["state"][3]) <= 1
        assert np.abs(xT - g["xT"]).max() <= 1e-10 * np.abs(g["xT"]).max()
