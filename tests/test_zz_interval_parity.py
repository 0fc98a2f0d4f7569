"""SURVEY.md 8(d), adaptive-mode parity over a whole first snapshot interval of the archived cases
(configs 1-2): the accepted-step sequence (t_i, h_i), steps / steps_total, the hand-over step size
and the final fields of the B200 path against what the UNMODIFIED reference produced for the same
Params and sphere pack (tests/golden/interval_*.npz, made by tests/golden/make_golden_interval.py:
the reference needs minutes of CPU per case, the GPU a second).  Bar: bit-exact; the fields are
compared through SHA-256 digests of their float64 bytes plus a strided sample for diagnostics.

(The file sorts last on purpose: these are the longest GPU tests.)"""
import hashlib
import os

import numpy as np
import pytest

from porousfreezethaw_b200 import cases
from porousfreezethaw_b200.solver import SlabSolver

HERE = os.path.dirname(os.path.abspath(__file__))
SPHERES = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "data", "spheres_positions.txt")


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<f8").tobytes()).hexdigest()


def _goldens():
    d = os.path.join(HERE, "golden")
    return sorted(f[len("interval_"):-len(".npz")] for f in os.listdir(d) if f.startswith("interval_") and f.endswith(".npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", _goldens())
def test_first_snapshot_interval_matches_reference(gpu_lib, name):
    if not os.path.exists(SPHERES):
        pytest.skip("oracle/_ref/data/spheres_positions.txt not staged")
    g = np.load(os.path.join(HERE, "golden", f"interval_{name}.npz"))
    spec = cases.archived_case(name)
    spec.spheres = np.loadtxt(SPHERES).reshape(-1, 3)
    T = float(g["T"])
    assert T == spec.final_time / (spec.saved_files - 1)
    sol = SlabSolver(spec, lib_path=gpu_lib)
    try:
        u, p, gl = spec.initial_fields()
        sol.upload_interior(u, p, gl)
        # the reference rasterises with libm tanh on the host; the product's drop-in plug-in does the same
        # (pft_equation_rasterise_host), the device rasteriser is within 2 ulp: take the host path here
        from oracle import pyoracle as po
        case = po.Case(spec.n1, spec.n2, spec.n3, spec.L1, spec.L2, spec.L3, spec.param_array(), spec.calc_mode)
        w = case.pack(u, p, gl)
        po.Oracle(case).rasterise(spec.spheres, w)
        sol.upload(w)
        rec = []
        sol.set_service(lambda ft, t, h, st, tot: rec.append((t, h)) or 0)
        r = sol.solve(T)
        assert r.code == 0
        want = g["state"]
        assert (r.t, r.h, r.steps, r.steps_total) == (want[0], want[1], int(want[2]), int(want[3])), \
            f"gpu {(r.t, r.h, r.steps, r.steps_total)} reference {tuple(want)}"
        got = np.array(rec)
        ref = g["record"].reshape(-1, 2)
        assert got.shape == ref.shape
        bad = np.nonzero((got != ref).any(axis=1))[0]
        assert bad.size == 0, f"first differing accepted step #{bad[0]}: gpu {got[bad[0]]} reference {ref[bad[0]]}"
        uo, po_, glo = sol.download_interior()
        assert np.array_equal(uo[::7, ::5, ::3], g["sample_u"]), f"u sample differs, max {np.abs(uo[::7, ::5, ::3] - g['sample_u']).max():.3e}"
        assert np.array_equal(po_[::7, ::5, ::3], g["sample_p"]), f"p sample differs, max {np.abs(po_[::7, ::5, ::3] - g['sample_p']).max():.3e}"
        assert _digest(glo) == str(g["sha_gl"])
        assert _digest(uo) == str(g["sha_u"]) and _digest(po_) == str(g["sha_p"])
    finally:
        sol.close()
