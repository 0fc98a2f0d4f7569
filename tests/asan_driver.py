"""Run by tests/test_emulated_asan.py in a subprocess with libasan preloaded: the launch orders of the device loop
(speculative S2..S5 S1' and classic S1..S5), budgets that end a solve call after one attempt, re-entry, ragged tiles,
several z-chunks, a three-slab group on the peer path, the binned rasteriser and the standalone right-hand side."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np                                             # noqa: E402
import parity                                                  # noqa: E402
from porousfreezethaw_b200 import cases                        # noqa: E402
from porousfreezethaw_b200.solver import GroupSolver, SlabSolver   # noqa: E402

lib = sys.argv[1]
os.environ.update(PFT_TX="8", PFT_TY="4", PFT_ZCHUNKS="3")
out = []
for mode in (1, 2, 11):
    spec = cases.small_case(13, 7, 11, calc_mode=mode)
    u, p, gl = spec.initial_fields()
    for on in ("1", "0"):
        os.environ["PFT_SPEC"] = on
        sol = SlabSolver(spec, lib_path=lib)
        sol.upload_interior(u, p, gl)
        sol.rasterise(spec.spheres)
        for b in (1, 2, 5, 0):
            r = sol.solve(6.0, max_attempts=b)
        out.append((mode, r.t, r.steps, r.steps_total))
        sol.debug_fetch(1)
        x = np.zeros(3 * sol.SIZE)
        sol.download(x)
        sol.rhs_host(r.t, x)
        sol.close()
assert out[0] == out[1] and out[2] == out[3] and out[4] == out[5], out
os.environ["PFT_SPEC"] = "1"
for k in ("PFT_TX", "PFT_TY", "PFT_ZCHUNKS"):
    del os.environ[k]
spec = cases.small_case(12, 10, 19, calc_mode=1)
case, orc, w = parity.make(spec)
g = GroupSolver(spec, 3, lib_path=lib)
g.upload(w.copy())
r = g.solve(10.0)
got = w.copy()
g.download(got)
g.close()
print("ASAN-DRIVER-OK", out[0], (r.t, r.steps, r.steps_total))
