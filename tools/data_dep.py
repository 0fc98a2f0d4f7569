#!/usr/bin/env python
"""tools/data_dep.py -- does the stage kernels' speed depend on the field values? (development aid)
Times one thin 200x200x50 slab filled with the lower / the upper half of the 200x200x100 case's state."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from porousfreezethaw_b200.solver import SlabSolver
from bench import make_spec

big, _, _ = make_spec("slab-200x100", 1)
spec, _, _ = make_spec("slab-200x50", 1)
sb = SlabSolver(big)
sb.upload_interior(*big.initial_fields())
sb.rasterise(big.spheres)
sb.solve(big.final_time, max_attempts=30)
U, P, G = sb.download_interior()
print("big case after 30 attempts: t =", sb.sys.t, "h =", sb.sys.h)
sb.close()
for name, sl in (("lower half", slice(0, 50)), ("upper half", slice(50, 100))):
    u, p, g = (np.ascontiguousarray(a[sl]) for a in (U, P, G))
    tiny = np.count_nonzero((np.abs(p) < 1e-290) & (p != 0))
    print(f"{name}: p in [{p.min():.3e}, {p.max():.3e}], {np.count_nonzero(p == 0)} zeros, {tiny} tiny non-zero; u in [{u.min():.6f}, {u.max():.6f}]")
    sol = SlabSolver(spec)
    sol.upload_interior(u, p, g)
    sol.solve(spec.final_time, max_attempts=10)
    sol.reset_stats()
    sol.solve(spec.final_time, max_attempts=100)
    st = sol.stats()
    print(f"   {st['device_ms_total'] / 100 * 1e3:7.1f} us/attempt  accepted {st['accepted']}/100", flush=True)
    for which, nm in ((1, "K1"), (3, "K3"), (4, "K4")):
        ku, kp = sol.debug_fetch(which)
        for a, an in ((ku, "u"), (kp, "p")):
            nz = a[a != 0]
            print(f"   {nm}{an}: zeros {np.count_nonzero(a == 0)}, |x|<1e-290 non-zero {np.count_nonzero(np.abs(nz) < 1e-290)}, min|nz| {np.abs(nz).min() if nz.size else 0:.3e}")
    sol.close()
