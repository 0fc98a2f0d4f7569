#!/usr/bin/env python
"""Regenerates tests/golden/partial_<case>.npz from the UNMODIFIED reference (oracle/_ref/libpft_ref.so):
the head of an archived production case (reference sphere pack, adaptive mode) up to the first
ACCEPTED step at which steps_total >= N -- the reference's own Service_Callback break
(RK_MPI_SAsolver_hybrid2.c:676-705: the solver saves t, h = new_h and returns 1).  Used for the
cases whose whole first snapshot interval costs hours of CPU: the "400" high-resolution cases
(the benchmarked workload), the mode-2 "Temp" cases and the PhysRevE-2025 "200" cases.

Stored: the accepted-step record (t_i, h_i), the state at the break, SHA-256 digests of the final
u, p, gl interiors, a strided sample of u and p (diagnostics, and the tolerance bar of mode 2),
and -- for bench.py's multi-GPU `parity` key -- the same after K2 < N attempts where asked.

    python tests/golden/make_golden_partial.py 400-SigmaP1-P-smallsigma:100 200-Temp:400 ...
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from porousfreezethaw_b200 import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
SPHERES = os.path.join(ROOT, "oracle", "_ref", "data", "spheres_positions.txt")
STRIDE = (7, 5, 3)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<f8").tobytes()).hexdigest()


def plane_sums(a: np.ndarray) -> np.ndarray:
    """per z-plane sum (mod 2^64) of the float64 bit patterns: a checksum every slab owner can compute for
    its own planes (bench.py's multi-GPU `parity` key)"""
    return np.ascontiguousarray(a, dtype="<f8").view(np.uint64).reshape(a.shape[0], -1).sum(axis=1, dtype=np.uint64)


def main():
    for arg in sys.argv[1:]:
        name, n = arg.rsplit(":", 1)
        n = int(n)
        spec = cases.archived_case(name)
        T = spec.final_time / (spec.saved_files - 1)            # intertrack.c:2271
        case = po.Case(spec.n1, spec.n2, spec.n3, spec.L1, spec.L2, spec.L3, spec.param_array(), spec.calc_mode)
        u, p, gl = spec.initial_fields()
        ref = po.Reference()
        ref.setup(case, u, p, gl, SPHERES, t0=0.0, tau=spec.tau, tau_min=spec.tau_min, delta=spec.delta, use_callback=True)
        ref.lib.pftref_break_after(n)
        t0 = time.time()
        rc = ref.solve(T)
        assert rc == 1, rc                                       # interrupted by the callback
        t, h, steps, total = ref.state()
        x = case.interior(ref.solution().copy())
        rec = ref.record()
        ref.close()
        sz, sy, sx = STRIDE
        np.savez_compressed(os.path.join(HERE, f"partial_{name}.npz"), name=name, T=T, break_after=n,
                            state=np.array([t, h, steps, total]), record=rec,
                            sha_u=digest(x[0]), sha_p=digest(x[1]), sha_gl=digest(x[2]),
                            sample_u=x[0][::sz, ::sy, ::sx].copy(), sample_p=x[1][::sz, ::sy, ::sx].copy(),
                            plane_sum_u=plane_sums(x[0]), plane_sum_p=plane_sums(x[1]), plane_sum_gl=plane_sums(x[2]),
                            absmax=np.array([np.abs(x[0]).max(), np.abs(x[1]).max()]))
        print(name, "break_after", n, "t", t, "steps", steps, "total", total, "h", h, f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
