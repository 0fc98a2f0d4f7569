#!/usr/bin/env python
"""Regenerates tests/golden/interval_<case>.npz from the UNMODIFIED reference (oracle/_ref/libpft_ref.so):
an archived production case with the reference's own sphere pack, integrated adaptively over its
whole FIRST SNAPSHOT INTERVAL (t: 0 -> final_time/99, SURVEY.md 8(d) "adaptive mode parity").  Stored:
the accepted-step record (t_i, h_i) from a Service_Callback, the solver state at the end, SHA-256
digests of the final u and p interiors (the fields themselves would be megabytes) and a strided
sample of both for diagnostics.  Run in the build container only (minutes per case):

    python tests/golden/make_golden_interval.py 100-GradP 100-SigmaP1-P [200-GradP ...]
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


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<f8").tobytes()).hexdigest()


def main():
    for name in sys.argv[1:]:
        spec = cases.archived_case(name)
        T = spec.final_time / (spec.saved_files - 1)            # intertrack.c:2271
        case = po.Case(spec.n1, spec.n2, spec.n3, spec.L1, spec.L2, spec.L3, spec.param_array(), spec.calc_mode)
        u, p, gl = spec.initial_fields()
        ref = po.Reference()
        ref.setup(case, u, p, gl, SPHERES, t0=0.0, tau=spec.tau, tau_min=spec.tau_min, delta=spec.delta, use_callback=True)
        t0 = time.time()
        rc = ref.solve(T)
        assert rc == 0
        t, h, steps, total = ref.state()
        x = case.interior(ref.solution().copy())
        rec = ref.record()
        ref.close()
        np.savez_compressed(os.path.join(HERE, f"interval_{name}.npz"), name=name, T=T, state=np.array([t, h, steps, total]),
                            record=rec, sha_u=digest(x[0]), sha_p=digest(x[1]), sha_gl=digest(x[2]),
                            sample_u=x[0][::7, ::5, ::3].copy(), sample_p=x[1][::7, ::5, ::3].copy())
        print(name, "T", T, "steps", steps, "total", total, "h", h, f"{time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
