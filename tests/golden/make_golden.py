#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libpft_ref.so, built
from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds, for one small case: the case parameters, the initial padded array (IC +
sphere pack as the reference's PrecalculateData rasterised it), K1 = f(0, x0) from the reference's
right-hand side, and the state after an adaptive RK_MPI_SA_solve to T with the accepted-step
record (t_i, h_i) captured by a Service_Callback.  Only interiors are stored (float64, exact)."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from porousfreezethaw_b200 import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("m0", (12, 10, 16), 0, 30.0), ("m1", (12, 10, 16), 1, 30.0), ("m2", (12, 10, 16), 2, 30.0),
         ("m10", (12, 10, 16), 10, 30.0), ("m11", (12, 10, 16), 11, 30.0), ("ragged_m0", (13, 7, 9), 0, 12.0)]


def main():
    for name, n, mode, T in CASES:
        spec = cases.small_case(*n, calc_mode=mode)
        case = po.Case(spec.n1, spec.n2, spec.n3, spec.L1, spec.L2, spec.L3, spec.param_array(), spec.calc_mode)
        u, p, gl = spec.initial_fields()
        f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
        for r in spec.spheres:
            f.write("%.17g %.17g %.17g\n" % tuple(r))
        f.close()
        ref = po.Reference()
        ref.setup(case, u, p, gl, f.name, t0=0.0, tau=spec.tau, tau_min=spec.tau_min, delta=spec.delta, use_callback=True)
        x0 = case.interior(ref.solution().copy())
        k1 = case.interior(ref.rhs(0.0))
        rc = ref.solve(T)
        t, h, steps, total = ref.state()
        xT = case.interior(ref.solution().copy())
        rec = ref.record()
        ref.close()
        os.unlink(f.name)
        assert rc == 0
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), n=np.array(n), mode=mode, T=T, L=np.array([spec.L1, spec.L2, spec.L3]),
                            param=spec.param_array(), spheres=spec.spheres, tau=spec.tau, tau_min=spec.tau_min, delta=spec.delta,
                            x0=x0, k1=k1[:2], xT=xT[:2], state=np.array([t, h, steps, total]), record=rec)
        print(name, "steps", steps, "total", total, "h", h)


if __name__ == "__main__":
    main()
