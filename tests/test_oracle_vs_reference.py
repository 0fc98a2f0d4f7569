"""Pins the oracle: oracle/pft_oracle.c against the UNMODIFIED reference compiled into
oracle/_ref/ (only where that build exists: this container and the GPU box it ships to)."""
import tempfile

import numpy as np
import pytest

from oracle import pyoracle as po
from porousfreezethaw_b200 import cases

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libpft_ref.so not built (needs /root/reference)")


def _setup(spec, variant="hybrid2", use_callback=False):
    case = po.Case(spec.n1, spec.n2, spec.n3, spec.L1, spec.L2, spec.L3, spec.param_array(), spec.calc_mode)
    u, p, gl = spec.initial_fields()
    f = tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False)
    for r in spec.spheres:
        f.write("%.17g %.17g %.17g\n" % tuple(r))
    f.close()
    ref = po.Reference(variant)
    ref.setup(case, u, p, gl, f.name, tau=spec.tau, tau_min=spec.tau_min, delta=spec.delta, use_callback=use_callback)
    orc = po.Oracle(case)
    w = case.pack(u, p, gl)
    orc.rasterise(po.read_spheres(f.name), w)
    return case, ref, orc, w


@pytest.mark.parametrize("mode", [0, 1, 2, 10, 11])
def test_rasteriser_bcond_rhs(mode):
    case, ref, orc, w = _setup(cases.small_case(14, 9, 17, calc_mode=mode))
    assert np.array_equal(case.interior(w), case.interior(ref.solution()))          # equation.c:508-529
    for t in (0.0, 20000.0):                                                          # both Dirichlet phases
        k_ref = ref.rhs(t)
        w2 = w.copy()
        k = orc.rhs(t, w2)
        assert np.array_equal(case.interior(k), case.interior(k_ref))
        assert np.array_equal(w2, ref.solution())                                     # whole ghost shell, both layers
    ref.close()


@pytest.mark.parametrize("mode,variant", [(0, "hybrid2"), (1, "hybrid2"), (2, "hybrid2"), (10, "hybrid2"), (11, "hybrid2"), (0, "hybrid")])
def test_adaptive_solve_sequence(mode, variant):
    """RK_MPI_SA_solve over three consecutive calls (the snapshot loop): identical step record,
    counters, hand-over h and fields; also against the twin module RK_MPI_SAsolver_hybrid."""
    if not po.have_ref(variant):
        pytest.skip("variant not built")
    case, ref, orc, w = _setup(cases.small_case(12, 10, 16, calc_mode=mode), variant, use_callback=True)
    rec = []
    s = orc.new_system(t=0.0, h=1.0, h_min=1e-6, delta=1e-3, service=lambda ft, t, h, st, tot, ctx: rec.append((t, h)) or 0)
    x = w.copy()
    for T in (7.0, 19.0, 30.0):
        assert ref.solve(T) == orc.solve(s, x, T) == 0
        assert ref.state() == (s.t, s.h, s.steps, s.steps_total)
    assert np.array_equal(ref.record(), np.array(rec))
    assert np.array_equal(case.interior(ref.solution()), case.interior(x))
    ref.close()


def test_no_callback_handover():
    """without a Service_Callback system->h is only written at NEXTFINISH (hybrid2.c:754)"""
    case, ref, orc, w = _setup(cases.small_case(12, 10, 16, calc_mode=0))
    s = orc.new_system(t=0.0, h=1.0, h_min=1e-6, delta=1e-3)
    x = w.copy()
    for T in (0.5, 11.0):          # the first call is a single truncated step: h stays 1.0
        assert ref.solve(T) == orc.solve(s, x, T) == 0
        assert ref.state() == (s.t, s.h, s.steps, s.steps_total)
    ref.close()


def test_nan_handling_matches():
    case, ref, orc, w = _setup(cases.small_case(10, 8, 12, calc_mode=0), use_callback=True)
    ref.lib.pftref_handle_nan(1)
    ref.set_state(0.0, 1e200, 0.0, 1e-3)
    ref.lib.pftref_break_after(6)
    s = orc.new_system(t=0.0, h=1e200, h_min=0.0, delta=1e-3, handle_nan=True,
                       service=lambda ft, t, h, st, tot, ctx: int(tot >= 6))
    x = w.copy()
    assert ref.solve(1e300) == orc.solve(s, x, 1e300)
    assert ref.state() == (s.t, s.h, s.steps, s.steps_total)
    ref.lib.pftref_handle_nan(0)
    ref.close()
