"""The Params front end (porousfreezethaw_b200/params.py) against the reference's own parsers:
operator precedences, the archived case files, and the initial-condition pass compared with
snapshot 0 written by the reference application."""
import glob
import os
import subprocess

import numpy as np
import pytest

from porousfreezethaw_b200 import cases, params
from porousfreezethaw_b200.snapshot import read_snapshot

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def test_expression_grammar():
    v = dict(L1=0.03, L2=0.03, L3=0.06, a=2.0, b=3.0, x=np.array([0.0, 1.0, 2.0]))
    ev = lambda s: params.evaluate(s, v)
    assert ev("1+2*3") == 7 and ev("(1+2)*3") == 9 and ev("2^3^1") == 8 and ev("-2^2") == -4
    assert ev("10 / (L1 max L2 max L3)") == 10 / 0.06
    assert ev("a*b max 10") == 2.0 * 10            # max (16) binds tighter than * (20)
    assert ev("5*60*60") == 18000 and ev("273.15 - 25") == 273.15 - 25
    assert list(ev("x>0.5 and x<1.5")) == [0.0, 1.0, 0.0]
    assert list(ev("x=1 or not (x<2)")) == [0.0, 1.0, 1.0]
    assert list(ev("x>0.5 ? 7 : 9")) == [9.0, 7.0, 7.0]
    assert ev("0.5*(1.0 + tanh(0.5/0.1*(0.2-0.055))") == 0.5 * (1.0 + np.tanh(0.5 / 0.1 * (0.2 - 0.055)))   # unclosed paren tolerated
    assert ev("sqrt(16) + abs(-2) + sgn(-3)") == 5.0
    with pytest.raises(params.ExprError):
        ev("nope + 1")
    assert params.to_int(199.99999999999997) == 200 and params.to_int(2.4) == 2


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "cases")), reason="archived case files not staged")
def test_archived_cases_parse_and_match_tables():
    files = sorted(glob.glob(os.path.join(REF, "cases", "Cases-*", "*", "Params"))) + [os.path.join(REF, "cases", "Params.shipped")]
    assert len(files) >= 10
    for f in files:
        spec, pc = params.case_from_params(f)
        assert spec.n1 * 2 == spec.n2 * 2 == spec.n3 and spec.delta == 1e-3 and spec.tau == 1.0 and spec.tau_min == 1e-6
        assert set(pc.icond) == {"u", "p", "gl"} and "logfile" in pc.options
    for name, rel in (("100-GradP", "Cases-LR/freeze-thaw-10h-GradP"), ("200-GradP", "Cases-MR/freeze-thaw-10h-GradP"),
                      ("200-Temp", "Cases-MR/freeze-thaw-10h-Temp"),
                      ("400-SigmaP1-P-smallsigma", "Cases-HR/freeze-thaw-10h-SigmaP1-P-smallsigma"), ("400-Temp", "Cases-HR/freeze-thaw-10h-Temp")):
        spec, _ = params.case_from_params(os.path.join(REF, "cases", rel, "Params"))
        tab = cases.archived_case(name)
        assert (spec.n1, spec.n2, spec.n3, spec.calc_mode) == (tab.n1, tab.n2, tab.n3, tab.calc_mode)
        assert np.array_equal(spec.param_array(), tab.param_array()), name
        assert spec.final_time == tab.final_time and spec.saved_files == tab.saved_files


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "intertrack_ref")), reason="reference application not built")
def test_initial_conditions_match_reference_snapshot0(tmp_path):
    """snapshot 0 of the reference application = IC formulas + sphere pack; the formula pass must agree"""
    text = open(os.path.join(REF, "cases", "Params.shipped")).read()
    text = text.replace("grid_nodes\t100", "grid_nodes\t24").replace("saved_files\t100", "saved_files\t2").replace("final_time\t10*hours", "final_time\t1e-6")
    text = "\n".join(l for l in text.splitlines() if not l.startswith("set debug_logfile") and not l.startswith("set snapshot_trigger"))
    (tmp_path / "Params").write_text(text)
    (tmp_path / "data").mkdir(); (tmp_path / "OUT").mkdir()
    open(tmp_path / "data" / "spheres_positions.txt", "w").close()          # empty pack: pure formula pass
    r = subprocess.run([os.path.join(REF, "intertrack_ref"), "Params"], cwd=tmp_path, env=dict(os.environ, OUTPUT="OUT", OMP_NUM_THREADS="2"),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    snap = read_snapshot(str(tmp_path / "OUT" / "image.000.ncd"))
    spec, pc = params.case_from_params(str(tmp_path / "Params"))
    assert (spec.n3, spec.n2, spec.n1) == snap["vars"]["u"].shape
    u, p, gl = params.initial_fields_from_formulas(spec, pc.icond)
    assert np.array_equal(u, snap["vars"]["u"]) and np.array_equal(p, snap["vars"]["p"])
    assert np.abs(gl - snap["vars"]["gl"]).max() <= 4 * np.finfo(float).eps       # numpy tanh vs glibc tanh
    for k in ("xi", "alpha", "beads_scaling", "ball_radius", "top_temp1", "phase_switch_time"):
        assert snap["attrs"][k] == spec.params[k], k
