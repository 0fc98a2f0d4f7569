#!/usr/bin/env python
"""tools/validate_published.py -- run an ARCHIVED production case of the reference end to end on
the B200 drop-in (the reference's own driver linked against the B200 solver module and plug-in:
oracle/_ref/intertrack_b200) and compare with what the reference published for that case:

  * the ice-volume-fraction series <case>.txt (100 values of mean(p > 0.5), scripts/avg.sh), and
  * the cumulative accepted/total Runge-Kutta step counts of OUTPUT/intertrack.log.

The published runs used another compiler/libm; the step-size controller is chaotic with respect
to last-bit differences, so step counts agree to a fraction of a percent, not exactly (SURVEY.md
facts table); the physics-level series is the robust check.

    python tools/validate_published.py Cases-LR/freeze-thaw-10h-GradP [--binary intertrack_b200]
"""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from porousfreezethaw_b200.snapshot import ice_fraction, read_snapshot  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("case")
    ap.add_argument("--binary", default="intertrack_b200")
    ap.add_argument("--snapshots", type=int, default=0, help="stop after this many snapshots (0 = all)")
    ap.add_argument("--keep", action="store_true")
    args = ap.parse_args()
    cdir = os.path.join(REF, "cases", args.case)
    work = tempfile.mkdtemp(prefix="pft_validate_")
    os.makedirs(os.path.join(work, "data"))
    os.makedirs(os.path.join(work, "OUT"))
    shutil.copy(os.path.join(REF, "data", "spheres_positions.txt"), os.path.join(work, "data", "spheres_positions.txt"))
    params = open(os.path.join(cdir, "Params")).read()
    params = re.sub(r"^set debug_logfile.*$", "# (debug log disabled for the validation run)", params, flags=re.M)
    params = re.sub(r"^set snapshot_trigger.*$", "", params, flags=re.M)
    if args.snapshots:
        # keep the snapshot schedule (intertrack.c:2271): stop early by shortening nothing -- just kill later
        pass
    open(os.path.join(work, "Params"), "w").write(params)
    env = dict(os.environ, OUTPUT="OUT", OMP_NUM_THREADS=str(os.cpu_count() or 8), OMP_SCHEDULE="static")
    t0 = time.time()
    proc = subprocess.Popen([os.path.join(REF, args.binary), "Params"], cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    pat = re.compile(r"Calculating snapshot (\d+) .*?(\d+) R-K steps \((\d+) total\)")
    ours = {}
    for line in proc.stdout:
        m = pat.search(line)
        if m:
            ours[int(m.group(1))] = (int(m.group(2)), int(m.group(3)))
            if args.snapshots and int(m.group(1)) >= args.snapshots:
                time.sleep(0.5)
                proc.terminate()
                break
    proc.wait()
    wall = time.time() - t0

    pub_steps = {int(a): (int(b), int(c)) for a, b, c in pat.findall(open(os.path.join(cdir, "OUTPUT", "intertrack.log")).read())}
    series_file = os.path.join(REF, "cases", args.case + ".txt")
    pub_series = np.loadtxt(series_file) if os.path.exists(series_file) else None
    fr = []
    k = 0
    while os.path.exists(os.path.join(work, "OUT", f"image.{k:03d}.ncd")) and (not args.snapshots or k <= args.snapshots):
        try:
            fr.append(ice_fraction(read_snapshot(os.path.join(work, "OUT", f"image.{k:03d}.ncd"))["vars"]["p"]))
        except Exception:
            break
        k += 1
    fr = np.array(fr)
    n = min(len(fr), len(pub_series)) if pub_series is not None else 0
    last = max(s for s in ours if s in pub_steps) if ours else 0
    out = {
        "case": args.case, "binary": args.binary, "wall_s": round(wall, 1), "snapshots": len(fr),
        "ice_fraction_max_abs_diff": float(np.abs(fr[:n] - pub_series[:n]).max()) if n else None,
        "ice_fraction_ours_last": float(fr[n - 1]) if n else None,
        "ice_fraction_published_last": float(pub_series[n - 1]) if n else None,
        "steps_ours": ours.get(last), "steps_published": pub_steps.get(last), "at_snapshot": last,
        "steps_total_rel_diff": (ours[last][1] - pub_steps[last][1]) / pub_steps[last][1] if last else None,
    }
    print(json.dumps(out), flush=True)
    if n:
        for i in range(0, n, max(1, n // 10)):
            print(f"  snapshot {i:3d}: ice fraction ours {fr[i]:.6f} published {pub_series[i]:.6f}")
    if not args.keep:
        shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
