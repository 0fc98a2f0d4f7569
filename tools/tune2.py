#!/usr/bin/env python
"""Round-2 kernel variant sweep on one B200: runs bench.py once per environment configuration and prints one
row per configuration (Gcell-RHS/s, ms per attempt, per-stage kernel times).  Usage: tools/tune2.py [workload]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORK = sys.argv[1] if len(sys.argv) > 1 else "400-SigmaP1-P-smallsigma"
CONFIGS = [
    ("v1", {"PFT_KERNEL": "1"}),
    ("v2 R1 nobar", {}),
    ("v2 R1 bar", {"PFT_NOBAR": "0"}),
    ("v2 R1 nobar generic-shape", {"PFT_FIXED_SHAPE": "0"}),
    ("v2 R2 nobar", {"PFT_ROWS": "2"}),
    ("v2 R2 bar", {"PFT_ROWS": "2", "PFT_NOBAR": "0"}),
    ("v2 R2 ty4 (tile 50x8)", {"PFT_ROWS": "2", "PFT_TY": "4"}),
    ("v2 R2 ty3 (tile 50x6)", {"PFT_ROWS": "2", "PFT_TY": "3"}),
    ("v2 R1 ty4", {"PFT_TY": "4"}),
    ("v2 rows 21212", {"PFT_ROWS": "21212"}),
    ("v2 rows 12121", {"PFT_ROWS": "12121"}),
    ("v2 R1 nobar runahead 2", {"PFT_RUNAHEAD": "2"}),
]
if os.environ.get("TUNE_ONLY"):
    keep = os.environ["TUNE_ONLY"].split(",")
    CONFIGS = [c for c in CONFIGS if any(k in c[0] for k in keep)]
rows = []
for name, env in CONFIGS:
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", WORK, "--steps", "40", "--warmup", os.environ.get("TUNE_WARMUP", "10"), "--repeats", "3",
                        "--also", "none", "--no-cpu-baseline", "--no-parity", "--no-dropin", "--profile-steps", "20"], capture_output=True, text=True, env=e, cwd=ROOT)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if r.returncode or not line:
        print(f"{name:32s} FAILED rc={r.returncode} {r.stderr[-300:]!r}", flush=True)
        continue
    d = json.loads(line[0])
    st = " ".join(f"S{s['stage']}={1e3 * s['avg_ms']:.0f}" for s in d["roofline"]["per_stage"])
    print(f"{name:32s} {d['value'] / 1e3:7.2f} G  {d['ms_per_step']:.4f} ms/attempt  frac {d['roofline_step']['frac']:.3f}  | {st} us  tile {d['config']['tile']} zc {d['config']['z_chunk']}", flush=True)
    rows.append((name, env, d["value"]))
