#!/usr/bin/env python
"""tools/tune.py -- sweep tile / z-chunk / flux-sharing settings of the stage kernels on a B200
and print Mcell-RHS/s + per-stage HBM GB/s for each (development aid; results are summarised
under profiles/)."""
import argparse
import itertools
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from porousfreezethaw_b200 import cases  # noqa: E402
from porousfreezethaw_b200.solver import SlabSolver  # noqa: E402
from bench import make_spec, STAGE_BYTES_HEAT, STAGE_BYTES_NOHEAT  # noqa: E402


def run(spec, fields, env, flags, attempts, t_h):
    for k in ("PFT_TX", "PFT_TY", "PFT_ZCHUNKS", "PFT_THREADS", "PFT_PROFILE_STAGES", "PFT_LEGACY", "PFT_SPLIT", "PFT_SHARE", "PFT_ROWS2"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    sol = SlabSolver(spec, flags=flags)
    sol.upload_interior(*fields)
    sol.sys.t, sol.sys.h = t_h
    sol.solve(spec.final_time, max_attempts=3)
    sol.reset_stats()
    e0 = sol.sys.rhs_evals
    sol.solve(spec.final_time, max_attempts=attempts)
    st = sol.stats()
    ev = sol.sys.rhs_evals - e0
    val = sol.cells * ev / (st["device_ms_total"] * 1e-3) / 1e6
    sb = STAGE_BYTES_HEAT if spec.calc_mode in (0, 1, 2) else STAGE_BYTES_NOHEAT
    per = []
    for i in range(5):
        n = st["stage_launches"][i]
        if n and st["stage_ms"][i] > 0:
            ms = st["stage_ms"][i] / n
            per.append(round(sol.cells * sb[i] / (ms * 1e-3) / 1e9))
        else:
            per.append(None)
    out = dict(env=env, flags=flags, tile=[st["tile_x"], st["tile_y"]], zc=st["z_chunk"], blocks=st["grid_blocks"],
               mcell_rhs_s=round(val, 1), ms_per_attempt=round(st["device_ms_total"] / attempts, 4), stage_gbs=per)
    sol.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="400-SigmaP1-P-smallsigma")
    ap.add_argument("--attempts", type=int, default=30)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--zsweep", action="store_true", help="sweep z-chunk counts for a few tiles only")
    args = ap.parse_args()
    spec, _, _ = make_spec(args.workload, 1)
    # a developed state: IC + sphere pack, 40 attempts in
    s0 = SlabSolver(spec)
    u, p, gl = spec.initial_fields()
    s0.upload_interior(u, p, gl)
    s0.rasterise(spec.spheres)
    s0.solve(spec.final_time, max_attempts=40)
    fields = s0.download_interior()
    t_h = (s0.sys.t, s0.sys.h)
    s0.close()
    P = {"PFT_PROFILE_STAGES": 1}
    configs = [({}, 0), ({"PFT_ROWS2": 1}, 0), (dict(P), 0)]
    tiles = [(50, 5), (50, 4), (50, 3), (40, 6), (100, 2), (66, 3), (40, 5), (50, 2)]
    if args.quick:
        tiles = tiles[:8]
    for tx, ty in tiles:
        if tx > spec.n1:
            continue
        configs.append(({"PFT_TX": tx, "PFT_TY": ty}, 0))
    if not args.quick:
        for tx, ty in ((50, 3), (50, 2)):
            for zc in (4, 8, 16, 37, 74):
                configs.append(({"PFT_TX": tx, "PFT_TY": ty, "PFT_ZCHUNKS": zc}, 0))
    if args.zsweep:
        configs = [({}, 0)]
        for tx, ty in ((50, 5), (50, 4), (50, 3), (100, 2), (40, 5)):
            if tx > spec.n1:
                continue
            for zc in (1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 20, 25, 32, 40, 50):
                if zc <= spec.n3:
                    configs.append(({"PFT_TX": tx, "PFT_TY": ty, "PFT_ZCHUNKS": zc}, 0))
    for env, flags in configs:
        try:
            print(json.dumps(run(spec, fields, env, flags, args.attempts, t_h)), flush=True)
        except Exception as e:
            print(json.dumps(dict(env=env, flags=flags, error=repr(e))), flush=True)


if __name__ == "__main__":
    main()
