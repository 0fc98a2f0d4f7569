#!/usr/bin/env python
"""tools/trace_timeline.py -- device timeline of the stage launches (development aid).

Needs `make -C porousfreezethaw_b200/csrc libpft_b200_trace.so`.  Run alone or under torchrun; prints,
per rank and launch: kernel span, gap to the previous kernel, the longest halo wait, when the boundary
planes were published relative to the kernel start, and the span of the error all-reduce kernel."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from porousfreezethaw_b200 import _lib  # noqa: E402
from porousfreezethaw_b200.solver import SlabSolver  # noqa: E402
from bench import make_spec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="400-SigmaP1-P-smallsigma")
    ap.add_argument("--attempts", type=int, default=12)
    ap.add_argument("--data", default="case", choices=["case", "trivial", "flush"])
    ap.add_argument("--cta", action="store_true", help="per-chunk CTA statistics of the first launches")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); N = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if os.environ.get("PFT_TRACE_REVERSE"):      # rank r on device N-1-r: is a slow rank slow because of its slab or its GPU?
        local = N - 1 - local
    torch.cuda.set_device(local)
    if N > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec, _, _ = make_spec(args.workload, N)
    sol = SlabSolver(spec, rank=rank, nranks=N, device=local, lib_path=_lib.TRACE_LIB_PATH)
    if N > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(sol.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        sol.comm_init(bytes(uid.cpu().numpy().tobytes()))
    u, p, gl = spec.initial_fields(first_row=sol.first_row, n3_local=sol.n3)
    if args.data == "trivial":          # same (boring) values on every rank: separates data effects from position effects
        u = np.full_like(u, 293.15); p = np.zeros_like(p)
    elif args.data == "flush":          # as the case, but the far tail of the phase field flushed to zero
        p = np.where(np.abs(p) < 1e-200, 0.0, p)
    print(f"rank {rank}: p in [{p.min():.3e}, {p.max():.3e}], zeros {np.count_nonzero(p == 0)}, 0<|p|<1e-280: {np.count_nonzero((p != 0) & (np.abs(p) < 1e-280))}", flush=True)
    sol.upload_interior(u, p, gl)
    sol.rasterise(spec.spheres)
    sol.solve(spec.final_time, max_attempts=30)
    sol.debug_trace_cta()
    sol.debug_trace()
    if N > 1:
        dist.barrier()
    sol.solve(spec.final_time, max_attempts=args.attempts)
    cta = sol.debug_trace_cta().astype(np.int64)
    stg, rec = sol.debug_trace()
    st = sol.stats()
    k1u, k1p = sol.debug_fetch(1)
    _, pf, _ = sol.download_interior()
    def bands(a):
        m = np.abs(a[a != 0])
        return (f"zeros {np.count_nonzero(a == 0)}, >=1e-289: {np.count_nonzero(m >= 1e-289)}, [1e-305,1e-289): "
                f"{np.count_nonzero((m < 1e-289) & (m >= 1e-305))}, <1e-305: {np.count_nonzero(m < 1e-305)}")
    print(f"rank {rank} bands  p: {bands(pf)} | K1p: {bands(k1p)} | K1u: {bands(k1u)}", flush=True)
    out = [f"rank {rank}: {len(stg)} launches, device {st['device_ms_total'] / args.attempts * 1e3:.1f} us/attempt"]
    r = rec.astype(np.int64)
    first = (~rec[:, 0]).astype(np.int64)
    prev_end = None
    t0 = first[0] if len(first) else 0
    for q in range(len(stg)):
        a = r[q]
        if a[1] == 0:        # skipped speculative stage 1
            out.append(f"  S{stg[q]}  skipped")
            continue
        span = (a[2] - first[q]) / 1e3
        gap = (first[q] - prev_end) / 1e3 if prev_end is not None else 0.0
        ramp = (a[1] - first[q]) / 1e3
        line = f"  S{stg[q]} t={(first[q] - t0) / 1e3:9.1f} gap={gap:6.1f} span={span:6.1f} last_start=+{ramp:5.1f}"
        if a[5]:
            line += f" maxwait={a[5] / 1e3:6.1f} pub_lo=+{(a[6] - first[q]) / 1e3 if a[6] else -1:5.1f} pub_hi=+{(a[7] - first[q]) / 1e3 if a[7] else -1:5.1f}"
        line += f" cta_body lo/hi/mid={a[11] / 1e3:5.1f}/{a[12] / 1e3:5.1f}/{a[13] / 1e3:5.1f} sig={a[10] / 1e3:4.1f}"
        prev_end = a[2]
        if a[8]:
            line += f" | eps exchange starts end+{(a[8] - a[2]) / 1e3:5.1f} takes {(a[9] - a[8]) / 1e3:5.1f}"
            prev_end = a[9]
        out.append(line)
    if args.cta:
        for q in range(min(len(cta), len(stg), 6)):
            c = cta[q]
            c = c[c[:, 0] > 0]
            if not len(c):
                continue
            t0q = c[:, 0].min()
            chunk = c[:, 3] & 0xffffffff
            sm = c[:, 3] >> 32
            out.append(f"  launch {q} S{stg[q]}: {len(c)} CTAs on {len(np.unique(sm))} SMs, span {(c[:, 2].max() - t0q) / 1e3:.1f} us")
            for ch in np.unique(chunk):
                m = c[chunk == ch]
                st_, body, sig = (m[:, 0] - t0q) / 1e3, (m[:, 1] - m[:, 0]) / 1e3, (m[:, 2] - m[:, 1]) / 1e3
                out.append(f"    chunk {ch}: n={len(m)} start {st_.min():5.1f}..{st_.max():5.1f} (mean {st_.mean():5.1f})  body mean {body.mean():5.1f} max {body.max():5.1f}"
                           f"  signal mean {sig.mean():4.1f} max {sig.max():4.1f}  end max {((m[:, 2] - t0q) / 1e3).max():5.1f}")
            early = c[(c[:, 0] - t0q) < 3000]
            late = c[(c[:, 0] - t0q) >= 3000]
            if len(late):
                out.append(f"    first wave {len(early)} CTAs body mean {((early[:, 1] - early[:, 0]) / 1e3).mean():.1f}; later {len(late)} CTAs body mean {((late[:, 1] - late[:, 0]) / 1e3).mean():.1f}")
    text = "\n".join(out)
    if N > 1:
        allt = [None] * N
        dist.all_gather_object(allt, text)
        if rank == 0:
            print("\n".join(allt))
    else:
        print(text)


if __name__ == "__main__":
    main()
