#!/usr/bin/env python
"""tools/multigpu_check.py -- run under torchrun on N GPUs: the z-slab decomposed run must be
bit-identical to the single-GPU run of the same case (the invariance the reference shows across
MPI rank counts).  Prints one PASS/FAIL line per case on rank 0 and exits non-zero on failure.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py [--attempts 40]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from porousfreezethaw_b200 import cases  # noqa: E402
from porousfreezethaw_b200.solver import SlabSolver  # noqa: E402
from bench import make_spec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--attempts", type=int, default=40)
    ap.add_argument("--workloads", default="100-GradP,200-SigmaP1-P,small11")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for name in args.workloads.split(","):
        if name == "small11":
            spec = cases.small_case(24, 20, 8 * world + 3, calc_mode=11)
        else:
            spec, _, _ = make_spec(name, world)
        u, p, gl = spec.initial_fields()
        # the sphere pack, rasterised once on rank 0's GPU so that every run starts from the same bits
        if rank == 0:
            s0 = SlabSolver(spec, device=local)
            s0.upload_interior(u, p, gl)
            s0.rasterise(spec.spheres)
            u, p, gl = s0.download_interior()
            r0 = s0.solve(spec.final_time, max_attempts=args.attempts)
            ref = s0.download_interior()
            s0.close()
        t = torch.from_numpy(np.ascontiguousarray(gl)).cuda()
        dist.broadcast(t, 0)
        gl = t.cpu().numpy()

        sol = SlabSolver(spec, rank=rank, nranks=world, device=local)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(sol.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        sol.comm_init(bytes(uid.cpu().numpy().tobytes()))
        sl = slice(sol.first_row, sol.first_row + sol.n3)
        sol.upload_interior(u[sl], p[sl], gl[sl])
        r = sol.solve(spec.final_time, max_attempts=args.attempts)
        out = sol.download_interior()
        st = sol.stats()
        sol.close()
        parts = [None] * world if rank == 0 else None
        dist.gather_object((sol.first_row, sol.n3, out, (r.t, r.h, r.steps, r.steps_total)), parts, 0)
        if rank == 0:
            good = all(pt[3] == (r0.t, r0.h, r0.steps, r0.steps_total) for pt in parts)
            for first, n3, o, _ in parts:
                for a, b in zip(o, ref):
                    good = good and np.array_equal(a, b[first:first + n3])
            ok = ok and good
            print(f"{'PASS' if good else 'FAIL'} {spec.name} {spec.n1}x{spec.n2}x{spec.n3} on {world} GPUs: "
                  f"{r0.steps} accepted / {r0.steps_total} attempts, slabs bit-identical to 1 GPU: {good}; "
                  f"{st['device_ms_total'] / max(1, args.attempts):.3f} ms/attempt", flush=True)
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
