"""Worker of tests/test_multirank_gloo.py: one z-slab per process, torch.distributed/gloo standing
in for NCCL (PFT_NCCL_LIB -> tests/cuda_emu/libnccl_emu.so), the CUDA library emulated on CPU."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from porousfreezethaw_b200 import cases  # noqa: E402
from porousfreezethaw_b200.solver import SlabSolver  # noqa: E402

EMU = os.path.join(ROOT, "tests", "cuda_emu")


def main():
    mode, T, out = int(sys.argv[1]), float(sys.argv[2]), sys.argv[3]
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    nccl = C.CDLL(os.path.join(EMU, "libnccl_emu.so"), mode=C.RTLD_GLOBAL)
    P2P = C.CFUNCTYPE(None, C.c_void_p, C.c_size_t, C.c_int, C.c_int)
    GRP = C.CFUNCTYPE(None, C.c_int)
    AMX = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_size_t)
    ops = []

    def view(ptr, nbytes):
        return torch.from_numpy(np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(nbytes // 8,)))

    def p2p(buf, nbytes, peer, is_send):
        t = view(buf, nbytes)
        ops.append(dist.P2POp(dist.isend if is_send else dist.irecv, t, peer))

    def group(end):
        if not end:
            ops.clear()
            return
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        ops.clear()

    def allmax(buf, count):
        t = torch.from_numpy(np.ctypeslib.as_array(buf, shape=(count,)))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)

    hooks = (P2P(p2p), GRP(group), AMX(allmax))
    nccl.emu_nccl_set_hooks(*hooks)

    spec = cases.small_case(12, 10, 18, calc_mode=mode)
    planes = [int(v) for v in os.environ["PFT_TEST_PLANES"].split(",")] if os.environ.get("PFT_TEST_PLANES") else None
    sol = SlabSolver(spec, rank=rank, nranks=world, lib_path=os.path.join(EMU, "libpft_b200_emu.so"), planes=planes)
    uid = [sol.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    sol.comm_init(uid[0])

    # every rank builds the global IC and takes its rows (the reference evaluates the formulas per rank)
    full = np.load(os.path.join(out, "ic.npz"))
    sl = slice(sol.first_row, sol.first_row + sol.n3)
    sol.upload_interior(full["u"][sl], full["p"][sl], full["gl"][sl])
    r = sol.solve(T)
    u, p, gl = sol.download_interior()
    np.savez(os.path.join(out, f"rank{rank}.npz"), u=u, p=p, gl=gl, state=np.array([r.code, r.t, r.h, r.steps, r.steps_total]),
             first_row=sol.first_row, n3=sol.n3)
    sol.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
