"""N > 1 on CPU: world_size-2 and -3 runs of the product host code (z-slab partition, per-stage
halo exchange order, global error max) with torch.distributed/gloo standing in for NCCL and the
kernels emulated (tests/cuda_emu).  Bar: the slab-decomposed run is bit-identical to the
single-slab oracle -- the GPU-count invariance the reference shows across MPI rank counts."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import parity
from porousfreezethaw_b200 import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,mode,planes", [(2, 0, None), (3, 1, None), (2, 11, None), (3, 0, "4,9,5")])
def test_slab_decomposition_is_bit_exact(world, mode, planes, emu_lib, tmp_path):
    """one process per slab, NCCL transport (gloo standing in); planes: an uneven partition (pft_case.max_n3)"""
    T = 6.0
    spec = cases.small_case(12, 10, 18, calc_mode=mode)
    case, orc, w = parity.make(spec)
    u, p, gl = case.interior(w)
    np.savez(tmp_path / "ic.npz", u=u, p=p, gl=gl)
    s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta)
    x = w.copy()
    assert orc.solve(s, x, T) == 0
    want = case.interior(x)

    port = _free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PFT_NCCL_LIB=os.path.join(ROOT, "tests", "cuda_emu", "libnccl_emu.so"), OMP_NUM_THREADS="1",
                   PFT_TEST_PLANES=planes or "")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_gloo_worker.py"), str(mode), str(T), str(tmp_path)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [pr.communicate(timeout=600)[0].decode() for pr in procs]
    for pr, o in zip(procs, outs):
        assert pr.returncode == 0, o[-3000:]

    got = np.zeros_like(want)
    for r in range(world):
        d = np.load(tmp_path / f"rank{r}.npz")
        assert list(d["state"]) == [0, s.t, s.h, s.steps, s.steps_total], f"rank {r}: {d['state']} vs oracle {(s.t, s.h, s.steps, s.steps_total)}"
        sl = slice(int(d["first_row"]), int(d["first_row"]) + int(d["n3"]))
        got[0, sl], got[1, sl], got[2, sl] = d["u"], d["p"], d["gl"]
    assert np.array_equal(got, want)
