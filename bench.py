#!/usr/bin/env python
"""bench.py -- Mcell-RHS evaluations/s of the Runge-Kutta-Merson + phase-field/heat RHS hot path.

    python bench.py --gpus N --steps K --warmup W            (N=1; N>1 under torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU code on the host cores)

A "step" is one Merson step ATTEMPT over the whole grid (5 right-hand-side evaluations in the
reference; 4 + the K1 evaluations actually needed here).  The workload is an
archived case of the reference (default: the "400" high-resolution case, 200x200x400 nodes,
calc_mode 1) or a synthetic sphere-pack cube; see --workload.

Prints ONE JSON line (rank 0).  `value` times the device-resident path, `e2e` the same K
attempts through upload -> solve -> download with pinned HOST buffers, `roofline` the dominant
stage kernel against the measured HBM peak, `cpu_baseline` the reference CPU code on this host.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from porousfreezethaw_b200 import cases  # noqa: E402

# algorithmic HBM bytes per cell of each fused stage kernel:
# float64, u and p evolve (16 B per array pair), gl static (8 B), one sweep per stage
STAGE_BYTES_HEAT = [40, 56, 40, 72, 88]      # sum 296 B per attempt = 59.2 B per cell-RHS
STAGE_BYTES_NOHEAT = [24, 32, 24, 40, 48]    # calc_mode 10/11: only p evolves; u, gl static reads


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_spec(name: str, nranks: int):
    if name.startswith("params:"):
        # any reference `Params` case file, through the front end of porousfreezethaw_b200/params.py
        from porousfreezethaw_b200 import params as pfront
        spec, pc = pfront.case_from_params(name[7:])
        sph = os.path.join(os.path.dirname(os.path.abspath(name[7:])), "data", "spheres_positions.txt")
        if not os.path.exists(sph):
            sph = os.path.join(ROOT, "oracle", "_ref", "data", "spheres_positions.txt")
        spec.spheres = np.loadtxt(sph).reshape(-1, 3) if os.path.exists(sph) else cases.synthetic_spheres(200, zmax=1.6)
        if isinstance(pc.options.get("icond_file"), str):        # `set icond_file=...` [`set continue_series`]
            path = pc.options["icond_file"]
            if not os.path.isabs(path):
                path = os.path.join(os.path.dirname(os.path.abspath(name[7:])), path)
            r = pfront.restart_from_snapshot(spec, path, continue_series=bool(pc.options.get("continue_series")))
            u, p, gl = r.u, r.p, r.gl
            spec.extra["t0"] = r.t
            if r.tau is not None:
                spec.tau, spec.final_time = r.tau, r.final_time
        else:
            u, p, gl = pfront.initial_fields_from_formulas(spec, pc.icond)
        spec.initial_fields = lambda first_row=0, n3_local=None, _f=(u, p, gl), _s=spec: tuple(
            a[first_row:first_row + (_s.n3 if n3_local is None else n3_local)] for a in _f)
        return spec, "strong", f"case file {name[7:]}"
    if name.startswith("slab-"):
        # slab-200x50: one thin 200x200x50 block, calc_mode 1 (what one of 8 ranks owns of the 400 case)
        n, nz = (int(v) for v in name[5:].split("x"))
        spec = cases.cube_case(n, nz=nz, calc_mode=1, spheres=200)
        return spec, "strong", "synthetic sphere pack (numpy default_rng(20251018), 200 spheres)"
    if name.startswith("cube-"):
        # cube-400 | cube-800 (strong scaling) | cube-800w (weak: 800x800x(100*N))
        tag = name[5:]
        if tag.endswith("w"):
            n = int(tag[:-1])
            spec = cases.cube_case(n, nz=100 * nranks)
            scaling = "weak"
        else:
            n = int(tag)
            spec = cases.cube_case(n)
            scaling = "strong"
        return spec, scaling, "synthetic sphere pack (numpy default_rng(20251018), 4000 spheres)"
    spec = cases.archived_case(name)
    sph = os.path.join(ROOT, "oracle", "_ref", "data", "spheres_positions.txt")
    if os.path.exists(sph):
        spec.spheres = np.loadtxt(sph).reshape(-1, 3)
        geo = "reference sphere pack (data/spheres_positions.txt, 200 spheres)"
    else:
        s = cases.synthetic_spheres(200, zmax=1.6)
        spec.spheres = s
        geo = "synthetic sphere pack (200 spheres; the reference's data file is not in this tree)"
    return spec, "strong", geo


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons, pw = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); smax.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# the reference CPU path (oracle/_ref when the reference was compiled in the container,
# else our restatement oracle/pft_oracle.c) -- used by cpu_baseline and by --impl reference

class _StdoutToStderr:
    """The reference's C code prints progress on fd 1; the bench line must be the only thing on stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference_run(spec, warmup: int, steps: int, variant: str = "hybrid2", threads: int | None = None):
    with _StdoutToStderr():
        return _cpu_reference_run(spec, warmup, steps, variant, threads)


def _cpu_reference_run(spec, warmup: int, steps: int, variant: str = "hybrid2", threads: int | None = None):
    """Times `steps` Merson attempts of the reference's own CPU implementation (after `warmup`
    attempts) with all host threads (or `threads`).  variant "hybrid2" = the reference's own flags
    (-O1), "fast" = the same sources at -O3 -march=x86-64-v3.  Returns (Mcell-RHS/s, attempts,
    seconds, dict describing the run)."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    if threads is not None:       # the OpenMP runtime is already up: change the team size through its API
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    os.environ.setdefault("OMP_SCHEDULE", "static")      # apps/intertrack-hybrid-S-freezing/Run:13-19
    os.environ.setdefault("OMP_PROC_BIND", "false")
    po.build("oracle")
    case = po.Case(spec.n1, spec.n2, spec.n3, spec.L1, spec.L2, spec.L3, spec.param_array(), spec.calc_mode)
    u, p, gl = spec.initial_fields()
    orc = po.Oracle(case)
    w = case.pack(u, p, gl)
    orc.rasterise(spec.spheres, w)            # OpenMP rasteriser of the restatement (bit-identical to the reference's)
    if po.have_ref(variant):
        kind = "reference"
        ref = po.Reference(variant)
        gl_r = case.interior(w)[2]
        ref.setup(case, u, p, gl_r, None, t0=0.0, tau=spec.tau, tau_min=spec.tau_min, delta=spec.delta, use_callback=True)
        del w

        def run_attempts(n):
            t0_, _, _, tot0 = ref.state()
            ref.lib.pftref_break_after(tot0 + n)
            t = time.perf_counter()
            ref.solve(spec.final_time)
            dt = time.perf_counter() - t
            return ref.state()[3] - tot0, dt
    else:
        kind = "port"
        s = orc.new_system(t=0.0, h=spec.tau, h_min=spec.tau_min, delta=spec.delta)
        work = np.zeros(5 * w.size)

        def run_attempts(n):
            tot0 = s.steps_total
            s.max_attempts = n
            t = time.perf_counter()
            orc.solve(s, w, spec.final_time, work)
            dt = time.perf_counter() - t
            return s.steps_total - tot0, dt
    if warmup > 0:
        run_attempts(warmup)
    n, dt = run_attempts(steps)
    value = spec.cells * 5.0 * n / dt / 1e6
    flags = "-O3 -march=x86-64-v3 -fopenmp" if variant == "fast" else "reference flags -O1 -fopenmp"
    info = {"value": value, "unit": "Mcell-RHS/s", "cores": int(threads if threads is not None else os.environ["OMP_NUM_THREADS"]), "kind": kind,
            "sample": f"{n} Merson attempts (5 RHS each) of {spec.name} {spec.n1}x{spec.n2}x{spec.n3} after {warmup} warm-up attempts, "
                      f"{flags}, {dt:.2f} s"}
    if threads is not None:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(os.environ["OMP_NUM_THREADS"]))
    return value, n, dt, info


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    spec, scaling, geo = make_spec(args.workload, max(1, args.gpus))
    value, n, dt, info = cpu_reference_run(spec, args.warmup, args.steps)
    line = {
        "impl": "reference", "metric": "Mcell-RHS evals/s", "value": value, "unit": "Mcell-RHS/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(n, 1),
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": spec.name, "grid": [spec.n1, spec.n2, spec.n3], "calc_mode": spec.calc_mode, "geometry": geo,
                   "mode": "adaptive"},
        "cpu_baseline": info,
        "e2e": {"value": value, "unit": "Mcell-RHS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default=os.environ.get("PFT_BENCH_WORKLOAD", "400-SigmaP1-P-smallsigma"))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--variant", default="strict", choices=["strict", "fast"],
                    help="strict = IEEE association, bit-exact (default); fast = FMA contraction allowed (fields within 1e-10, no step-sequence parity)")
    ap.add_argument("--fixed-step", action="store_true",
                    help="SURVEY.md 8(d) fixed-step mode: H = half the adaptive step reached after the warm-up, h_min = inf, one solve call per step")
    ap.add_argument("--cpu-steps", type=int, default=50)
    ap.add_argument("--profile-steps", type=int, default=40)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3          # the timing rules ask for W >= 3 (and 0 would mean "no attempt limit" to pft_solve)
    if args.steps < 1:
        args.steps = 1

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from porousfreezethaw_b200.solver import SlabSolver

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # TEST INFRASTRUCTURE ONLY (tests/test_host_logic.py): walk through this script on a machine without a
    # GPU against the CPU emulation of the CUDA sources (tests/cuda_emu), to check the line it prints.
    # The line says so in "data"; nothing measured that way is a benchmark.
    dry = os.environ.get("PFT_BENCH_DRYRUN_LIB")
    if dry:
        if world != 1:
            raise SystemExit("bench.py: the dry run is single-rank")
        print("bench.py: DRY RUN against the CPU emulation (PFT_BENCH_DRYRUN_LIB) -- test of this script, not a measurement",
              file=sys.stderr)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; porousfreezethaw_b200 has no CPU path")
        torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"bench.py: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    N = world

    spec, scaling, geo = make_spec(args.workload, N)
    from porousfreezethaw_b200 import _lib as _pftlib
    lib_path = _pftlib.FAST_LIB_PATH if args.variant == "fast" else None
    if os.environ.get("PFT_BENCH_LIB"):       # A/B builds (development)
        lib_path = os.environ["PFT_BENCH_LIB"]
    if dry:
        lib_path = dry
    sol = SlabSolver(spec, rank=rank, nranks=N, device=local, lib_path=lib_path)
    if N > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(sol.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        sol.comm_init(bytes(uid.cpu().numpy().tobytes()))

    # initial condition of this rank's slab, sphere pack rasterised on the device
    u, p, gl = spec.initial_fields(first_row=sol.first_row, n3_local=sol.n3)
    sol.upload_interior(u, p, gl)
    sol.rasterise(spec.spheres)
    del u, p, gl
    sol.sys.t, sol.sys.h = float(spec.extra.get("t0", 0.0)), spec.tau      # a restart dataset may set both

    def barrier():
        if dry:
            return
        torch.cuda.synchronize()
        if N > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x: float) -> float:
        if N == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x: float) -> float:
        if N == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    T_END = spec.final_time

    # ---- warm-up, then EXACTLY K attempts, device-resident --------------------------------
    sol.solve(T_END, max_attempts=args.warmup)
    H = 0.5 * sol.sys.h                                      # fixed-step mode: identical on every rank

    def advance(s, k):
        """k step attempts; returns this rank's device time (CUDA events on the solver's stream)"""
        if not args.fixed_step:
            s.solve(T_END, max_attempts=k)
            return s.stats()["device_ms_total"]
        ms = 0.0
        for _ in range(k):                                   # one always-accepted Merson step per call
            s.sys.h, s.sys.h_min = H, float("inf")
            s.solve(s.sys.t + H)
            ms += s.stats()["device_ms_total"]
        return ms

    if args.fixed_step:
        advance(sol, 3)
    st0 = sol.stats()
    ev0 = sol.sys.rhs_evals
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall = time.perf_counter()
    my_ms = advance(sol, args.steps)
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop() if sampler else None
    st1 = sol.stats()
    dev_ms = allmax(my_ms)                                   # max over ranks
    if dry and dev_ms <= 0:
        dev_ms = 1e3 * t_wall                                # the emulation has no device clock
    attempts = st1["attempts"] - st0["attempts"]
    rhs = sol.sys.rhs_evals - ev0
    total_cells = allsum(float(sol.cells))
    value = total_cells * rhs / (dev_ms * 1e-3) / 1e6
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    stage_launches = [a - b for a, b in zip(st1["stage_launches"], st0["stage_launches"])]
    # host side of the control loop, per attempt: time blocked on the device's decision / time issuing launches
    host_wait = (st1["host_wait_ms"] - st0["host_wait_ms"]) / max(attempts, 1)
    host_launch = (st1["host_launch_ms"] - st0["host_launch_ms"]) / max(attempts, 1)
    if N > 1:
        hw = [None] * N
        dist.all_gather_object(hw, (round(host_wait, 4), round(host_launch, 4), round(my_ms / max(attempts, 1), 4)))
    else:
        hw = [(round(host_wait, 4), round(host_launch, 4), round(my_ms / max(attempts, 1), 4))]

    # ---- per-stage device times (CUDA events around every stage kernel on its stream) ----
    os.environ["PFT_PROFILE_STAGES"] = "1"
    # (a second solver object: the profiling switch is read when a solver is created)
    u_, p_, gl_ = sol.download_interior()
    solp = SlabSolver(spec, rank=rank, nranks=N, device=local, lib_path=lib_path)
    if N > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(solp.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        solp.comm_init(bytes(uid.cpu().numpy().tobytes()))
    solp.upload_interior(u_, p_, gl_)
    solp.sys.t, solp.sys.h = sol.sys.t, sol.sys.h
    del u_, p_, gl_
    solp.sys.h_min = sol.sys.h_min
    advance(solp, 3)
    solp.reset_stats()
    barrier()
    advance(solp, args.profile_steps)
    barrier()
    stp = solp.stats()
    del os.environ["PFT_PROFILE_STAGES"]
    sb = STAGE_BYTES_HEAT if spec.calc_mode in (0, 1, 2) else STAGE_BYTES_NOHEAT
    peak, peak_src = peaks()
    stages = []
    for i in range(5):
        nl = stp["stage_launches"][i]
        if nl == 0:
            continue
        ms = stp["stage_ms"][i] / nl
        if dry and ms <= 0:
            ms = 1e-3
        gbs = sol.cells * sb[i] / (ms * 1e-3) / 1e9
        stages.append({"stage": i + 1, "launches": nl, "avg_ms": ms, "bytes_per_cell": sb[i], "achieved_gbs": gbs, "frac": gbs / peak})
    dom = max(stages, key=lambda s: s["avg_ms"] * s["launches"])
    roofline = {"bound": "hbm", "kernel": f"merson_stage<{dom['stage']}>", "achieved": dom["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": dom["frac"], "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": sol.cells * dom["bytes_per_cell"], "avg_launch_ms": dom["avg_ms"],
                "per_stage": stages}
    tr = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu capture
    if os.path.exists(tr):
        try:
            with open(tr) as f:
                tj = json.load(f)
            key = f"{spec.name}:stage{dom['stage']}"
            if key in tj:
                roofline["traffic"] = tj[key]
        except Exception:
            pass
    solp.close()

    # ---- end to end: pinned HOST buffers -> upload -> K attempts -> download ---------------
    xh = np.empty(3 * sol.SIZE) if dry else torch.empty(3 * sol.SIZE, dtype=torch.float64, pin_memory=True).numpy()
    sol.download(xh)                                         # a valid padded host image of the current state
    barrier()
    ev0 = sol.sys.rhs_evals
    a0 = sol.sys.steps_total
    t0 = time.perf_counter()
    sol.upload(xh)
    advance(sol, args.steps)
    sol.download(xh)
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_attempts = sol.sys.steps_total - a0
    e2e_value = total_cells * (sol.sys.rhs_evals - ev0) / e2e_s / 1e6
    nbytes = 3 * sol.SIZE * 8
    e2e = {"value": e2e_value, "unit": "Mcell-RHS/s", "h2d_bytes_per_step": nbytes / max(e2e_attempts, 1),
           "d2h_bytes_per_step": nbytes / max(e2e_attempts, 1), "attempts": e2e_attempts,
           "note": "one solve call = upload(x) + K attempts + download(x), as RK_MPI_SA_solve with a host-resident x"}

    # ---- the reference CPU path on this host (rank 0, N == 1 only) ------------------------
    cpu = None
    if rank == 0 and N == 1 and not args.no_cpu_baseline:
        try:
            _, _, _, cpu = cpu_reference_run(spec, 1, args.cpu_steps)
            # also the aggressive-flags build and one thread, on short samples
            others = []
            from oracle import pyoracle as _po
            if _po.have_ref("fast"):
                others.append(cpu_reference_run(spec, 1, max(2, args.cpu_steps // 5), variant="fast")[3])
            if _po.have_ref("hybrid"):      # a second copy of the reference (own globals) for the 1-thread sample
                others.append(cpu_reference_run(spec, 0, 1, variant="hybrid", threads=1)[3])
            cpu["others"] = others
        except Exception as e:      # the baseline is reported, never fatal
            cpu = {"value": None, "unit": "Mcell-RHS/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        st = sol.stats()
        line = {
            "metric": "Mcell-RHS evals/s", "value": value, "unit": "Mcell-RHS/s", "n_gpus": N,
            "steps": attempts, "warmup": args.warmup, "ms_per_step": dev_ms / max(attempts, 1),
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
            "data": "DRY RUN on the CPU emulation (tests only): not a measurement" if dry else "synthetic",
            "config": {"workload": spec.name, "grid": [spec.n1, spec.n2, spec.n3], "calc_mode": spec.calc_mode,
                       "geometry": geo, "mode": f"fixed step H={H:.6g}" if args.fixed_step else "adaptive (delta 1e-3)", "variant": args.variant, "cache": "inputs larger than L2 (no flush needed)"
                       if sol.cells * 13 * 8 > 2 * 126e6 else "working set fits L2",
                       "parallelism": f"z-slabs x{N}", "tile": [st["tile_x"], st["tile_y"]], "z_chunk": st["z_chunk"],
                       "rhs_evals_per_attempt": rhs / max(attempts, 1),
                       "accepted": st1["accepted"] - st0["accepted"]},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "wall_ms_per_step": 1e3 * t_wall / max(attempts, 1),
            "host_loop_ms_per_step": {"per_rank[wait, launch, device]": hw},
            "roofline_step": {"achieved": total_cells / N * sum(b * n for b, n in zip(sb, stage_launches)) / (dev_ms * 1e-3) / 1e9,
                              "peak": peak, "unit": "GB/s",
                              "frac": total_cells / N * sum(b * n for b, n in zip(sb, stage_launches)) / (dev_ms * 1e-3) / 1e9 / peak,
                              "note": "algorithmic bytes of every stage kernel launched in the timed region / device time, per GPU (includes launch gaps and host round trips)"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    sol.close()
    if N > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
