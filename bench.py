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
    # all host cores, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1): the environment for a runtime
    # that is not up yet, the API for one that is
    os.environ["OMP_NUM_THREADS"] = str(cores)
    import ctypes
    ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL).omp_set_num_threads(int(threads if threads is not None else cores))
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
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
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

def plane_sums(a):
    """per z-plane sum (mod 2^64) of the float64 bit patterns (tests/golden/make_golden_partial.py)"""
    return np.ascontiguousarray(a, dtype="<f8").view(np.uint64).reshape(a.shape[0], -1).sum(axis=1, dtype=np.uint64)


class Bench:
    """one workload on this rank's slab: creation, initial state, the timed regions"""

    def __init__(self, args, spec, rank, N, local, lib_path, dry, dist, torch):
        from porousfreezethaw_b200.solver import SlabSolver
        self.args, self.spec, self.rank, self.N, self.local, self.dry, self.dist, self.torch = args, spec, rank, N, local, dry, dist, torch
        self.lib_path = lib_path
        self.SlabSolver = SlabSolver
        self.planes = None                     # planes per slab; None = the reference's rule
        self.balance = None
        if N > 1 and not dry and not getattr(args, "no_balance", False):
            self.calibrate()
        self.sol = self.new_solver()

    def calibrate(self):
        """Load balance (the multi-process counterpart of pft_group_balance): slabs differ in cost per plane (interior slabs
        push two halo planes, the division's cost depends on the data) and every attempt ends with all slabs waiting for the
        slowest.  A throw-away solver with per-stage device timing runs warm-up + 30 attempts from the initial state; the
        planes are then dealt out in inverse proportion to the measured cost per plane (cases.balanced_planes)."""
        from porousfreezethaw_b200.cases import balanced_planes, partition
        os.environ["PFT_PROFILE_STAGES"] = "1"
        sol = self.new_solver()
        del os.environ["PFT_PROFILE_STAGES"]
        try:
            self.initial_state(sol)
            sol.solve(self.spec.final_time, max_attempts=max(self.args.warmup, 3))
            sol.reset_stats()
            self.barrier()
            sol.solve(self.spec.final_time, max_attempts=30)
            st = sol.stats()
        finally:
            sol.close()
        mine = sum(st["stage_ms"][i] / max(st["stage_launches"][i], 1) for i in range(4))      # stage 5 contains the wait for the others
        allc = [None] * self.N
        self.dist.all_gather_object(allc, float(mine))
        planes0 = [partition(self.spec.n3, self.N, r)[0] for r in range(self.N)]
        self.balance = {"cost_ms_stages_1_4": [round(c, 4) for c in allc], "planes_before": planes0}
        if max(allc) > 1.03 * (sum(allc) / self.N) and min(allc) > 0:
            self.planes = balanced_planes(planes0, allc)
        self.balance["planes"] = self.planes or planes0

    def new_solver(self):
        sol = self.SlabSolver(self.spec, rank=self.rank, nranks=self.N, device=self.local, lib_path=self.lib_path, planes=self.planes)
        if self.N > 1:
            torch, dist = self.torch, self.dist
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                uid.copy_(torch.frombuffer(bytearray(sol.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            sol.comm_init(bytes(uid.cpu().numpy().tobytes()))
        return sol

    def barrier(self):
        if self.dry:
            return
        self.torch.cuda.synchronize()
        if self.N > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def allreduce(self, x: float, op: str) -> float:
        if self.N == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def initial_state(self, sol, device_rasteriser=True):
        spec = self.spec
        u, p, gl = spec.initial_fields(first_row=sol.first_row, n3_local=sol.n3)
        sol.upload_interior(u, p, gl)
        if device_rasteriser:
            t0 = time.perf_counter()
            sol.rasterise(spec.spheres)
            self.rasterise_ms = 1e3 * (time.perf_counter() - t0)       # device rasteriser (spheres binned), incl. host-side binning and sync
        sol.sys.t, sol.sys.h = float(spec.extra.get("t0", 0.0)), spec.tau      # a restart dataset may set both
        sol.sys.steps = sol.sys.steps_total = 0

    def advance(self, s, k, H=None):
        """k step attempts; returns this rank's device time (CUDA events on the solver's stream)"""
        if H is None:
            s.solve(self.spec.final_time, max_attempts=k)
            return s.stats()["device_ms_total"]
        ms = 0.0
        for _ in range(k):                                   # one always-accepted Merson step per call
            T = s.sys.t + H
            s.sys.h, s.sys.h_min = T - s.sys.t, float("inf")   # h_eff = fl(fl(t+H) - t): |T - t| <= |h| holds, so the call is ONE attempt
            s.solve(T)
            ms += s.stats()["device_ms_total"]
        return ms

    def timed(self, s, k, repeats, H=None):
        """`repeats` timed regions of exactly k attempts each, every one bracketed by a barrier and a
        synchronize on both sides; per region the MAX over ranks of the device time; returns the list"""
        out = []
        for _ in range(repeats):
            self.barrier()
            t0 = time.perf_counter()
            my_ms = self.advance(s, k, H)
            if not self.dry:
                self.torch.cuda.synchronize()
            my_wall = 1e3 * (time.perf_counter() - t0)
            self.barrier()
            dev = self.allreduce(my_ms, "max")
            wall = self.allreduce(my_wall, "max")
            if self.dry and dev <= 0:
                dev = wall                                   # the emulation has no device clock
            out.append((dev, wall, my_ms))
        return out


def parity_check(b: Bench, golden_path: str):
    """The benchmarked workload against the UNMODIFIED reference (tests/golden/partial_<case>.npz: state and
    per-plane checksums of u and p after the first n attempts): fresh initial state with the sphere pack
    rasterised by libm tanh on the host (the drop-in plug-in's rasteriser), n attempts on the device(s),
    every rank checks its own planes."""
    from porousfreezethaw_b200 import _lib
    g = np.load(golden_path)
    n = int(g["break_after"])
    spec, sol0 = b.spec, b.sol
    sol = b.new_solver()
    try:
        u, p, gl = spec.initial_fields(first_row=sol.first_row, n3_local=sol.n3)
        w = np.zeros((3, sol.N3, sol.N2, sol.N1))
        for q, a in enumerate((u, p, gl)):
            w[q, 2:-2, 2:-2, 2:-2] = a
        w = w.reshape(-1)
        try:      # torchrun exports OMP_NUM_THREADS=1: give the host rasteriser this rank's share of the cores
            import ctypes
            ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL).omp_set_num_threads(max(1, (os.cpu_count() or 1) // max(b.N, 1)))
        except OSError:
            pass
        _lib.rasterise_host(sol._case, spec.spheres, w)
        gl_r = w.reshape(3, sol.N3, sol.N2, sol.N1)[2, 2:-2, 2:-2, 2:-2]
        sol.upload_interior(u, p, np.ascontiguousarray(gl_r))
        sol.sys.t, sol.sys.h = 0.0, spec.tau
        r = sol.solve(spec.final_time, max_attempts=n)
        uo, po_, glo = sol.download_interior()
        sl = slice(sol.first_row, sol.first_row + sol.n3)
        mine = (r.t == float(g["state"][0]) and r.h == float(g["state"][1]) and r.steps == int(g["state"][2]) and r.steps_total == int(g["state"][3])
                and np.array_equal(plane_sums(uo), g["plane_sum_u"][sl]) and np.array_equal(plane_sums(po_), g["plane_sum_p"][sl])
                and np.array_equal(plane_sums(glo), g["plane_sum_gl"][sl]))
        ok = b.allreduce(0.0 if mine else 1.0, "sum") == 0.0
        return {"ok": bool(ok), "against": "unmodified reference, " + os.path.relpath(golden_path, ROOT), "attempts": n,
                "state": {"t": r.t, "h": r.h, "steps": r.steps, "steps_total": r.steps_total},
                "reference_state": {"t": float(g["state"][0]), "h": float(g["state"][1]), "steps": int(g["state"][2]), "steps_total": int(g["state"][3])},
                "fields": "per-plane checksums of u, p, gl: every rank its own planes, bit-exact"}
    finally:
        sol.close()


def dropin_e2e(spec, ngpus: int, t_probe: float):
    """End to end THROUGH THE REFERENCE'S SOLVER-MODULE ABI: RK_MPI_SA_solve(final_time, &system) of librk_b200.so with
    the padded host array `system->x` (pageable numpy memory, page-locked by the module on first use), the chunk
    table of intertrack.c:2149-2156 and the plug-in's host rasteriser -- exactly the calls the reference driver
    makes.  ngpus > 1: this ONE process drives all of them (PFT_NGPUS).  The first call (page-locking, glass field
    upload) is not timed; the second one is."""
    from porousfreezethaw_b200.dropin import DropInRun
    run = DropInRun(spec, ngpus)
    try:
        u, p, gl = spec.initial_fields()
        w = np.zeros((3, run.N3, run.N2, run.N1))
        for q, a in enumerate((u, p, gl)):
            w[q, 2:-2, 2:-2, 2:-2] = a
        del u, p, gl
        w = w.reshape(-1)
        run.rasterise_host(w)
        rc = run.solve(w, 0.25 * t_probe)
        if rc != 0:
            raise RuntimeError(f"RK_MPI_SA_solve returned {rc}")
        a0, s0 = run.system.steps_total, run.system.steps
        t0 = time.perf_counter()
        rc = run.solve(w, t_probe)
        dt = time.perf_counter() - t0
        if rc != 0:
            raise RuntimeError(f"RK_MPI_SA_solve returned {rc}")
        att, acc = run.system.steps_total - a0, run.system.steps - s0
        rhs = 4 * att + acc
        nbytes = 2 * spec.n3 * run.N1 * run.N2 * 8
        return {"value": spec.cells * rhs / dt / 1e6, "unit": "Mcell-RHS/s", "gpus": ngpus, "attempts": att, "seconds": dt,
                "h2d_bytes_per_step": nbytes / max(att, 1), "d2h_bytes_per_step": nbytes / max(att, 1),
                "note": "one RK_MPI_SA_solve call of librk_b200.so (solver-module ABI of include/RK_MPI_SAsolver.h) on a host-resident x"}
    finally:
        run.close()


def stage_profile(b: Bench, H, peak, peak_src):
    """per-stage device times: CUDA events around every stage kernel on its stream (programmatic dependent
    launch off, one attempt in flight).  With several ranks a stage's time includes waiting for the slower
    neighbour -- use the numbers as kernel times at N = 1 only."""
    spec = b.spec
    os.environ["PFT_PROFILE_STAGES"] = "1"
    solp = b.new_solver()          # (the profiling switch is read when a solver is created)
    del os.environ["PFT_PROFILE_STAGES"]
    try:
        u_, p_, gl_ = b.sol.download_interior()
        solp.upload_interior(u_, p_, gl_)
        del u_, p_, gl_
        solp.sys.t, solp.sys.h, solp.sys.h_min = b.sol.sys.t, b.sol.sys.h, b.sol.sys.h_min
        b.advance(solp, 3, H)
        solp.reset_stats()
        b.barrier()
        b.advance(solp, b.args.profile_steps, H)
        b.barrier()
        stp = solp.stats()
    finally:
        solp.close()
    sb = STAGE_BYTES_HEAT if spec.calc_mode in (0, 1, 2) else STAGE_BYTES_NOHEAT
    stages = []
    for i in range(5):
        nl = stp["stage_launches"][i]
        if nl == 0:
            continue
        ms = stp["stage_ms"][i] / nl
        if b.dry and ms <= 0:
            ms = 1e-3
        gbs = b.sol.cells * sb[i] / (ms * 1e-3) / 1e9
        stages.append({"stage": i + 1, "launches": nl, "avg_ms": ms, "bytes_per_cell": sb[i], "achieved_gbs": gbs, "frac": gbs / peak})
    dom = max(stages, key=lambda s: s["avg_ms"] * s["launches"])
    gen = {0: "merson_stage", 1: "merson_stage_tma", 2: "merson_stage_v2"}.get(stp.get("kernel_generation", 1), "merson_stage")
    roofline = {"bound": "hbm", "kernel": f"{gen}<{dom['stage']}>", "achieved": dom["achieved_gbs"], "peak": peak,
                "unit": "GB/s", "frac": dom["frac"], "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": b.sol.cells * dom["bytes_per_cell"], "avg_launch_ms": dom["avg_ms"],
                "per_stage": stages}
    if b.N > 1:
        roofline["note"] = "N > 1: a stage's event time includes waiting for the neighbouring slabs' boundary planes (stage 5: for every slab's error maximum)"
        mine = [round(stp["stage_ms"][i] / max(stp["stage_launches"][i], 1), 4) for i in range(5)]
        allr = [None] * b.N
        b.dist.all_gather_object(allr, (b.sol.n3, mine))
        roofline["per_rank[planes, stage_ms x5]"] = allr
    tr = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu capture
    if os.path.exists(tr):
        try:
            with open(tr) as f:
                tj = json.load(f)
            key = f"{spec.name}:stage{dom['stage']}"
            if key in tj:      # captured on one GPU holding the whole grid; a slab moves its share
                roofline["traffic"] = tj[key] * b.sol.cells / spec.cells
        except Exception:
            pass
    return roofline


def run_workload(args, name, rank, N, local, lib_path, dry, dist, torch, steps, warmup, repeats, fixed_step, full):
    """returns (line fragment, Bench).  full: also stage profile, end to end, parity."""
    spec, scaling, geo = make_spec(name, N)
    b = Bench(args, spec, rank, N, local, lib_path, dry, dist, torch)
    sol = b.sol
    b.initial_state(sol)
    peak, peak_src = peaks()
    sb = STAGE_BYTES_HEAT if spec.calc_mode in (0, 1, 2) else STAGE_BYTES_NOHEAT

    # ---- warm-up, then `repeats` regions of EXACTLY `steps` attempts, device-resident ---------------
    H = None
    if fixed_step:
        # SURVEY.md 8(d)-5 asks for half the first accepted adaptive step; from tau = 1 the first attempts blow the
        # error estimate up and the controller answers with steps of 1e-17 .. 1e-31 before it recovers, so H is
        # taken after the controller has settled: half the adaptive step reached after 40 attempts (same on every rank)
        sol.solve(spec.final_time, max_attempts=max(40, warmup))
        H = 0.5 * sol.sys.h
        b.advance(sol, 3, H)
    else:
        sol.solve(spec.final_time, max_attempts=warmup)
    st0 = sol.stats()
    ev0 = sol.sys.rhs_evals
    regions = b.timed(sol, steps, repeats, H)
    st1 = sol.stats()
    devs = sorted(r[0] for r in regions)
    dev_ms = statistics.median(devs)
    wall_ms = statistics.median(sorted(r[1] for r in regions))
    attempts_all = st1["attempts"] - st0["attempts"]
    rhs_all = sol.sys.rhs_evals - ev0
    rhs_per_region = rhs_all / max(repeats, 1)
    total_cells = b.allreduce(float(sol.cells), "sum")
    value = total_cells * rhs_per_region / (dev_ms * 1e-3) / 1e6
    stage_launches = [(a - c) / max(repeats, 1) for a, c in zip(st1["stage_launches"], st0["stage_launches"])]
    # real work only: launches that found the integration finished are not counted as bytes moved
    noop = (st1["noop_launches"] - st0["noop_launches"]) / max(repeats, 1)
    skipped1 = attempts_all / max(repeats, 1) - (rhs_per_region - 4 * attempts_all / max(repeats, 1))      # stage 1 launches that returned at once (K1 reused)
    eff = [stage_launches[0] - noop / 5 - skipped1] + [n_ - noop / 5 for n_ in stage_launches[1:]]
    step_gbs = total_cells / N * sum(bb * n_ for bb, n_ in zip(sb, eff)) / (dev_ms * 1e-3) / 1e9
    frag = {
        "workload": spec.name, "grid": [spec.n1, spec.n2, spec.n3], "calc_mode": spec.calc_mode, "scaling": scaling,
        "mode": f"fixed step H={H:.6g}" if fixed_step else "adaptive (delta 1e-3)",
        "value": value, "unit": "Mcell-RHS/s", "steps": steps, "repeats": repeats, "ms_per_step": dev_ms / steps,
        "device_ms_regions": [round(d, 4) for d in devs], "wall_ms_per_step": wall_ms / steps,
        "rhs_evals_per_attempt": rhs_per_region / steps,
        "rasterise_ms": {"ms": round(getattr(b, "rasterise_ms", 0.0), 3), "spheres": int(len(spec.spheres)), "cells_per_gpu": int(sol.cells)},
        "slab_balance": b.balance,
        "roofline_step": {"achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak,
                          "note": "algorithmic bytes of the stage kernels that did work in a timed region / its device time, per GPU (launch gaps included)"},
    }
    info = {"spec": spec, "scaling": scaling, "geo": geo, "H": H, "st0": st0, "st1": st1, "regions": regions, "peak": peak, "peak_src": peak_src,
            "launches": (st1["kernel_launches"] - st0["kernel_launches"]) / max(repeats, 1), "total_cells": total_cells}
    return frag, b, info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--repeats", type=int, default=7, help="timed regions of --steps attempts each; the median is reported")
    ap.add_argument("--workload", default=os.environ.get("PFT_BENCH_WORKLOAD", "400-SigmaP1-P-smallsigma"))
    ap.add_argument("--also", default=os.environ.get("PFT_BENCH_ALSO", "default"),
                    help="comma-separated secondary workloads reported under `also` ('none' to skip); default: the 100 case (one GPU only), the PhysRevE-2025 200 case and "
                         "the SURVEY 8(d)-5 weak-scaling slab 800x800x(100 N) in fixed-step mode")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-balance", action="store_true", help="N > 1: keep the reference's even slab partition (default: planes dealt out by measured cost)")
    ap.add_argument("--no-dropin", action="store_true", help="skip the end-to-end figure through RK_MPI_SA_solve (e2e_dropin)")
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
    if args.repeats < 1:
        args.repeats = 1

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    # the bench line must be the only thing on stdout (NCCL prints its version banner there)
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

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

    from porousfreezethaw_b200 import _lib as _pftlib
    lib_path = _pftlib.FAST_LIB_PATH if args.variant == "fast" else None
    if os.environ.get("PFT_BENCH_LIB"):       # A/B builds (development)
        lib_path = os.environ["PFT_BENCH_LIB"]
    if dry:
        lib_path = dry

    # clocks / throttle reasons are sampled over the whole measurement (started BEFORE any barrier that
    # precedes a timed region, stopped after the last one), on rank 0's GPU
    sampler = ClockSampler(local) if (rank == 0 and not dry) else None

    frag, b, info = run_workload(args, args.workload, rank, N, local, lib_path, dry, dist, torch,
                                 args.steps, args.warmup, args.repeats, args.fixed_step, True)
    spec, sol, H = info["spec"], b.sol, info["H"]
    st0, st1 = info["st0"], info["st1"]
    attempts = args.steps
    # host side of the control loop, per attempt: time asleep waiting for a free slot of the launch queue / time issuing launches
    tot_att = max(st1["attempts"] - st0["attempts"], 1)
    host_wait = (st1["host_wait_ms"] - st0["host_wait_ms"]) / tot_att
    host_launch = (st1["host_launch_ms"] - st0["host_launch_ms"]) / tot_att
    my_dev = statistics.median(sorted(r[2] for r in info["regions"])) / attempts
    if N > 1:
        hw = [None] * N
        dist.all_gather_object(hw, (round(host_wait, 4), round(host_launch, 4), round(my_dev, 4)))
    else:
        hw = [(round(host_wait, 4), round(host_launch, 4), round(my_dev, 4))]

    clocks = sampler.stop() if sampler else None            # the headline regions are over

    roofline = stage_profile(b, H, info["peak"], info["peak_src"])

    # ---- end to end: pinned HOST buffers -> upload -> K attempts -> download ---------------
    # what RK_MPI_SA_solve does per call with a host-resident x: the temperature and phase fields travel both
    # ways, the static glass field is resident (csrc/rk_mpi_sa_b200.c)
    from porousfreezethaw_b200.solver import FIELD_U, FIELD_P
    xh = np.empty(3 * sol.SIZE) if dry else torch.empty(3 * sol.SIZE, dtype=torch.float64, pin_memory=True).numpy()
    sol.download(xh)                                         # a valid padded host image of the current state
    e2e_runs = []
    for k_e2e in (attempts, 10 * attempts):
        b.barrier()
        ev0 = sol.sys.rhs_evals
        a0 = sol.sys.steps_total
        t0 = time.perf_counter()
        sol.upload_fields(xh, FIELD_U | FIELD_P)
        b.advance(sol, k_e2e, H)
        sol.download_fields(xh, FIELD_U | FIELD_P)
        b.barrier()
        e2e_s = b.allreduce(time.perf_counter() - t0, "max")
        e2e_runs.append((sol.sys.steps_total - a0, info["total_cells"] * (sol.sys.rhs_evals - ev0) / e2e_s / 1e6, e2e_s))
    nbytes = 2 * sol.n3 * sol.N1 * sol.N2 * 8               # u and p, the slab's planes incl. their x/y ghost nodes
    e2e = {"value": e2e_runs[0][1], "unit": "Mcell-RHS/s", "h2d_bytes_per_step": nbytes / max(e2e_runs[0][0], 1),
           "d2h_bytes_per_step": nbytes / max(e2e_runs[0][0], 1), "attempts": e2e_runs[0][0],
           "long_call": {"attempts": e2e_runs[1][0], "value": e2e_runs[1][1], "seconds": e2e_runs[1][2]},
           "note": "one solve call = upload(u, p) + K attempts + download(u, p) through the C ABI with pinned host buffers, as RK_MPI_SA_solve "
                   "with a host-resident x (gl is static and stays on the device); long_call = the same with 10 K attempts"}

    # ---- the same through the reference's solver-module ABI (one host process; all GPUs it can see) ----
    e2e_dropin = None
    if N == 1 and not dry and not args.fixed_step and args.variant == "strict" and os.path.exists(_pftlib.RK_LIB_PATH) and not args.no_dropin:
        e2e_dropin = []
        t_probe = float(sol.sys.t)
        counts = [1] + ([torch.cuda.device_count()] if torch.cuda.device_count() > 1 else [])
        for g_ in counts:
            try:
                e2e_dropin.append(dropin_e2e(spec, g_, t_probe))
            except Exception as e:
                e2e_dropin.append({"gpus": g_, "error": repr(e)})

    # ---- parity of THIS workload against the unmodified reference -------------------------
    parity = None
    gpath = os.path.join(ROOT, "tests", "golden", f"partial_{spec.name}.npz")
    if not args.no_parity and not args.fixed_step and args.variant == "strict" and os.path.exists(gpath) and "plane_sum_u" in np.load(gpath) \
            and os.path.exists(_pftlib.RK_LIB_PATH) and not dry:
        try:
            parity = parity_check(b, gpath)
        except Exception as e:
            parity = {"ok": False, "error": repr(e)}

    # ---- secondary workloads -------------------------------------------------------------
    also = []
    # default: the PhysRevE-2025 "200" case and the weak-scaling slab at every N; the 250 k-cell "100" case (the reference's
    # CPU-runnable configuration) on one GPU only -- there is nothing to learn from 12 planes per GPU
    names = [] if args.also in ("none", "") else ((["100-GradP"] if N == 1 else []) + ["2025-200-GradP", "cube-800w"] if args.also == "default" else args.also.split(","))
    if dry:
        names = []
    for nm in names:
        try:
            fixed = nm.endswith("w")
            f2, b2, i2 = run_workload(args, nm, rank, N, local, lib_path, dry, dist, torch,
                                      30 if fixed else args.steps, 10 if fixed else args.warmup, 3, fixed, False)
            b2.sol.close()
            also.append(f2)
        except Exception as e:
            also.append({"workload": nm, "error": repr(e)})

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
            "metric": "Mcell-RHS evals/s", "value": frag["value"], "unit": "Mcell-RHS/s", "n_gpus": N,
            "steps": attempts, "warmup": args.warmup, "ms_per_step": frag["ms_per_step"],
            "higher_is_better": True, "scaling": info["scaling"], "vs_baseline": None, "dtype": "f64",
            "data": "DRY RUN on the CPU emulation (tests only): not a measurement" if dry else "synthetic",
            "config": {"workload": spec.name, "grid": [spec.n1, spec.n2, spec.n3], "calc_mode": spec.calc_mode,
                       "geometry": info["geo"], "mode": frag["mode"], "variant": args.variant,
                       "cache": ("inputs larger than L2 (no flush needed)" if sol.cells * 13 * 8 > 2 * 126e6 else
                                 "working set %d MB per GPU against 126 MB of L2, no flush: an attempt re-reads what the "
                                 "previous one wrote, as in production" % (sol.cells * 13 * 8 // 1000000)),
                       "parallelism": f"z-slabs x{N}", "tile": [st["tile_x"], st["tile_y"]], "z_chunk": st["z_chunk"],
                       "rhs_evals_per_attempt": frag["rhs_evals_per_attempt"],
                       "timing": f"median of {args.repeats} regions of {attempts} attempts, each: barrier + synchronize, CUDA events on the solver stream, max over ranks",
                       "controller": "device (twin of the host libm pow), attempts enqueued ahead: %d" % st["runahead"] if st["device_pow"] else "host pow per attempt",
                       "launch_order": "S2 S3 S4 S5 S1' (K1 of the next attempt evaluated speculatively from stage 5's candidate)"
                                       if st.get("speculative_stage1") else "S1 S2 S3 S4 S5"},
            "e2e": e2e, "gpu_launches": info["launches"], "clocks": clocks, "roofline": roofline,
            "device_ms_regions": frag["device_ms_regions"],
            "wall_ms_per_step": frag["wall_ms_per_step"],
            "host_loop_ms_per_step": {"per_rank[asleep, launching, device]": hw},
            "roofline_step": frag["roofline_step"],
            "e2e_dropin": e2e_dropin,
            "rasterise_ms": frag["rasterise_ms"],
            "slab_balance": frag["slab_balance"],
            "parity": parity,
            "also": also,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    sol.close()
    if N > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
