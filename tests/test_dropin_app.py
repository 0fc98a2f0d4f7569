"""The drop-in proof: the reference's UNMODIFIED driver (intertrack.c, model.c, the Params
parsers) linked against the B200 solver module (rk_mpi_sa_b200.c) and model plug-in
(equation_plugin_b200.c) instead of RK_MPI_SAsolver_hybrid2.c / equation.c, run on a Params
case file next to the reference application itself: the snapshot datasets must be byte-identical
and the logged step counts equal.

 * on the GPU box (-m gpu): oracle/_ref/intertrack_b200 (prebuilt here, linked to the nvcc library)
 * in the build container: the same objects linked to the CPU emulation of the CUDA library
"""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

PARAMS = """\
icond u = "293.15"
icond p = "z>0.052 and z<0.058 and ((x-L1/2)^2+(y-L2/2)^2 < (L1/3)^2)"
icond gl = "0.5*(1.0 + tanh(0.5/xi_gl*(z-0.055))"
set logfile = $OUTPUT/intertrack.log
set out_file = $OUTPUT/image out_file_suffix = .ncd
{debug}
hours		60*60
L1		0.03
L2		0.03
L3		0.06
u_noise_amp	0
water_cp	4.18e3
ice_cp		2.05e3
glass_cp	0.84e3
water_lambda	0.6
ice_lambda	2.22
glass_lambda	1.1
water_rho	997
ice_rho		917
glass_rho	2500
u_star		273.15
L		3.34e5
xi_gl           L3/300
zeta            2.0
beads_scaling   L1
ball_radius	0.1*beads_scaling
beads_offset_x  0
beads_offset_y  beads_offset_x
beads_offset_z  beads_offset_x
xi		L3/{xi_div}
a		2
b		1
alpha		water_rho*water_cp
mu		1e-4
p_eps0		0.05
p_eps1		0.2
gamma		2
top_temp1	        273.15 - 25
top_temp2	        273.15 + 20
phase_switch_time       5*hours
calc_mode	{mode}
final_time	{final_time}
saved_files	{saved}
delta		1e-3
tau_min		1e-6
tau		1
grid_nodes	{grid}
multiplier	grid_nodes / (L1 max L2 max L3)
n1		L1 * multiplier
n2		L2 * multiplier
n3		L3 * multiplier
set comment="drop-in test"
"""


def _run(binary, workdir, outdir, threads=8, ngpus=None):
    os.makedirs(os.path.join(workdir, outdir), exist_ok=True)
    env = dict(os.environ, OUTPUT=outdir, OMP_NUM_THREADS=str(threads), OMP_SCHEDULE="static")
    env.pop("PFT_NGPUS", None)
    if ngpus:
        env["PFT_NGPUS"] = str(ngpus)      # ONE host process cuts the block into z-slabs, one per GPU
    r = subprocess.run([binary, "Params"], cwd=workdir, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    return r.stdout


def _compare(workdir, a, b, saved):
    for k in range(saved):
        fa, fb = (os.path.join(workdir, d, f"image.{k:03d}.ncd") for d in (a, b))
        assert open(fa, "rb").read() == open(fb, "rb").read(), f"snapshot {k} differs"
    pat = re.compile(r"(\d+) R-K steps \((\d+) total\)")
    la, lb = (pat.findall(open(os.path.join(workdir, d, "intertrack.log")).read()) for d in (a, b))
    assert la == lb and len(la) == saved


def _case(tmp_path, mode, grid, final_time, saved, debug=False, full_grid=False):
    (tmp_path / "data").mkdir()
    shutil.copy(os.path.join(REF, "data", "spheres_positions.txt"), tmp_path / "data" / "spheres_positions.txt")
    dbg = "set debug_logfile = $OUTPUT/RK.log" if debug else ""
    if full_grid:
        dbg += "\ngrid full"       # intertrack.c:971-973: export the padded grid incl. ghost shell -> bcond_setup (:2511)
    (tmp_path / "Params").write_text(PARAMS.format(mode=mode, grid=grid, final_time=final_time, saved=saved, debug=dbg, xi_div=grid))


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "intertrack_ref")), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "obj")), reason="reference objects only exist in the build container")
@pytest.mark.parametrize("mode,debug,ngpus,full_grid", [(0, False, None, False), (1, True, None, False), (0, False, 2, False), (1, True, 3, False),
                                                        (0, False, None, True), (1, False, 2, True)])
def test_dropin_app_on_emulation(tmp_path, emu_lib, mode, debug, ngpus, full_grid):
    """debug=True installs the reference's Service_Callback (RKService: per-step log), ngpus drives several
    emulated devices from the one process (worker threads, peer-memory protocol), full_grid makes the driver
    export the ghost shell too (bcond_setup of the plug-in -> pft_equation_bcond_host)"""
    objs = [os.path.join(REF, "obj", o) for o in ("intertrack_b200.o", "rk_mpi_sa_b200.o", "equation_b200.o")]
    objs += [o for o in glob.glob(os.path.join(REF, "obj", "*.o"))
             if os.path.basename(o).split(".")[0] in ("pparser", "cparser", "evsubst", "mprintf", "ee_wrapper", "exp_all", "_AVS", "shim_rt")
             or os.path.basename(o).startswith("str_")]
    exe = str(tmp_path / "intertrack_b200_emu")
    subprocess.run(["gcc", "-fopenmp", "-o", exe, *objs, emu_lib, "-Wl,-rpath," + os.path.dirname(emu_lib), "-lstdc++", "-lm"], check=True)
    _case(tmp_path, mode, 24, 0.6, 3, debug, full_grid)
    _run(os.path.join(REF, "intertrack_ref"), str(tmp_path), "OUT_ref")
    _run(exe, str(tmp_path), "OUT_b200", ngpus=ngpus)
    _compare(str(tmp_path), "OUT_ref", "OUT_b200", 3)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("mode,grid,final_time,debug", [(0, 100, 3.0, False), (1, 100, 3.0, True), (11, 60, 5.0, False)])
def test_dropin_app_on_gpu(tmp_path, mode, grid, final_time, debug):
    """e.g. grid 100 = the "100" low-resolution case (50x50x100): reference CPU app vs the same
    driver on the B200 drop-in, first seconds of the run, 4 snapshots"""
    exe = os.path.join(REF, "intertrack_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/intertrack_b200 not built")
    _case(tmp_path, mode, grid, final_time, 4, debug)
    _run(os.path.join(REF, "intertrack_ref"), str(tmp_path), "OUT_ref", threads=os.cpu_count() or 8)
    _run(exe, str(tmp_path), "OUT_b200")
    _compare(str(tmp_path), "OUT_ref", "OUT_b200", 4)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("ngpus,mode,grid,final_time,debug", [(2, 1, 100, 3.0, False), (2, 0, 100, 2.0, True), (4, 1, 100, 3.0, False), (8, 1, 200, 0.5, False)])
def test_dropin_app_multi_gpu(tmp_path, ngpus, mode, grid, final_time, debug):
    """the UNMODIFIED reference driver on 2/4/8 B200s of one box through RK_MPI_SA_solve: PFT_NGPUS=n makes the
    one host process cut its block into n z-slabs (intertrack.c:1780-1787) -- snapshots byte-identical to the
    reference application's, logged step counts equal.  Skipped on boxes with fewer GPUs."""
    import torch
    if torch.cuda.device_count() < ngpus:
        pytest.skip(f"needs {ngpus} GPUs")
    exe = os.path.join(REF, "intertrack_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/intertrack_b200 not built")
    _case(tmp_path, mode, grid, final_time, 4, debug)
    _run(os.path.join(REF, "intertrack_ref"), str(tmp_path), "OUT_ref", threads=os.cpu_count() or 8)
    _run(exe, str(tmp_path), "OUT_b200", ngpus=ngpus)
    _compare(str(tmp_path), "OUT_ref", "OUT_b200", 4)
