"""NetCDF classic output / restart input: the dataset layout the reference writes
(apps/intertrack-hybrid-S-freezing/intertrack.c:2331-2447) in the file format it writes it in
(nc_create(..., NC_CLOBBER, ...)).  scipy.io.netcdf_file is the independent reader/writer."""
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as po
from porousfreezethaw_b200 import snapshot as sn

scipy_io = pytest.importorskip("scipy.io")
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")


def _fields(shape=(7, 5, 6), seed=3):
    rng = np.random.default_rng(seed)
    return tuple(rng.random(shape) for _ in range(3))


def test_written_dataset_opens_in_an_independent_reader(tmp_path):
    u, p, gl = _fields()
    attrs = {"xi": 2.5e-4, "alpha": 3.0, "calc_mode": 1, "delta": 1e-3, "tau": 1.0, "t": 363.6363636363636,
             "final_time": 36000.0, "snapshot": 1, "total_snapshots": 100}
    sn.write_snapshot(str(tmp_path / "image.001.ncd"), u, p, gl, (0.03, 0.03, 0.06), attrs, title="freeze-thaw")
    sn.convert(str(tmp_path / "image.001.ncd"), str(tmp_path / "image.001.nc"))
    f = scipy_io.netcdf_file(str(tmp_path / "image.001.nc"), "r", mmap=False)
    assert f.version_byte == 1                                     # classic, as NC_CLOBBER gives
    assert list(f.dimensions.items()) == [("n3", 7), ("n2", 5), ("n1", 6)]          # intertrack.c:2340-2346
    assert list(f.variables) == ["n3", "n2", "n1", "u", "p", "gl"]                  # :2350-2354
    for name, a in (("u", u), ("p", p), ("gl", gl)):
        v = f.variables[name]
        assert v.dimensions == ("n3", "n2", "n1") and v.typecode() == "d"
        assert np.array_equal(v[:], a)
    assert np.array_equal(f.variables["n1"][:], 0.03 * (np.arange(6) + 0.5) / 6)    # cell centres, :2420-2440
    assert f.L3 == 0.06 and f.xi == 2.5e-4 and f.t == attrs["t"]
    assert int(f.calc_mode) == 1 and int(f.total_snapshots) == 100 and f.calc_mode.dtype.kind == "i"
    assert f.title == b"freeze-thaw"
    f.close()


def test_round_trips(tmp_path):
    u, p, gl = _fields((4, 3, 9), seed=5)
    sn.write_snapshot(str(tmp_path / "a.ncd"), u, p, gl, (1.0, 2.0, 3.0), {"t": 0.1 + 0.2, "snapshot": 4}, title="")
    sn.convert(str(tmp_path / "a.ncd"), str(tmp_path / "a.nc"))
    sn.convert(str(tmp_path / "a.nc"), str(tmp_path / "b.ncd"))
    a, b = sn.read_snapshot(str(tmp_path / "a.ncd")), sn.read_snapshot(str(tmp_path / "b.ncd"))
    assert a["dims"] == b["dims"] and a["attrs"] == b["attrs"]
    assert all(np.array_equal(a["vars"][k], b["vars"][k]) for k in a["vars"])
    assert open(tmp_path / "a.ncd", "rb").read() == open(tmp_path / "b.ncd", "rb").read()


def test_reads_datasets_written_by_another_netcdf_implementation(tmp_path):
    """what an MPI+netCDF build of the reference would hand us as `icond_file` (intertrack.c:1589): classic
    and 64-bit-offset files, attributes of every numeric type"""
    u, p, gl = _fields((3, 4, 5), seed=9)
    for version in (1, 2):
        path = str(tmp_path / f"ic{version}.nc")
        g = scipy_io.netcdf_file(path, "w", version=version)
        for d, n in (("n3", 3), ("n2", 4), ("n1", 5)):
            g.createDimension(d, n)
        for name, a in (("u", u), ("p", p), ("gl", gl)):
            g.createVariable(name, "d", ("n3", "n2", "n1"))[:] = a
        g.L1, g.L2, g.L3, g.t = 0.03, 0.03, 0.06, 12.5
        g.calc_mode = np.int32(10)
        g.title = "restart"
        g.close()
        s = sn.read_netcdf(path)
        assert s["dims"] == {"n3": 3, "n2": 4, "n1": 5}
        assert all(np.array_equal(s["vars"][k], a) for k, a in (("u", u), ("p", p), ("gl", gl)))
        assert s["attrs"]["t"] == 12.5 and s["attrs"]["calc_mode"] == 10 and s["attrs"]["title"] == "restart"
        assert sn.ice_fraction(s["vars"]["p"]) == float((p > 0.5).mean())


def test_large_dataset_switches_to_64bit_offsets():
    """the classic format stops at 2 GiB (the 800x800x100 block of one GPU is 1.5 GiB of fields);
    beyond that the writer emits the 64-bit-offset variant"""
    dims = {"n3": 100, "n2": 800, "n1": 800}
    vd = {k: ["n3", "n2", "n1"] for k in ("u", "p", "gl")}
    one = 8 * 100 * 800 * 800
    hdr, version = sn._nc_header(dims, {"t": 0.0}, vd, {k: one for k in vd})
    assert version == 1 and hdr[:4] == b"CDF\x01"
    dims["n3"] = 200
    hdr, version = sn._nc_header(dims, {"t": 0.0}, vd, {k: 2 * one for k in vd})
    assert version == 2 and hdr[:4] == b"CDF\x02"
    assert int.from_bytes(hdr[-8:], "big") == len(hdr) + 4 * one          # begin of the last variable, 64 bits


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_application_snapshot_converts(tmp_path):
    """a dataset written by the reference's own driver (through the netcdf.h stand-in) -> NetCDF classic"""
    text = open(os.path.join(REF, "cases", "Params.shipped")).read()
    text = text.replace("grid_nodes\t100", "grid_nodes\t16").replace("saved_files\t100", "saved_files\t2").replace("final_time\t10*hours", "final_time\t1e-6")
    text = "\n".join(l for l in text.splitlines() if not l.startswith("set debug_logfile") and not l.startswith("set snapshot_trigger"))
    (tmp_path / "Params").write_text(text)
    (tmp_path / "data").mkdir(); (tmp_path / "OUT").mkdir()
    open(tmp_path / "data" / "spheres_positions.txt", "w").close()
    r = subprocess.run([os.path.join(REF, "intertrack_ref"), "Params"], cwd=tmp_path, env=dict(os.environ, OUTPUT="OUT", OMP_NUM_THREADS="2"),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    src = str(tmp_path / "OUT" / "image.000.ncd")
    sn.convert(src, str(tmp_path / "image.000.nc"))
    a = sn.read_snapshot(src)
    f = scipy_io.netcdf_file(str(tmp_path / "image.000.nc"), "r", mmap=False)
    assert [(d, n) for d, n in f.dimensions.items()] == list(a["dims"].items())
    for k, v in a["vars"].items():
        assert np.array_equal(f.variables[k][:], v), k
    for k, v in a["attrs"].items():
        got = getattr(f, k)
        assert (got.decode() if isinstance(got, bytes) else got) == v, k
    f.close()


def test_restart_input_from_either_container(tmp_path):
    """`set icond_file` / `set continue_series` (intertrack.c:1585-1669, 2036-2110) in the Params front end"""
    from porousfreezethaw_b200 import cases, params
    spec = cases.small_case(6, 5, 9, calc_mode=1)
    u, p, gl = _fields((9, 5, 6), seed=21)
    attrs = {"t": 123.5, "tau": 0.25, "final_time": 36000.0, "snapshot": 17, "total_snapshots": 100, "calc_mode": 1}
    sn.write_snapshot(str(tmp_path / "image.017.ncd"), u, p, gl, (spec.L1, spec.L2, spec.L3), attrs, title="t")
    sn.convert(str(tmp_path / "image.017.ncd"), str(tmp_path / "image.017.nc"))
    for path in ("image.017.ncd", "image.017.nc"):
        r = params.restart_from_snapshot(spec, str(tmp_path / path))
        assert np.array_equal(r.u, u) and np.array_equal(r.p, p) and np.array_equal(r.gl, gl)
        assert (r.t, r.tau, r.snapshot) == (0.0, None, 0)                    # plain icond_file: fields only
        r = params.restart_from_snapshot(spec, str(tmp_path / path), continue_series=True, first_row=3, n3_local=4)
        assert np.array_equal(r.p, p[3:7]) and r.u.flags["C_CONTIGUOUS"]      # one rank's block (:2046-2062)
        assert (r.t, r.tau, r.final_time, r.snapshot, r.total_snapshots) == (123.5, 0.25, 36000.0, 17, 100)
    with pytest.raises(params.ExprError):
        params.restart_from_snapshot(cases.small_case(6, 5, 8, calc_mode=1), str(tmp_path / "image.017.nc"))
    sn.write_snapshot(str(tmp_path / "bare.ncd"), u, p, gl, (1, 1, 1), {"t": 1.0}, title="")
    with pytest.raises(params.ExprError):
        params.restart_from_snapshot(spec, str(tmp_path / "bare.ncd"), continue_series=True)
