"""Snapshot / restart datasets.

The reference writes one NetCDF dataset per snapshot (apps/intertrack-hybrid-S-freezing/
intertrack.c:2331-2548): dims n3,n2,n1; coordinate variables n3,n2,n1; u,p,gl as
double[n3][n2][n1]; global attributes L1..L3, every model parameter, calc_mode, delta, tau, t,
final_time, snapshot, total_snapshots, title -- and restarts from one (`set icond_file`,
`set continue_series`, :1642-1669).  libnetcdf is not available in this image, so the reference
driver built here writes the same content through its netcdf.h stand-in into the "PFTSNAP 1"
container (text header + raw little-endian float64 payload, documented in oracle/shim/shim_rt.c).
This module reads and writes that container, reads and writes real NetCDF classic files of the
same layout (second half of the file), converts between the two, and provides the
post-processing quantity of scripts/avg.sh:9-20 (ice volume fraction = mean(p > 0.5)).
"""
from __future__ import annotations

import numpy as np

MAGIC = b"PFTSNAP 1\n"


def read_snapshot(path: str) -> dict:
    """-> {"dims": {name: len}, "vars": {name: ndarray}, "attrs": {name: value}}"""
    with open(path, "rb") as f:
        if f.readline() != MAGIC:
            raise ValueError(f"{path}: not a PFTSNAP 1 dataset")
        dims, dim_order, var_defs, attrs = {}, [], [], {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated header")
            tok = line.decode("utf-8", "replace").split()
            if not tok:
                continue
            if tok[0] == "end":
                break
            if tok[0] == "dim":
                dims[tok[1]] = int(tok[2]); dim_order.append(tok[1])
            elif tok[0] == "var":
                nd = int(tok[2])
                var_defs.append((tok[1], [dim_order[int(i)] for i in tok[3:3 + nd]]))
            elif tok[0] == "attd":
                attrs[tok[1]] = float(tok[2])
            elif tok[0] == "atti":
                attrs[tok[1]] = int(tok[2])
            elif tok[0] == "attt":
                n = int(tok[2])
                attrs[tok[1]] = f.read(n).decode("utf-8", "replace")
                f.read(1)
        out = {}
        for name, dnames in var_defs:
            shape = tuple(dims[d] for d in dnames)
            cnt = int(np.prod(shape)) if shape else 1
            out[name] = np.frombuffer(f.read(8 * cnt), dtype="<f8").reshape(shape).copy()
    return {"dims": dims, "vars": out, "attrs": attrs}


def write_snapshot(path: str, u, p, gl, L, attrs: dict, title: str = "") -> None:
    """Writes a dataset with the reference's layout (intertrack.c:2331-2447): coordinates are the
    cell centres L_d*(i+0.5)/n_d."""
    u, p, gl = (np.ascontiguousarray(a, dtype="<f8") for a in (u, p, gl))
    n3, n2, n1 = u.shape
    with open(path, "wb") as f:
        f.write(MAGIC)
        for name, n in (("n3", n3), ("n2", n2), ("n1", n1)):
            f.write(f"dim {name} {n}\n".encode())
        f.write(b"var n3 1 0\nvar n2 1 1\nvar n1 1 2\n")
        for name in ("u", "p", "gl"):
            f.write(f"var {name} 3 0 1 2\n".encode())
        for k, v in zip(("L1", "L2", "L3"), L):
            f.write(f"attd {k} {float(v):.17g}\n".encode())
        for k, v in attrs.items():
            if isinstance(v, (int, np.integer)):
                f.write(f"atti {k} {int(v)}\n".encode())
            else:
                f.write(f"attd {k} {float(v):.17g}\n".encode())
        t = title.encode()
        f.write(f"attt title {len(t)}\n".encode() + t + b"\n")
        f.write(b"end\n")
        for n, Ld in ((n3, L[2]), (n2, L[1]), (n1, L[0])):
            f.write((Ld * (np.arange(n) + 0.5) / n).astype("<f8").tobytes())
        for a in (u, p, gl):
            f.write(a.tobytes())


def ice_fraction(p) -> float:
    """scripts/avg.sh:9-20: the mean over all nodes of (p > 0.5)"""
    return float((np.asarray(p) > 0.5).mean())


# ---------------------------------------------------------------------------------------------
# NetCDF classic: the format the reference itself writes (nc_create(..., NC_CLOBBER, ...),
# intertrack.c:2331) and restarts from (:1589, :2040).  Self-contained writer / reader of the
# classic (CDF-1) and 64-bit-offset (CDF-2) file format -- fixed-size variables, global attributes
# of type double / int / char, which is all the reference's datasets contain -- so that snapshots
# produced here open in ncdump / ncview / scripts/avg.sh-style tools, and datasets written by an
# MPI+netCDF build of the reference can seed a run here.

import struct

_NC_DIMENSION, _NC_VARIABLE, _NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C
_NC_CHAR, _NC_INT, _NC_DOUBLE = 2, 4, 6
_NC_TYPES = {1: ("b", 1), 2: ("c", 1), 3: (">i2", 2), 4: (">i4", 4), 5: (">f4", 4), 6: (">f8", 8)}


def _pad4(n: int) -> int:
    return (n + 3) & ~3


def _nc_name(s: str) -> bytes:
    b = s.encode()
    return struct.pack(">i", len(b)) + b + b"\0" * (_pad4(len(b)) - len(b))


def _nc_attr(name: str, v) -> bytes:
    if isinstance(v, str):
        b = v.encode()
        return _nc_name(name) + struct.pack(">ii", _NC_CHAR, len(b)) + b + b"\0" * (_pad4(len(b)) - len(b))
    if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
        return _nc_name(name) + struct.pack(">iii", _NC_INT, 1, int(v))
    return _nc_name(name) + struct.pack(">ii", _NC_DOUBLE, 1) + struct.pack(">d", float(v))


def write_netcdf(path: str, snap: dict) -> None:
    """Writes {"dims", "vars", "attrs"} (as read_snapshot returns it) as a NetCDF classic dataset;
    dimension order and attribute order are kept, every variable is double.  Variables whose
    dimensions are not recorded in snap["var_dims"] are matched to the dimensions by shape, with
    the reference's convention first: a 1-D variable named like a dimension is its coordinate."""
    dims = snap["dims"]
    dnames = list(dims)
    given = snap.get("var_dims", {})
    var_dims = {}
    for name, a in snap["vars"].items():
        if name in given:
            var_dims[name] = list(given[name])
            continue
        if a.ndim == 1 and name in dims and dims[name] == a.shape[0]:
            var_dims[name] = [name]
        else:
            if tuple(dims[d] for d in dnames[-a.ndim:]) != a.shape and tuple(dims[d] for d in dnames[:a.ndim]) != a.shape:
                raise ValueError(f"cannot match variable {name} of shape {a.shape} to the dimensions {dims}")
            var_dims[name] = dnames[:a.ndim] if tuple(dims[d] for d in dnames[:a.ndim]) == a.shape else dnames[-a.ndim:]
    sizes = {n: _pad4(8 * int(np.prod(a.shape))) for n, a in snap["vars"].items()}
    hdr, _ = _nc_header(dims, snap["attrs"], var_dims, sizes)
    with open(path, "wb") as f:
        f.write(hdr)
        for name, a in snap["vars"].items():
            b = np.ascontiguousarray(a, dtype=">f8").tobytes()
            f.write(b + b"\0" * (sizes[name] - len(b)))


def _nc_header(dims: dict, attrs: dict, var_dims: dict, sizes: dict):
    """header bytes and format version for double variables `var_dims` (name -> dimension names) of
    `sizes` bytes each, laid out back to back after the header: classic (CDF-1) while every offset
    fits 31 bits, else 64-bit offsets (CDF-2)"""
    dnames = list(dims)

    def header(version: int, begins: dict) -> bytes:
        h = b"CDF" + bytes([version]) + struct.pack(">i", 0)
        h += struct.pack(">ii", _NC_DIMENSION, len(dnames)) if dnames else struct.pack(">ii", 0, 0)
        for d in dnames:
            h += _nc_name(d) + struct.pack(">i", dims[d])
        h += struct.pack(">ii", _NC_ATTRIBUTE, len(attrs)) if attrs else struct.pack(">ii", 0, 0)
        for k, v in attrs.items():
            h += _nc_attr(k, v)
        h += struct.pack(">ii", _NC_VARIABLE, len(var_dims)) if var_dims else struct.pack(">ii", 0, 0)
        for name, dn in var_dims.items():
            ids = [dnames.index(d) for d in dn]
            h += _nc_name(name) + struct.pack(">i", len(ids)) + b"".join(struct.pack(">i", i) for i in ids)
            h += struct.pack(">ii", 0, 0)                                    # no per-variable attributes
            h += struct.pack(">i", _NC_DOUBLE) + struct.pack(">I", min(sizes[name], 0xFFFFFFFF))
            h += struct.pack(">i" if version == 1 else ">q", begins.get(name, 0))
        return h

    for version in (1, 2):
        begins, off = {}, len(header(version, {}))
        for name in var_dims:
            begins[name] = off
            off += sizes[name]
        if version == 1 and off > 0x7FFFFFFF:
            continue                                                         # the classic format stops at 2 GiB
        return header(version, begins), version
    raise AssertionError("unreachable")


def read_netcdf(path: str) -> dict:
    """Reads a NetCDF classic / 64-bit-offset dataset without record variables into the same
    {"dims", "vars", "attrs", "var_dims"} form as read_snapshot."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:3] != b"CDF" or buf[3] not in (1, 2):
        raise ValueError(f"{path}: not a NetCDF classic dataset")
    version, pos = buf[3], 8

    def i4():
        nonlocal pos
        v = struct.unpack_from(">i", buf, pos)[0]
        pos += 4
        return v

    def name():
        nonlocal pos
        n = i4()
        s = buf[pos:pos + n].decode()
        pos += _pad4(n)
        return s

    def attrs():
        nonlocal pos
        tag, n = i4(), i4()
        out = {}
        if tag == 0:
            return out
        if tag != _NC_ATTRIBUTE:
            raise ValueError(f"{path}: malformed attribute list")
        for _ in range(n):
            k = name()
            t, cnt = i4(), i4()
            dt, sz = _NC_TYPES[t]
            raw = buf[pos:pos + cnt * sz]
            pos += _pad4(cnt * sz)
            if t == _NC_CHAR:
                out[k] = raw.decode("utf-8", "replace")
            else:
                v = np.frombuffer(raw, dtype=dt)
                v = v.astype(np.int64 if t in (1, 3, 4) else np.float64)
                out[k] = (int(v[0]) if t in (1, 3, 4) else float(v[0])) if cnt == 1 else v
        return out

    dims, dnames = {}, []
    tag, n = i4(), i4()
    if tag == _NC_DIMENSION:
        for _ in range(n):
            d = name()
            dims[d] = i4()
            dnames.append(d)
    gatt = attrs()
    out, var_dims = {}, {}
    tag, n = i4(), i4()
    if tag == _NC_VARIABLE:
        for _ in range(n):
            v = name()
            nd = i4()
            ids = [i4() for _ in range(nd)]
            attrs()
            t = i4()
            pos += 4                                                         # vsize
            begin = struct.unpack_from(">i" if version == 1 else ">q", buf, pos)[0]
            pos += 4 if version == 1 else 8
            dt, sz = _NC_TYPES[t]
            shape = tuple(dims[dnames[i]] for i in ids)
            cnt = int(np.prod(shape)) if shape else 1
            out[v] = np.frombuffer(buf, dtype=dt, count=cnt, offset=begin).astype(np.float64).reshape(shape)
            var_dims[v] = [dnames[i] for i in ids]
    return {"dims": dims, "vars": out, "attrs": gatt, "var_dims": var_dims}


def convert(src: str, dst: str) -> None:
    """PFTSNAP 1 -> NetCDF classic, or NetCDF classic -> PFTSNAP 1 (by the source's magic)."""
    with open(src, "rb") as f:
        magic = f.read(10)
    if magic == MAGIC:
        write_netcdf(dst, read_snapshot(src))
        return
    s = read_netcdf(src)
    L = [s["attrs"].get(k, 1.0) for k in ("L1", "L2", "L3")]
    attrs = {k: v for k, v in s["attrs"].items() if k not in ("L1", "L2", "L3", "title") and not isinstance(v, str)}
    write_snapshot(dst, s["vars"]["u"], s["vars"]["p"], s["vars"]["gl"], L, attrs, title=str(s["attrs"].get("title", "")))
