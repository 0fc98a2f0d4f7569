"""Snapshot / restart datasets.

The reference writes one NetCDF dataset per snapshot (apps/intertrack-hybrid-S-freezing/
intertrack.c:2331-2548): dims n3,n2,n1; coordinate variables n3,n2,n1; u,p,gl as
double[n3][n2][n1]; global attributes L1..L3, every model parameter, calc_mode, delta, tau, t,
final_time, snapshot, total_snapshots, title -- and restarts from one (`set icond_file`,
`set continue_series`, :1642-1669).  libnetcdf is not available in this image, so the same
content travels in the "PFTSNAP 1" container (text header + raw little-endian float64 payload,
documented in oracle/shim/shim_rt.c, which is also what the reference driver writes here through
its netcdf.h stand-in).  This module reads and writes that container and provides the
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
