"""`Params` case-file front end, independent of the reference's parsers.

The reference reads a case file line by line (modules/pparser/pparser.c:60-114): a line is either
a command for the application's handler (`set opt[=value] ...`, `icond VAR = "formula"`, `grid`,
... -- tables at apps/intertrack-hybrid-S-freezing/intertrack.c:925-998) or `name expression`,
evaluated by libsource/exprsion and defined as a variable.  Parameters are then fetched by name
(intertrack.c:1491-1575) and the initial-condition formulas are evaluated per node
(:1934-1968, multi-pass so that one quantity may refer to another).

Operator table and binding strength as in libsource/exprsion/exp_all.cc:405-445 and
ee_wrapper.cc:320-335 (smaller number binds tighter):
    ^ root 14 | unary - + functions max min sgn 16 | * / 20 | + - 22 | < > = 24 | not 25 |
    and or 26 | ? 27 | : 28
`max` / `min` are infix (`a max b`).  An expression that ends with open parentheses is accepted
(the archived cases contain `"0.5*(1.0 + tanh(0.5/xi_gl*(z-0.055))"`; exp_all.cc:866).

Expressions are evaluated with numpy, so the initial-condition formulas run over whole grids.
"""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass, field

import numpy as np

from .cases import PARAM_NAMES, CaseSpec

_FUNCS = {
    "int": np.trunc, "floor": np.floor, "ceil": np.ceil, "round": np.round, "abs": np.abs,
    "sin": np.sin, "cos": np.cos, "tan": np.tan, "asin": np.arcsin, "acos": np.arccos, "atan": np.arctan,
    "sinh": np.sinh, "cosh": np.cosh, "tanh": np.tanh, "asinh": np.arcsinh, "acosh": np.arccosh, "atanh": np.arctanh,
    "log": np.log10, "ln": np.log, "sqrt": np.sqrt, "exp": np.exp, "pow10": lambda x: 10.0 ** x,
    "sgn": np.sign, "not": lambda x: (np.asarray(x) == 0) * 1.0,
}
_BINARY = {
    "^": (14, lambda a, b: np.power(a, b)), "root": (14, lambda a, b: np.power(b, 1.0 / a)),
    "max": (16, np.maximum), "min": (16, np.minimum),
    "*": (20, np.multiply), "/": (20, np.divide),
    "+": (22, np.add), "-": (22, np.subtract),
    "<": (24, lambda a, b: (a < b) * 1.0), ">": (24, lambda a, b: (a > b) * 1.0), "=": (24, lambda a, b: (a == b) * 1.0),
    "and": (26, lambda a, b: ((np.asarray(a) != 0) & (np.asarray(b) != 0)) * 1.0),
    "or": (26, lambda a, b: ((np.asarray(a) != 0) | (np.asarray(b) != 0)) * 1.0),
}
_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)|([A-Za-z_][A-Za-z_0-9]*)|(.))")


class ExprError(ValueError):
    pass


def _tokens(s: str):
    out, pos = [], 0
    while pos < len(s):
        m = _TOKEN.match(s, pos)
        if not m:
            break
        pos = m.end()
        num, name, op = m.groups()
        if num is not None:
            out.append(("num", float(num)))
        elif name is not None:
            out.append(("name", name))
        elif op.strip():
            out.append(("op", op))
    return out


class _Parser:
    def __init__(self, toks, variables):
        self.t, self.i, self.v = toks, 0, variables

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("end", None)

    def next(self):
        tok = self.peek()
        self.i += 1
        return tok

    def expr(self, limit=100):
        """parse operators binding tighter than `limit` (smaller number = tighter)"""
        left = self.prefix()
        while True:
            kind, val = self.peek()
            if kind == "op" and val == "?" and 27 < limit:
                self.next()
                a = self.expr(28)
                k2, v2 = self.next()
                if not (k2 == "op" and v2 == ":"):
                    raise ExprError("'?' without ':'")
                b = self.expr(28)
                left = np.where(np.asarray(left) != 0, a, b)
                continue
            key = val if kind in ("op", "name") else None
            if key in _BINARY and _BINARY[key][0] < limit:
                prec, fn = _BINARY[key]
                self.next()
                # left-associative: the right operand takes only tighter operators ('^' included)
                right = self.expr(prec)
                left = fn(left, right)
                continue
            return left

    def prefix(self):
        kind, val = self.next()
        if kind == "num":
            return val
        if kind == "op" and val == "(":
            x = self.expr(100)
            k2, v2 = self.peek()
            if k2 == "op" and v2 == ")":
                self.next()
            elif k2 != "end":
                raise ExprError(f"expected ')' near token {self.i}")
            return x          # an unclosed parenthesis at the end of the formula is tolerated
        if kind == "op" and val in "+-":
            x = self.expr(16 + 0.5)      # unary sign: binds like a function (16), '^' (14) is tighter
            return -x if val == "-" else x
        if kind == "name":
            if val in _FUNCS:
                prec = 25.5 if val == "not" else 16.5
                return _FUNCS[val](self.expr(prec))
            if val == "pi":
                return math.pi
            if val == "e" and "e" not in self.v:
                return math.e
            if val in self.v:
                return self.v[val]
            raise ExprError(f"undefined variable '{val}'")
        raise ExprError(f"unexpected token {val!r}")


def evaluate(expr: str, variables: dict):
    """Evaluate one expression; variables may hold floats or numpy arrays."""
    p = _Parser(_tokens(expr), variables)
    x = p.expr(100)
    while p.peek() == ("op", ")"):       # stray closers are ignored like stray openers
        p.next()
    if p.peek()[0] != "end":
        raise ExprError(f"trailing input in {expr!r}")
    return x


def to_int(x: float) -> int:
    """intertrack.c:673-684: round to nearest"""
    r = math.floor(x)
    return int(r + 1) if x - r >= 0.5 else int(r)


@dataclass
class ParsedCase:
    variables: dict = field(default_factory=dict)        # `name expression` lines
    options: dict = field(default_factory=dict)          # `set opt[=value]`
    icond: dict = field(default_factory=dict)            # `icond VAR = "formula"`
    grid_mode: str = "inner"


def _subst_env(s: str) -> str:
    """modules/evsubst/evsubst.c: $VAR and ${VAR}"""
    return re.sub(r"\$\{?([A-Za-z_][A-Za-z_0-9]*)\}?", lambda m: os.environ.get(m.group(1), ""), s)


def parse_case_text(text: str) -> ParsedCase:
    pc = ParsedCase()
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line.startswith("#"):
            continue
        head = line.split(None, 1)
        cmd, rest = head[0], (head[1] if len(head) > 1 else "")
        if cmd == "icond":
            m = re.match(r'\s*([A-Za-z_]\w*)\s*=\s*"(.*)"\s*$', rest)
            if not m:
                raise ExprError(f"bad icond line: {raw!r}")
            pc.icond[m.group(1)] = m.group(2)
        elif cmd == "set":
            rest = rest.split("#", 1)[0]
            for m in re.finditer(r'([A-Za-z_]\w*)\s*(?:=\s*("[^"]*"|\S+))?', rest):
                val = m.group(2)
                pc.options[m.group(1)] = True if val is None else _subst_env(val.strip('"'))
        elif cmd == "grid":
            pc.grid_mode = rest.strip() or "inner"
        elif cmd in ("mnemonic", "continue_if", "break"):
            continue          # batch-mode commands: not part of a single case
        else:
            pc.variables[cmd] = float(evaluate(rest.split("#", 1)[0], pc.variables))
    return pc


def case_from_params(path_or_text: str, name: str | None = None) -> tuple[CaseSpec, ParsedCase]:
    """Build the CaseSpec the hot path consumes from a reference `Params` case file."""
    if "\n" not in path_or_text and os.path.exists(path_or_text):
        text = open(path_or_text).read()
        name = name or os.path.basename(os.path.dirname(os.path.abspath(path_or_text))) or "case"
    else:
        text = path_or_text
    pc = parse_case_text(text)
    v = pc.variables
    missing = [k for k in ("L1", "L2", "L3", "n1", "n2", "n3", "final_time", "saved_files", "delta", "tau") if k not in v]
    if missing:
        raise ExprError(f"case file lacks {missing}")
    params = {k: v[k] for k in PARAM_NAMES if k in v}
    absent = [k for k in PARAM_NAMES if k not in v]
    if absent:
        raise ExprError(f"case file lacks the model parameters {absent}")
    spec = CaseSpec(name=name or "case", n1=to_int(v["n1"]), n2=to_int(v["n2"]), n3=to_int(v["n3"]),
                    L1=v["L1"], L2=v["L2"], L3=v["L3"], calc_mode=to_int(v.get("calc_mode", 0)), params=params,
                    tau=v["tau"], tau_min=v.get("tau_min", 0.0), delta=v["delta"], final_time=v["final_time"],
                    saved_files=to_int(v["saved_files"]))
    spec.extra["icond"] = dict(pc.icond)
    return spec, pc


def initial_fields_from_formulas(spec: CaseSpec, formulas: dict, first_row: int = 0, n3_local: int | None = None):
    """The initial-condition pass of intertrack.c:1850-2020 on rows [first_row, first_row+n3_local):
    variables x,y,z (metres), _x,_y,_z (0..1), L1..L3 and every model parameter; quantities may
    refer to each other (multi-pass).  Returns u, p, gl as [n3][n2][n1] arrays."""
    n3l = spec.n3 if n3_local is None else n3_local
    _x = (0.5 + np.arange(spec.n1)) / spec.n1
    _y = (0.5 + np.arange(spec.n2)) / spec.n2
    _z = (0.5 + np.arange(n3l) + first_row) / spec.n3
    Z, Y, X = np.meshgrid(_z, _y, _x, indexing="ij")
    env = dict(spec.params)
    env.update(L1=spec.L1, L2=spec.L2, L3=spec.L3, _x=X, _y=Y, _z=Z, x=spec.L1 * X, y=spec.L2 * Y, z=spec.L3 * Z)
    for q in range(20):
        env[f"i{q + 1}"] = 1.0
    done: dict = {}
    names = ["u", "p", "gl"]
    for _ in range(len(names) + 1):
        progress = False
        for nm in names:
            if nm in done:
                continue
            try:
                val = evaluate(formulas[nm], {**env, **done})
            except ExprError:
                continue
            done[nm] = np.broadcast_to(np.asarray(val, dtype=np.float64), X.shape).copy()
            progress = True
        if len(done) == len(names) or not progress:
            break
    if len(done) != len(names):
        raise ExprError(f"cannot evaluate initial conditions for {[n for n in names if n not in done]}")
    return done["u"], done["p"], done["gl"]


@dataclass
class Restart:
    """what `set icond_file=...` [`set continue_series`] hands to a run (intertrack.c:1585-1669, 2036-2110)"""
    u: np.ndarray
    p: np.ndarray
    gl: np.ndarray
    t: float = 0.0
    tau: float | None = None
    final_time: float | None = None
    snapshot: int = 0
    total_snapshots: int | None = None


def restart_from_snapshot(spec: CaseSpec, path: str, continue_series: bool = False, first_row: int = 0,
                          n3_local: int | None = None) -> Restart:
    """Initial fields of rows [first_row, first_row+n3_local) from a snapshot dataset (NetCDF classic
    or the PFTSNAP container), as the reference loads its `icond_file`: the dataset's dimensions must
    equal the case's grid (intertrack.c:1601-1633); with `continue_series` the starting time, the
    step size, final_time and the snapshot numbering are taken from the dataset's attributes
    (:1642-1669), and the dataset must carry them."""
    from . import snapshot as sn
    with open(path, "rb") as f:
        magic = f.read(10)
    ds = sn.read_snapshot(path) if magic == sn.MAGIC else sn.read_netcdf(path)
    want = {"n1": spec.n1, "n2": spec.n2, "n3": spec.n3}
    for d, n in want.items():
        if ds["dims"].get(d) != n:
            raise ExprError(f"{path}: dimension {d} is {ds['dims'].get(d)}, the case has {n}")
    for v in ("u", "p", "gl"):
        if v not in ds["vars"] or ds["vars"][v].shape != (spec.n3, spec.n2, spec.n1):
            raise ExprError(f"{path}: variable {v} missing or not double[n3][n2][n1]")
    n3l = spec.n3 if n3_local is None else n3_local
    rows = slice(first_row, first_row + n3l)
    r = Restart(*(np.ascontiguousarray(ds["vars"][v][rows], dtype=np.float64) for v in ("u", "p", "gl")))
    if continue_series:
        a = ds["attrs"]
        missing = [k for k in ("snapshot", "total_snapshots", "t", "final_time", "tau") if k not in a]
        if missing:
            raise ExprError(f"{path}: continue_series needs the attributes {missing}")
        r.t, r.tau, r.final_time = float(a["t"]), float(a["tau"]), float(a["final_time"])
        r.snapshot, r.total_snapshots = int(a["snapshot"]), int(a["total_snapshots"])
    return r
