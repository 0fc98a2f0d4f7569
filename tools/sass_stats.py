#!/usr/bin/env python3
"""Registers / spills of every stage kernel (ptxas report) and the per-plane instruction mix of the unrolled
plane bodies of merson_stage_v2 (cuobjdump -sass): fp64-pipe instructions hold the issue port for two
cycles, everything else for one -- `slots` = 2*fp64 + other is the issue-slot cost of one plane per warp."""
import re
import subprocess
import sys
from collections import Counter

LIB = sys.argv[1] if len(sys.argv) > 1 else "porousfreezethaw_b200/csrc/libpft_b200.so"
REP = "porousfreezethaw_b200/csrc/ptxas_report.txt"


def regs():
    t = open(REP).read()
    out = {}
    for b in t.split("ptxas info    : Compiling entry function ")[1:]:
        name = b.split("'")[1]
        used = re.search(r"Used (\d+) registers", b)
        sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        if used:
            out[name] = (int(used.group(1)), int(sp.group(1)) if sp else 0)
    return out


def sass(name):
    r = subprocess.run(["cuobjdump", "-sass", "-fun", name, LIB], capture_output=True, text=True)
    lines = []
    for l in r.stdout.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
        if m:
            lines.append((int(m.group(1), 16), m.group(2).strip()))
    return lines


def mix(body):
    ops = []
    for _, ins in body:
        parts = ins.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        ops.append(op)
    c = Counter(o.split(".")[0] for o in ops)
    fp64 = sum(v for k, v in c.items() if k in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU"))
    return len(ops), fp64, c


def main():
    rg = regs()
    hist_all = Counter()
    for name in sorted(rg):
        if "merson_stage" not in name:
            continue
        m = re.search(r"merson_stage_v2ILi(\d)ELi(\d)ELi(\d)ELi(\d+)ELi(\d+)ELb(\d)", name)
        if not m:
            continue
        stage, model, R, ctx, cty, nobar = (int(x) for x in m.groups())
        if model != 1 or ctx == 0:
            continue
        lines = sass(name)
        for _, ins in lines:
            parts = ins.split()
            hist_all[(parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]] += 1
        # plane bodies end with the CTA barrier (BAR.SYNC) or the arrive on the `empty` mbarrier (NOBAR)
        marks = [i for i, (_, ins) in enumerate(lines) if ("BAR.SYNC" in ins if not nobar else ("SYNCS.ARRIVE" in ins and "A1T0" in ins))]
        bodies = [lines[a + 1:b + 1] for a, b in zip(marks[:-1], marks[1:])]
        bodies = [b for b in bodies if len(b) > 150]
        if not bodies:
            continue
        best = min(bodies, key=len)
        n, f, c = mix(best)
        print(f"stage {stage} R {R} nobar {nobar} regs {rg[name][0]:3d} spill {rg[name][1]:3d} | plane body: {n} instr, {f} fp64, {n - f} other, "
              f"slots/cell {(n + f) / R:.0f} | LDS {c['LDS']} STG {c['STG']} FSEL {c['FSEL']} IMAD {c['IMAD']} MOV {c['MOV']} BRA {c['BRA']} ISETP {c['ISETP']}")
    print("opcode histogram over the listed kernels:", {k: v for k, v in hist_all.most_common() if k in
          ("UTMALDG", "UTMASTG", "SYNCS", "DADD", "DMUL", "DFMA", "LDS", "STS", "STG", "LDG", "BAR", "SHFL")})


if __name__ == "__main__":
    main()
