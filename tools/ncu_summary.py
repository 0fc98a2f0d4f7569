#!/usr/bin/env python
"""tools/ncu_summary.py -- condense an .ncu-rep (raw + source pages) into the few numbers the
design decisions use: duration, DRAM bytes, pipe utilisation, stall mix, hottest SASS lines."""
import csv
import io
import subprocess
import sys


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
            "launch__block_size", "launch__grid_size", "smsp__inst_executed.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps"]
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
    for k, r in enumerate(rows[2:]):
        print(f"== launch {k}: {r[idx['Kernel Name']]}")
        for w in want:
            if w in idx:
                print(f"   {w:66s} {r[idx[w]]} {units[idx[w]]}")
        vals = sorted(((float(r[idx[n]]), n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for n in stall), reverse=True)
        print("   stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in vals[:8]))
    if top <= 0:
        return
    for k in range(len(rows) - 2):
        src = page(rep, "source", ["--launch-skip", str(k), "--launch-count", "1"])
        if len(src) < 3:
            continue
        h = src[1]
        ix = {n: i for i, n in enumerate(h)}
        data = {}
        for r in src[2:]:
            if len(r) >= len(h) - 2 and r[0].startswith("0x"):
                data.setdefault(r[0], r)
        data = list(data.values())
        tot = sum(int(r[ix["# Samples"]]) for r in data)
        st = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        print(f"== source launch {k}: {src[0][1][:70]}  samples {tot}  instr {sum(int(r[ix['Instructions Executed']]) for r in data)}")
        for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
            s = sorted(((int(r[ix[n]]), n[6:]) for n in st if r[ix[n]].isdigit() and int(r[ix[n]]) > 0), reverse=True)[:2]
            print(f"   {int(r[ix['# Samples']]):6d} {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']].strip()[:58]:58s} {s}")


if __name__ == "__main__":
    main()
