#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a per-kernel table for profiles/.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_step_summary.md
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("lts__t_sector_hit_rate.pct", "l2hit_%"),
    ("sm__inst_executed_pipe_xu.sum", "xu_inst"),
    ("sm__inst_executed.sum", "inst"),
]


def conv(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return v
    u = unit.lower()
    if u in ("ns", "nsecond"): x /= 1e3
    elif u in ("ms", "msecond"): x *= 1e3
    elif u in ("s", "second"): x *= 1e6
    elif u == "byte": x /= 1e6
    elif u == "kbyte": x /= 1e3
    elif u == "gbyte": x *= 1e3
    return f"{x:.2f}" if abs(x) < 1e6 else f"{x:.3e}"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["| kernel | grid | block | " + " | ".join(n for _, n in COLS) + " |", "|" + "---|" * (3 + len(COLS))]
    for d in data:
        name = d[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = [conv(d[idx[m]], units[idx[m]]) if m in idx else "-" for m, _ in COLS]
        lines.append(f"| {name} | {d[idx['Grid Size']]} | {d[idx['Block Size']]} | " + " | ".join(cells) + " |")
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    open(out, "w").write(f"# ncu --set full summary of `{rep}`\n\n{note}\n\n" + "\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
