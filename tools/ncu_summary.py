#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into a per-kernel table for profiles/ and into the JSON that
bench.py loads its roofline side-fields from (DRAM traffic per launch, issue-active %, shared-memory wavefront %).

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r2_step_ncu_full.md "note" \
           [--json profiles/ncu_step.json] [--launches gpurun_out/launches.csv]
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess

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
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wave_%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_%"),
    ("lts__t_sector_hit_rate.pct", "l2hit_%"),
    ("smsp__inst_executed.sum", "warp_inst"),
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


def short_name(full):
    n = full.split("(")[0].replace("void ", "").replace("dvs::", "").strip()
    return re.sub(r"\(int\)|\(bool\)", "", n).replace(" ", "")


def launch_list(path):
    """ncu --metrics gpu__time_duration.sum --csv launch list -> {kernel: [us, ...]}"""
    out = {}
    rows = [r for r in csv.reader(open(path)) if r]
    hdr = next((i for i, r in enumerate(rows) if "Kernel Name" in r), None)
    if hdr is None:
        return out
    h = {c: i for i, c in enumerate(rows[hdr])}
    for r in rows[hdr + 1:]:
        if len(r) <= h["Metric Value"]:
            continue
        if r[h["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[h["Metric Value"]].replace(",", ""))
        unit = r[h["Metric Unit"]].lower()
        v = v / 1e3 if unit.startswith("n") else v * 1e3 if unit.startswith("m") else v
        out.setdefault(short_name(r[h["Kernel Name"]]), []).append(v)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep"); ap.add_argument("out"); ap.add_argument("note", nargs="?", default="")
    ap.add_argument("--json", default=None); ap.add_argument("--launches", default=None)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = ["| kernel | grid | block | " + " | ".join(n for _, n in COLS) + " |", "|" + "---|" * (3 + len(COLS))]
    agg = {}
    for d in data:
        name = short_name(d[idx["Kernel Name"]])
        cells = [conv(d[idx[m]], units[idx[m]]) if m in idx else "-" for m, _ in COLS]
        lines.append(f"| {name} | {d[idx['Grid Size']]} | {d[idx['Block Size']]} | " + " | ".join(cells) + " |")
        k = agg.setdefault(name, {"launches": 0, "time_us": 0.0, "dram_MB": 0.0, "issue_active_pct": 0.0, "smem_wavefront_pct": 0.0,
                                  "warp_inst": 0.0})
        f = lambda c: float(cells[[n for _, n in COLS].index(c)]) if cells[[n for _, n in COLS].index(c)] not in ("-", "") else 0.0
        k["launches"] += 1
        k["time_us"] += f("time_us"); k["dram_MB"] += f("dram_rd_MB") + f("dram_wr_MB")
        k["issue_active_pct"] += f("issue_%"); k["smem_wavefront_pct"] += f("smem_wave_%"); k["warp_inst"] += f("warp_inst")
    for k in agg.values():
        for f in ("time_us", "dram_MB", "issue_active_pct", "smem_wavefront_pct", "warp_inst"):
            k[f] = round(k[f] / k["launches"], 3)
    text = f"# ncu --set full summary of `{a.rep}`\n\n{a.note}\n\n" + "\n".join(lines) + "\n"
    ll = None
    if a.launches and os.path.exists(a.launches):
        ll = launch_list(a.launches)
        text += (f"\n## Launch list of the same command (`{a.launches}`, `--metrics gpu__time_duration.sum --clock-control none`): "
                 "mean device time per kernel, cold-cache and serialised — compare SHARES with the CUDA-event stage times\n\n"
                 "| kernel | launches | mean us | share of one step |\n|---|---|---|---|\n")
        per_step = {k: sum(v) / len(v) for k, v in ll.items()}
        tot = sum(per_step.values())
        for k, v in sorted(per_step.items(), key=lambda kv: -kv[1]):
            text += f"| {k} | {len(ll[k])} | {v:.2f} | {100 * v / tot:.1f}% |\n"
    open(a.out, "w").write(text)
    if a.json:
        json.dump({"source": os.path.basename(a.rep), "note": a.note, "summary": os.path.basename(a.out), "kernels": agg,
                   "launch_list_mean_us": ({k: round(sum(v) / len(v), 3) for k, v in ll.items()} if ll else None)},
                  open(a.json, "w"), indent=1)
    print(text)


if __name__ == "__main__":
    main()
