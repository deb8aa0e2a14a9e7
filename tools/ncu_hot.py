#!/usr/bin/env python
"""Hot source lines of one kernel in an .ncu-rep (read on the CPU box): instructions executed and stall samples per CUDA line.

usage: python tools/ncu_hot.py gpurun_out/x.ncu-rep <kernel regex> [top N] [launch index among matches]
Needs -lineinfo at compile time and --import-source on at capture time."""
import csv
import io
import subprocess
import sys


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{pat}",
                          "--launch-count", "1"] , capture_output=True, text=True).stdout
    # the dump is a sequence of per-file tables: "File Name",... then a header row "Line No","Source",metrics...
    rows = list(csv.reader(io.StringIO(raw)))
    out = []
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] in ("File Name", "File Path"):
            fname = r[1].split("/")[-1]; hdr = None; continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = {}
            for i, h in enumerate(r):
                hdr.setdefault(h, i)
            continue
        if hdr is None or fname is None or not r[0].isdigit():  # rows without a line number are the SASS detail
            continue
        try:
            inst = float(r[hdr["Instructions Executed"]]); samp = float(r[hdr["# Samples"]])
        except (KeyError, ValueError, IndexError):
            continue
        if inst or samp:
            d = {"file": fname, "line": r[0], "src": r[1].strip()[:110], "inst": inst, "samples": samp}
            for k in ("stall_long_sb", "stall_short_sb", "stall_mio", "stall_lg", "stall_barrier", "stall_math", "stall_wait",
                      "stall_not_selected", "stall_branch_resolving"):
                if k in hdr:
                    try: d[k] = float(r[hdr[k]])
                    except ValueError: d[k] = 0.0
            out.append(d)
    ti, ts = sum(d["inst"] for d in out), sum(d["samples"] for d in out)
    print(f"total warp instructions {ti:.4g}, samples {ts:.4g}")
    print("by instructions:")
    for d in sorted(out, key=lambda d: -d["inst"])[:top]:
        print(f"  {100*d['inst']/ti:5.1f}% inst {100*d['samples']/max(ts,1):5.1f}% smp  {d['file']}:{d['line']:>4}  {d['src']}")
    print("by stall samples:")
    for d in sorted(out, key=lambda d: -d["samples"])[:top]:
        st = {k[6:]: v for k, v in d.items() if k.startswith("stall_") and v > 0.15 * d["samples"]}
        print(f"  {100*d['samples']/max(ts,1):5.1f}% smp {100*d['inst']/ti:5.1f}% inst  {d['file']}:{d['line']:>4}  {d['src'][:70]}  {st}")


if __name__ == "__main__":
    main()
