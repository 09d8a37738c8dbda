#!/usr/bin/env python
"""Summarises an `ncu --set full --import-source on` report per CUDA source line.

    python scripts/ncu_source_summary.py gpurun_out/prof.ncu-rep [top_n]

Runs `ncu -i REP --page source --csv --print-source cuda,sass`, sums instructions executed and
warp-stall samples per (file, line), and prints the hottest lines plus per-file totals and the
stall-reason mix.  This is how the per-round profiles under profiles/ are produced.
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    per_line = defaultdict(lambda: [0, 0, ""])  # (file, line) -> [inst, samples, text]
    per_file = defaultdict(lambda: [0, 0])
    stalls = defaultdict(int)
    fname, hdr, cur = None, None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr is None:
            continue
        if r[0]:
            cur = (fname, int(r[0]))
            per_line[cur][2] = r[1].strip()[:110]
            continue  # the line's own row repeats the sum of its SASS rows
        if cur is None or len(r) <= i_inst or not r[i_inst].isdigit():
            continue
        inst, samp = int(r[i_inst]), int(r[i_samp]) if r[i_samp].isdigit() else 0
        per_line[cur][0] += inst
        per_line[cur][1] += samp
        per_file[fname][0] += inst
        per_file[fname][1] += samp
        for i, h in stall_cols:
            if i < len(r) and r[i].isdigit():
                stalls[h] += int(r[i])
    tot_i = sum(v[0] for v in per_file.values()) or 1
    tot_s = sum(v[1] for v in per_file.values()) or 1
    print(f"total warp instructions {tot_i}  stall samples {tot_s}")
    print("per file:")
    for f, (i, s) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:22s} inst {100 * i / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%")
    print("stall mix (all samples):")
    ts = sum(stalls.values()) or 1
    for h, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {h:24s} {100 * v / ts:5.1f}%")
    print(f"top {top} lines by stall samples:")
    for (f, ln), (i, s, t) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"  {f}:{ln:<4d} samp {100 * s / tot_s:5.1f}% inst {100 * i / tot_i:5.1f}%  {t}")
    print(f"top {top} lines by instructions:")
    for (f, ln), (i, s, t) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {f}:{ln:<4d} inst {100 * i / tot_i:5.1f}% samp {100 * s / tot_s:5.1f}%  {t}")


if __name__ == "__main__":
    main()
