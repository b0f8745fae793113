#!/usr/bin/env python
"""Aggregate warp-stall samples of an .ncu-rep by CUDA source line (needs -lineinfo + --import-source on).
usage: tools/ncu_hot_lines.py report.ncu-rep [top_n]"""
import collections
import csv
import os
import subprocess
import sys


def main(path, top=45):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = csv.reader(out.splitlines())
    fname, si, ii, ti = "?", None, None, None
    samples, insts, thr, text = collections.Counter(), collections.Counter(), collections.Counter(), {}
    kernel = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            if kernel is None:
                kernel = r[1]
            elif r[1] != kernel:
                break  # first kernel instance only
            continue
        if r[0] == "Line No":
            si = r.index("Warp Stall Sampling (All Samples)")
            ii = r.index("Instructions Executed")
            ti = r.index("Thread Instructions Executed")
            continue
        if si is None or len(r) <= si or r[0] == "":
            continue
        try:
            key = (fname, int(r[0]))
            s, n, t = int(r[si] or 0), int(r[ii] or 0), int(r[ti] or 0)
        except ValueError:
            continue
        samples[key] += s
        insts[key] += n
        thr[key] += t
        text[key] = r[1].strip()[:100]
    total = sum(samples.values())
    by_file = collections.Counter()
    for (f, _), v in samples.items():
        by_file[f] += v
    print("total samples", total, " warp-instructions", sum(insts.values()))
    print("by file:", ", ".join(f"{f} {100 * v / total:.1f}%" for f, v in by_file.most_common()))
    for k, v in samples.most_common(top):
        act = thr[k] / insts[k] if insts[k] else 0
        print(f"{100 * v / total:5.1f}%  inst {insts[k]:9d} act {act:4.1f}  {k[0]}:{k[1]:<4d} {text[k]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
