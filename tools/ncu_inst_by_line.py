#!/usr/bin/env python
"""Instruction counts (warp-level and thread-level) of an .ncu-rep by source line, sorted by warp instructions.
usage: tools/ncu_inst_by_line.py report.ncu-rep [top_n]"""
import collections, csv, os, subprocess, sys

def main(path, top=60):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    fname, si = "?", None
    I, T, S, text = collections.Counter(), collections.Counter(), collections.Counter(), {}
    kernel = None
    for r in csv.reader(out.splitlines()):
        if not r: continue
        if r[0] == "File Path": fname = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            if kernel is None: kernel = r[1]
            elif r[1] != kernel: break
            continue
        if r[0] == "Line No":
            si = r.index("Warp Stall Sampling (All Samples)"); ii = r.index("Instructions Executed"); ti = r.index("Thread Instructions Executed"); continue
        if si is None or len(r) <= si or r[0] == "": continue
        try: key = (fname, int(r[0])); s, n, t = int(r[si] or 0), int(r[ii] or 0), int(r[ti] or 0)
        except ValueError: continue
        I[key] += n; T[key] += t; S[key] += s; text[key] = r[1].strip()[:90]
    tot = sum(I.values()); ts = sum(S.values())
    print("warp-instructions", tot, " thread-instructions", sum(T.values()), " avg active", sum(T.values()) / max(1, tot))
    acc = 0
    for k, v in I.most_common(top):
        acc += v
        print(f"{100*v/tot:5.1f}% (cum {100*acc/tot:5.1f}%) stall {100*S[k]/ts:4.1f}% act {T[k]/max(1,v):4.1f}  {k[0]}:{k[1]:<4d} {text[k]}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
