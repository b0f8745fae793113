#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics ...` launch list by kernel: launches, total time, share; optional DRAM bytes.
usage: tools/launch_summary.py launches.csv"""
import collections, csv, re, sys

def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, ni, mi, ui, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    t, n, dr, dw = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float), collections.defaultdict(float)
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[start:]:
        if len(r) <= mi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        v = float(r[mi].replace(",", "")) * scale.get(r[ui], 1.0)
        if r[ni] == "gpu__time_duration.sum":
            t[name] += v
            n[name] += 1
        elif r[ni] == "dram__bytes_read.sum":
            dr[name] += v
        elif r[ni] == "dram__bytes_write.sum":
            dw[name] += v
    tot = sum(t.values())
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}  dram rd/wr per launch (MB)")
    for k, v in sorted(t.items(), key=lambda x: -x[1]):
        extra = f"  {dr[k] / n[k] / 1e6:9.1f} / {dw[k] / n[k] / 1e6:9.1f}" if k in dr else ""
        print(f"{k:44s} {n[k]:8d} {v / 1e3:10.2f} {100 * v / tot:6.1f}% {v / n[k]:9.1f}{extra}")
    print(f"{'total':44s} {sum(n.values()):8d} {tot / 1e3:10.2f}")

if __name__ == "__main__":
    main(sys.argv[1])
