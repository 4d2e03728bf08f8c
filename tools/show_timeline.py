#!/usr/bin/env python
"""Prints one steady-state step of a tools/fit_timeline.py CSV: start, end, duration, stream, kernel."""
import csv
import sys

def short(k):
    k = k.replace("dvt::(anonymous namespace)::", "").replace("dvt::", "").replace("void ", "")
    return k.split("(")[0][:60]

path, phase = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "1")
R = [r for r in csv.DictReader(open(path)) if r["phase"] == phase]
enc = [i for i, r in enumerate(R) if "fit_encode_kernel" in r["kernel"]]
a, b = enc[24], enc[25]
t0 = float(R[a]["start_us"])
print(f"step length {float(R[b]['start_us']) - t0:.1f} us")
for r in R[a:b + 1]:
    s = float(r["start_us"]) - t0
    d = float(r["dur_us"])
    print(f"{s:8.1f} {s + d:8.1f} {d:6.1f}  st{r['stream']:>3}  {short(r['kernel'])}")
