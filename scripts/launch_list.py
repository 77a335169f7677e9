"""Print one training step's launches from an `ncu --metrics gpu__time_duration.sum --csv` log (gpurun_out/launches.csv)."""
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
rows = list(csv.reader(open(path)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i
        break
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
out = [(r[ki][:78], r[gi], float(r[vi].replace(",", ""))) for r in rows[start + 2:] if len(r) > vi]
idx = [i for i, o in enumerate(out) if "normalize" in o[0]]
a, b = idx[-2], idx[-1]
tot = 0.0
for o in out[a:b]:
    print(f"{o[2] / 1000:8.1f} us  {o[1]:14s} {o[0]}")
    tot += o[2]
print(f"one step: {b - a} launches, {tot / 1000:.1f} us of kernel time")
