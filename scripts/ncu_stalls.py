"""Summarise the per-instruction stall samples of `ncu -i X.ncu-rep --page source --csv --launch-skip K --launch-count 1 > X.csv`:
    python scripts/ncu_stalls.py X.csv [top_n]
Prints the kernel, the stall-reason totals (all samples), and the top instructions with their dominant reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr) - 1]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {s: 0 for s in stalls}
samples = 0
num = lambda r, k: int(r[idx[k]]) if r[idx[k]].isdigit() else 0
for r in data:
    samples += num(r, "# Samples")
    for s in stalls:
        tot[s] += num(r, s)
print(rows[0][1][:140])
print("total samples", samples, " instructions executed (warp)", sum(num(r, "Instructions Executed") for r in data))
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {s:26s} {v:8d} {100 * v / max(samples, 1):5.1f}%")
for r in sorted(data, key=lambda r: -num(r, "# Samples"))[:topn]:
    st = {s: num(r, s) for s in stalls if num(r, s) > 0}
    main = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(r[idx["Address"]][-5:], str(num(r, "# Samples")).rjust(6), str(num(r, "Instructions Executed")).rjust(8), r[idx["Source"]].strip()[:70].ljust(70),
          [(k.replace("stall_", ""), v) for k, v in main])
