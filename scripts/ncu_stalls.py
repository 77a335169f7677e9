"""Summarise the per-instruction stall samples of `ncu -i X.ncu-rep --page source --csv > X.csv`: python scripts/ncu_stalls.py X.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]; data = rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = {s: 0 for s in stalls}
samples = 0
for r in data:
    if not r[idx['# Samples']].isdigit(): continue
    samples += int(r[idx['# Samples']])
    for s in stalls:
        if r[idx[s]].isdigit(): tot[s] += int(r[idx[s]])
print(rows[0][1][:120]); print("total samples", samples)
for s, v in sorted(tot.items(), key=lambda x: -x[1])[:12]: print(f"{s:28s} {v:8d} {100*v/samples:5.1f}%")
top = sorted([r for r in data if r[idx['# Samples']].isdigit()], key=lambda r: -int(r[idx['# Samples']]))[:topn]
for r in top:
    st = {s: int(r[idx[s]]) for s in stalls if r[idx[s]].isdigit() and int(r[idx[s]]) > 0}
    main = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(r[idx['Address']][-5:], r[idx['# Samples']].rjust(6), r[idx['Instructions Executed']].rjust(8), r[idx['Source']][:64].ljust(64), main)
