"""Opcode mix (weighted by executed warp instructions) of one kernel from `ncu --page source --csv`: python scripts/ncu_opmix.py X.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
seen, mix, tot = set(), collections.Counter(), 0
for r in rows[2:]:
    if len(r) < len(hdr) - 1 or r[idx["Address"]] in seen: continue
    seen.add(r[idx["Address"]])
    n = int(r[idx["Instructions Executed"]]) if r[idx["Instructions Executed"]].isdigit() else 0
    src = r[idx["Source"]].strip().split()
    op = src[1] if src and src[0].startswith("@") and len(src) > 1 else (src[0] if src else "?")
    mix[op.split(".")[0]] += n; tot += n
print(rows[0][1][:120], "total warp instr", tot)
for op, n in mix.most_common(28): print(f"  {op:14s} {n:10d} {100*n/tot:5.1f}%")
