"""Static evidence that the built kernels use the Blackwell paths (run here, no GPU): per kernel family, how many tcgen05 MMA
(UTC*MMA), TMEM load (LDTM), TMA tensor load (UTMALDG), bulk copy (UBLKCP), tcgen05 commit (UTCBAR) and mbarrier (SYNCS) instructions
`cuobjdump -sass` finds in climsim_b200/libclimsim_b200.so.     python scripts/sass_evidence.py > profiles/r01_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "climsim_b200", "libclimsim_b200.so")
PAT = {"UTC*MMA": r"\bUTC[A-Z]*MMA", "UTCBAR": r"\bUTCBAR", "LDTM": r"\bLDTM", "UTMALDG": r"\bUTMALDG", "UBLKCP": r"\bUBLKCP",
       "UTMAPF/CCTL": r"\bUTMA(PF|CCTL)", "SYNCS": r"\bSYNCS", "UCGABAR": r"\bUCGABAR", "2CTA": r"\.2CTA"}
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
fam = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        key = re.sub(r"\(.*", "", name)
        key = re.sub(r"<.*", "<...>", key) if "gemm_" not in key else key
        cur = fam.setdefault(key, collections.Counter())
        cur["instances"] += 1
        continue
    if cur is None:
        continue
    for k, p in PAT.items():
        if re.search(p, line):
            cur[k] += 1
print(f"# cuobjdump -sass {os.path.basename(LIB)} (sm_100a): instruction counts per kernel instantiation")
print("# " + "  ".join(f"{k:>11s}" for k in PAT) + "  kernel")
for key, c in fam.items():
    if not any(c[k] for k in PAT):
        continue
    print("  " + "  ".join(f"{c[k]:11d}" for k in PAT) + "  " + key)
