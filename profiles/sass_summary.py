#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md): runs `cuobjdump -sass`
on the built library; no GPU needed.  Usage: python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

LIB = Path(__file__).resolve().parent.parent / "chimeralm_b200" / "libchimeralm_b200.so"
WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "MUFU.TANH",
        "MUFU.EX2", "MUFU.RCP", "FFMA2", "FMUL2", "FADD2", "HMMA", "LDG", "STG", "LDS", "STS", "BAR.SYNC", "ELECT", "USETMAXREG"]
out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
counts, total, name = collections.defaultdict(collections.Counter), collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("clm::", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and name:
        op = m.group(1)
        total[name] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[name][w] += 1
print(f"# cuobjdump -sass {LIB.name} (sm_100a): instruction counts per kernel; columns with no hit anywhere are dropped")
cols = [w for w in WANT if any(c[w] for c in counts.values())]
print("kernel".ljust(64) + "".join(w.rjust(11) for w in ["SASS"] + cols))
for k in sorted(total, key=lambda k: -total[k]):
    if total[k] < 40 and not any(counts[k].values()):
        continue
    print(k[:63].ljust(64) + str(total[k]).rjust(11) + "".join(str(counts[k][w] or ".").rjust(11) for w in cols))
