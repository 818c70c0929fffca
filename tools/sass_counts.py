#!/usr/bin/env python
"""Instruction-class counts per kernel from `cuobjdump -sass` of the in-tree library -> profiles/<tag>_sass_counts.md
(tag = argv[1], default r3)."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(REPO, "gaussctrl_b200", "libgaussctrl_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "HMMA", "HFMA2", "MUFU.EX2", "SYNCS", "VOTE", "ATOMS", "ATOMG", "REDG", "BAR")


def demangle(n):
    for tool in ("cu++filt", "c++filt"):
        try:
            r = subprocess.run([tool, n], capture_output=True, text=True)
            if r.returncode == 0 and r.stdout.strip():
                return r.stdout.strip()
        except FileNotFoundError:
            continue
    return n


rows = []
for part in re.split(r"\n\s*Function : ", txt)[1:]:
    name = part.split("\n", 1)[0].strip()
    c = collections.Counter()
    n = 0
    for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", part, re.M):
        n += 1
        for k in KEYS:
            if m.group(1).startswith(k):
                c[k] += 1
    d = demangle(name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
    d = re.sub(r"\(.*", "", d.replace("(int)", "").replace("(bool)", ""))
    rows.append((d, n, c))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r3"
out = [f"# SASS evidence ({TAG}): instruction counts per kernel of `libgaussctrl_b200.so`", "",
       "`cuobjdump -sass gaussctrl_b200/libgaussctrl_b200.so` (nvcc 12.9, `-gencode arch=compute_100a,code=sm_100a`), counted by "
       "`tools/sass_counts.py`.", "UTCHMMA = `tcgen05.mma` (kind::f16), LDTM/STTM = `tcgen05.ld/st` (TMEM), UTMALDG/UTMASTG = TMA bulk "
       "tensor load/store, SYNCS = mbarrier ops, HMMA = `mma.sync`, HFMA2 = packed-half FMA (the attention kernel's polynomial exponentials), MUFU.EX2 = exp2 on the special-function unit, VOTE = warp ballots "
       "(radix ranking), ATOMS/ATOMG/REDG = shared / global atomics.", "",
       "| kernel | SASS instr | UTCHMMA | LDTM | STTM | UTMALDG | UTMASTG | HMMA | HFMA2 | MUFU.EX2 | SYNCS | VOTE | ATOMS/ATOMG/REDG |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
for d, n, c in sorted(rows, key=lambda r: -r[1]):
    out.append(f"| `{d[:70]}` | {n} | {c['UTCHMMA']} | {c['LDTM']} | {c['STTM']} | {c['UTMALDG']} | {c['UTMASTG']} | {c['HMMA']} | {c['HFMA2']} | "
               f"{c['MUFU.EX2']} | {c['SYNCS']} | {c['VOTE']} | {c['ATOMS']}/{c['ATOMG']}/{c['REDG']} |")
path = os.path.join(REPO, "profiles", f"{TAG}_sass_counts.md")
open(path, "w").write("\n".join(out) + "\n")
print("\n".join(out[:24]))
