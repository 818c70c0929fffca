"""Summarise `ncu -i X.ncu-rep --page source --csv` of a kernel: the top stalled instructions with their dominant
stall reasons, per-opcode execution counts, and the instructions executed per hot-loop trip.
usage: python tools/summarize_source.py source.csv [executions of one loop trip]"""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if len(r) > 10)
hdr = rows[h]
ix = {n: i for i, n in enumerate(hdr)}
ins = [r for r in rows[h + 1:] if len(r) == len(hdr)]
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = sum(int(r[ix["# Samples"]]) for r in ins)
print(f"{len(ins)} instructions, {tot} samples")
agg = collections.Counter()
for r in ins:
    for s in stalls:
        agg[s] += int(r[ix[s]] or 0)
print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in agg.most_common(9)))
ex = collections.Counter()
hot = int(sys.argv[2]) if len(sys.argv) > 2 else max(int(r[ix["Instructions Executed"]]) for r in ins)
nhot = 0
for r in ins:
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0]
    n = int(r[ix["Instructions Executed"]])
    ex[op] += n
    if n >= 0.9 * hot:
        nhot += 1
print(f"instructions executed >= 0.9 x {hot} times (hot loop body): {nhot}")
print("executed by opcode (x hot):", ", ".join(f"{k} {v / hot:.1f}" for k, v in ex.most_common(22)))
print("top stalled instructions:")
for j, r in sorted(enumerate(ins), key=lambda t: -int(t[1][ix["# Samples"]]))[:24]:
    n = int(r[ix["# Samples"]])
    why = sorted(((int(r[ix[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"  #{j:4d} {100 * n / tot:5.2f}%  {r[ix['Source']].strip()[:60]:60s} " + ", ".join(f"{w} {c}" for c, w in why))
