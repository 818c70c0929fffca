#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_*] --csv` launch list: one line per launch (or, with
--by-name, per kernel name) with time and DRAM bytes."""
import collections
import csv
import re
import sys


def load(path):
    allrows = list(csv.reader(open(path)))
    hdr = next(r for r in allrows if r and r[0] == "ID")
    col = {n: i for i, n in enumerate(hdr)}
    kn, gs, bs, mn, mv = col["Kernel Name"], col["Grid Size"], col["Block Size"], col["Metric Name"], col["Metric Value"]
    by = collections.OrderedDict()
    for r in allrows:
        if len(r) != len(hdr) or not r[0].isdigit():
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("<unnamed>::", "").replace("void ", "")
        by.setdefault(int(r[0]), {"name": name, "grid": r[gs], "block": r[bs]})[r[mn]] = float(r[mv].replace(",", ""))
    return by


def main():
    path = sys.argv[1]
    by_name = "--by-name" in sys.argv
    lo = int(sys.argv[sys.argv.index("--from") + 1]) if "--from" in sys.argv else 0
    hi = int(sys.argv[sys.argv.index("--to") + 1]) if "--to" in sys.argv else 1 << 30
    by = {k: v for k, v in load(path).items() if lo <= k < hi}
    tot = sum(v.get("gpu__time_duration.sum", 0) for v in by.values()) / 1e3
    if by_name:
        agg = collections.OrderedDict()
        for v in by.values():
            a = agg.setdefault(v["name"][:60], [0, 0.0, 0.0, 0.0])
            a[0] += 1
            a[1] += v.get("gpu__time_duration.sum", 0) / 1e3
            a[2] += v.get("dram__bytes_read.sum", 0) / 1e6
            a[3] += v.get("dram__bytes_write.sum", 0) / 1e6
        for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"{a[1]:10.1f} us {100 * a[1] / tot:5.1f}%  x{a[0]:<5d} rd {a[2]:9.1f} MB wr {a[3]:9.1f} MB  {n}")
    else:
        for k, v in by.items():
            print(f"{k:4d} {v.get('gpu__time_duration.sum', 0) / 1e3:9.1f} us  rd {v.get('dram__bytes_read.sum', 0) / 1e6:8.1f} MB "
                  f"wr {v.get('dram__bytes_write.sum', 0) / 1e6:8.1f} MB  {v['grid']:>16s} {v['name'][:60]}")
    print(f"total {tot:.1f} us over {len(by)} launches")


if __name__ == "__main__":
    main()
