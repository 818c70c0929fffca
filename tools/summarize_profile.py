"""Summarise the ncu outputs of tools/gpu_profile.sh into profiles/ (tracked): per-kernel share of one denoise step
from the launch list, and the headline metrics of the --set full captures."""
import collections
import csv
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs("profiles", exist_ok=True)
out = [f"# ncu summary {tag}", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include profiled/ "
       "python tools/profile_step.py 1` (one denoise step, eager: reference pass CFG batch 8 + one view batch CFG "
       "batch 6, 512^2). Per-launch times are cold-cache and serialised: compare SHARES.", ""]
lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
out += [f"Total device time of the step: {tot / 1e3:.2f} ms over {sum(n for n, _ in agg.values())} launches", "",
        "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    out.append(f"| `{k[:70]}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for rep in sorted(f for f in os.listdir("gpurun_out") if f.endswith(".ncu-rep")):
    r = subprocess.run(["ncu", "-i", os.path.join("gpurun_out", rep), "--page", "raw", "--csv"], capture_output=True,
                       text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    out += ["", f"## {rep} (`ncu --set full --clock-control none --import-source on`)", ""]
    for row in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, row, units)}
        out.append("- " + re.sub(r"\(.*", "", d["Kernel Name"][0]).replace("<unnamed>::", "") + ": " +
                   ", ".join(f"{k.split('.')[0]}={d[k][0]} {d[k][1]}" for k in want if k in d))
open(f"profiles/{tag}_ncu_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out))
