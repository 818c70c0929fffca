#!/bin/bash
# r1m (final): ncu --set full of the attention kernel at the new launch shape (B=72 rows) -> DRAM bytes per launch into
# profiles/attn_ncu_traffic.json, then the default bench (view batch 40)
mkdir -p gpurun_out
GCB_PROFILE_BQ=72 timeout 300 ncu --set full --clock-control none -k regex:attn_tc_kernel -s 1 -c 1 -o gpurun_out/attn_b72 -f python tools/profile_attn.py 3 > gpurun_out/attn_b72_ncu.log 2>&1; echo "== ncu exit $?"
ncu -i gpurun_out/attn_b72.ncu-rep --page details > gpurun_out/attn_b72_ncu.txt 2>/dev/null
python - <<'PY'
import csv, io, json, subprocess
raw = subprocess.run(["ncu", "-i", "gpurun_out/attn_b72.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda name: (vals[hdr.index(name)], units[hdr.index(name)])
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
rd, wr = to_bytes(*get("dram__bytes_read.sum")), to_bytes(*get("dram__bytes_write.sum"))
ent = {"shape": "B=72 rows (36 views x 2 CFG halves), N=4096, d=40, 8 heads, 5 K/V sources (self + 4 cached refs)",
       "dram_bytes_per_launch": rd + wr,
       "tensor_pipe_pct": float(get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")[0]) if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr else None,
       "xu_pipe_pct": float(get("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")[0]) if "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active" in hdr else None,
       "duration": list(get("gpu__time_duration.sum")),
       "source": "GCB_PROFILE_BQ=72 ncu --set full --clock-control none -k regex:attn_tc_kernel -s 1 -c 1 python tools/profile_attn.py 3 (profiles/r1m_attn_b72_ncu.txt)"}
tj = json.load(open("profiles/attn_ncu_traffic.json"))
tj["by_rows"]["72"] = ent
json.dump(tj, open("profiles/attn_ncu_traffic.json", "w"), indent=1)
json.dump(tj, open("gpurun_out/attn_ncu_traffic.json", "w"), indent=1)
print("traffic", ent["dram_bytes_per_launch"], ent["tensor_pipe_pct"], ent["xu_pipe_pct"], ent["duration"])
PY
timeout 700 python bench.py > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo "== bench exit $?"; head -c 300 gpurun_out/bench_r1m.json; echo
python -c "
import json; b=json.load(open('gpurun_out/bench_r1m.json')); print('value', b['value'], 'e2e', b['e2e'], 'roof', b['roofline']['achieved'], b['roofline']['frac'], b['roofline']['traffic'], 'cpu', b['cpu_baseline']['value'], 'ft', b['extra']['finetune'])"
