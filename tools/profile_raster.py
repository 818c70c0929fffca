"""Render a few 512x512 views of the synthetic 1M-Gaussian scene (eval mode: fused rgb+depth) for ncu / timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_scene, orbit_c2w
from gaussctrl_b200.gc_model import render_gaussians
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
P = {k: v.cuda() for k, v in synthetic_scene(n).items()}
bg = torch.zeros(3, device="cuda")
views = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for i in range(views):
    if i == 1:
        torch.cuda.synchronize(); torch.cuda.nvtx.range_push("profiled"); t0 = time.perf_counter()
    with torch.no_grad():
        out = render_gaussians(P, orbit_c2w(i, 40), 539.05, 538.17, 258.74, 239.35, 512, 512, 3, bg)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / max(1, views - 1)
vis = int((out["accumulation"] > 0).sum())
print(f"wall per view {dt*1e3:.3f} ms; covered pixels {vis}")
