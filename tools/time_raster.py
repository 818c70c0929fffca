#!/usr/bin/env python
"""Eval-render timing of the 1 M-Gaussian bench scene: V views enqueued back to back (no host synchronisation inside),
CUDA events around the whole loop -> ms per view and achieved GB/s against the SURVEY §8d byte formula."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import orbit_c2w, synthetic_scene  # noqa: E402
from gaussctrl_b200 import gsplat_ops as go  # noqa: E402
from gaussctrl_b200.gc_model import render_gaussians  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    V = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    P = {k: v.cuda() for k, v in synthetic_scene(n, seed=0).items()}
    bg = torch.zeros(3, device="cuda")
    intr = (539.05, 538.17, 258.74, 239.35)
    c2ws = [torch.cat([orbit_c2w(i, V), torch.tensor([[0, 0, 0, 1.0]])]) for i in range(V)]

    n_streams = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    streams = [torch.cuda.Stream() for _ in range(n_streams)] if n_streams > 1 else None

    def run(defer):
        outs = []
        with torch.no_grad():
            if streams is None or not defer:
                for c in c2ws:
                    outs.append(render_gaussians(P, c, *intr, 512, 512, 3, bg, state={"defer_check": defer}))
            else:
                outs = go.render_views_multistream(
                    lambda c: render_gaussians(P, c, *intr, 512, 512, 3, bg, state={"defer_check": True}), c2ws, streams)
        if defer:
            go.check_deferred_overflow()
        return outs

    run(True)
    torch.cuda.synchronize()
    res = {}
    for name, defer in (("pipelined", True), ("checked_per_view", False)):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        run(defer)
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"gpu_ms_per_view": e0.elapsed_time(e1) / V, "wall_ms_per_view": (time.perf_counter() - t0) * 1e3 / V}
    M = go.LAST_M[0]
    byts = 244.0 * n + 48.0 * n + 152.0 * M + 5.2e6
    # batched C-side loop (gcb_render_eval_batch), what GaussCtrlPipeline.render_views uses
    from gaussctrl_b200.gc_model import projection_matrix, viewmat_from_c2w
    import math
    W = H = 512
    vms = [viewmat_from_c2w(c) for c in c2ws]
    pm = projection_matrix(0.001, 1000, 2 * math.atan(W / (2 * intr[0])), 2 * math.atan(H / (2 * intr[1])))
    pms = torch.stack([pm @ v for v in vms])
    orgs = torch.stack([c[:3, 3] for c in c2ws])
    it = torch.tensor([list(intr)] * V)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        with torch.no_grad():
            go.render_eval_batch(P, (torch.stack(vms), pms, orgs, it), H, W, 3, bg, streams=streams)
        e1.record()
        torch.cuda.synchronize()
        go.check_deferred_overflow()
        res["batched"] = {"gpu_ms_per_view": e0.elapsed_time(e1) / V, "wall_ms_per_view": (time.perf_counter() - t0) * 1e3 / V}
    res["streams"] = n_streams
    res["intersections_last_view"] = M
    res["algorithmic_bytes"] = byts
    res["achieved_gbs_pipelined"] = byts / (res["pipelined"]["gpu_ms_per_view"] / 1e3) / 1e9
    res["achieved_gbs_batched"] = byts / (res["batched"]["gpu_ms_per_view"] / 1e3) / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    main()
