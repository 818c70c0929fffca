#!/usr/bin/env python
"""bench.py - edited-views/sec of the GaussCtrl hot path (512^2, 20 DDIM steps, chunk=3, R=4) on N B200s.

One "step" = one full pass of `edit_images` over the workload's V views (cross-view-attention DDIM sampling through
ControlNet+UNet for every view, then VAE decode): `value` = V / step time with z_T / depth already in HBM;
`e2e` = the same pass through GaussCtrlPipeline.edit_images() with host (numpy) train_data in and host images out.
`--impl reference` times the reference's algorithm (oracle restatement: un-fused 5-pass attention, refs recomputed
per chunk) on the host CPU cores on a bounded sample.  See DESIGN.md §6 for what each key means."""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "edited-views/sec (512^2, 20 DDIM steps, chunk=3)"
UNIT = "views/s"
S_STEPS, CHUNK, REFS, GUIDANCE = 20, 3, 4, 5.0
HW_LAT, HW_IMG = 64, 512
# algorithmic FLOPs per batch row per denoise step: ControlNet + UNet with cross-view attention
ATTN_N, ATTN_C, ATTN_D, ATTN_HEADS = 4096, 320, 40, 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--views", type=int, default=40, help="views per GPU (bear-like scene: 40)")
    ap.add_argument("--gaussians", type=int, default=1_000_000)
    ap.add_argument("--ddim-steps", type=int, default=S_STEPS)
    ap.add_argument("--view-batch", type=int, default=40,
                    help="views denoised per launch in the refs-once schedule (results do not depend on it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------- synthetic workload
def synthetic_scene(n: int, seed: int = 0):
    """means U(-1,1)^3, log-scales N(ln 0.01, 0.3^2), random unit quats, opacity logits U(-2,4), SH deg 3."""
    g = torch.Generator().manual_seed(seed)
    return dict(means=torch.rand((n, 3), generator=g) * 2 - 1,
                scales=torch.randn((n, 3), generator=g) * 0.3 + math.log(0.01),
                quats=torch.randn((n, 4), generator=g),
                opacities=torch.rand((n, 1), generator=g) * 6 - 2,
                features_dc=torch.randn((n, 3), generator=g) * 0.5,
                features_rest=torch.randn((n, 15, 3), generator=g) * 0.05)


def orbit_c2w(i: int, n: int, radius: float = 2.2):
    az = 2 * math.pi * i / n
    eye = torch.tensor([radius * math.cos(az), radius * math.sin(az), 0.5])
    fwd = -eye / eye.norm()
    right = torch.linalg.cross(fwd, torch.tensor([0.0, 0.0, 1.0]))
    right = right / right.norm()
    up = torch.linalg.cross(right, fwd)
    m = torch.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = right, up, -fwd, eye
    return m[:3]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(r) > 3 + j and r[3 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def reference_step_cpu(models, seed: int = 0, chunk: int = CHUNK):
    """One DDIM step of one reference-schedule chunk (R=4 refs + `chunk` views, CFG batch 2(R+chunk), literal 5-pass
    attention) on the host cores.  Returns seconds."""
    from oracle import pipeline as opipe, sd15
    unet, cnet = models
    g = torch.Generator().manual_seed(seed)
    F = REFS + chunk
    lat = torch.randn((F, 4, HW_LAT, HW_LAT), generator=g)
    disp = torch.rand((F, 1, HW_IMG, HW_IMG), generator=g).repeat(1, 3, 1, 1)
    pos, neg = torch.randn((1, 77, 768), generator=g), torch.randn((1, 77, 768), generator=g)
    t0 = time.perf_counter()
    opipe.edit_chunk(unet, cnet, None, sd15.DDIMTables(), lat, disp, pos, neg, 1, GUIDANCE, REFS, decode=False)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import sd15
    unet, cnet, _ = sd15.seeded_models(seed=0, with_vae=False)
    cores = torch.get_num_threads()
    # bounded sample: the full chunk (c=3) when few steps are requested, a 1-view chunk otherwise, so that the whole
    # --steps/--warmup run stays within a few minutes on the host cores (cost is proportional to the CFG batch)
    c_s = CHUNK if args.steps + args.warmup <= 4 else 1
    for _ in range(args.warmup):
        reference_step_cpu((unet, cnet), chunk=c_s)
    times = [reference_step_cpu((unet, cnet), chunk=c_s) for _ in range(args.steps)]
    t = sum(times) / len(times)
    vps = c_s / (t * S_STEPS)  # a chunk edits c views in S such steps (VAE decode and rasterisation not counted)
    sample = (f"1 DDIM step of one reference-schedule chunk (R={REFS}+c={c_s} frames, CFG batch {2 * (REFS + c_s)}, "
              f"literal 5-pass attention, fp32) per step; views/s = c / (S x t_step), S={S_STEPS}")
    line = {"impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": f"bear-like synthetic scene: {args.gaussians} Gaussians, {args.views} views/GPU at 512x512, "
                        f"chunk_size={CHUNK}, ref_view_num={REFS}, {args.ddim_steps} DDIM steps, guidance {GUIDANCE}, "
                        f"seeded random-init SD1.5 UNet + ControlNet-depth + VAE (BASELINE.json configs[1])",
            "views_per_gpu": args.views, "gaussians": args.gaussians, "ddim_steps": args.ddim_steps,
            "l2_policy": "inputs larger than L2 (activations of one denoise step ~ 6 GB > 126 MB)"}


# ----------------------------------------------------------------------------------------------- B200 arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from gaussctrl_b200 import ops
    from gaussctrl_b200._compat import Cameras
    from gaussctrl_b200.gc_model import GaussCtrlModel, GaussCtrlModelConfig
    from gaussctrl_b200.gc_pipeline import GaussCtrlPipeline, GaussCtrlPipelineConfig, SimpleDataManager
    from gaussctrl_b200.sd15_spec import synthetic_weights

    V, S = args.views * world, args.ddim_steps   # weak scaling: views grow with the GPU count, references are shared
    # ---- scene + cameras + pipeline (public API objects)
    scene = synthetic_scene(args.gaussians, seed=0)
    model = GaussCtrlModel(GaussCtrlModelConfig(), num_points=args.gaussians)
    with torch.no_grad():
        for k, v in scene.items():
            getattr(model, k).data = v
    model = model.to(dev)
    model.background_color = torch.zeros(3)
    c2w = torch.stack([orbit_c2w(i, V) for i in range(V)])
    cams = Cameras(c2w, 539.05, 538.17, 258.74, 239.35, HW_IMG, HW_IMG)
    dm = SimpleDataManager(cams)
    cfg = GaussCtrlPipelineConfig(edit_prompt="a photo of a polar bear in the forest",
                                  reverse_prompt="a photo of a bear statue in the forest", guidance_scale=GUIDANCE,
                                  num_inference_steps=S, chunk_size=CHUNK, ref_view_num=REFS)
    pipe = GaussCtrlPipeline(cfg, dev, world_size=world, local_rank=local, datamanager=dm, model=model,
                             weights=synthetic_weights(0))
    pipe.view_batch = args.view_batch

    # ---- stage-A products (untimed setup): rasterise depth for the ControlNet condition; z_T ~ N(0,1)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    g = torch.Generator().manual_seed(1)
    raster_ms = []
    for i in range(V):
        e0, e1 = ev(), ev()
        e0.record()
        out = model.get_outputs_for_camera(cams[i])
        e1.record()
        torch.cuda.synchronize()
        raster_ms.append(e0.elapsed_time(e1))
        dm.train_data[i]["unedited_image"] = out["rgb"].to(torch.float16).cpu()
        dm.train_data[i]["depth_image"] = out["depth"].permute(2, 0, 1).cpu().to(torch.float32).numpy()
        dm.train_data[i]["z_0_image"] = torch.randn((1, 4, HW_LAT, HW_LAT), generator=g).numpy()
    raster_ms = sorted(raster_ms[1:]) if len(raster_ms) > 1 else raster_ms
    from gaussctrl_b200 import gsplat_ops as _go
    n_isect = int(_go.LAST_M[0])
    # algorithmic bytes of one eval render: 244 N + 48 N_v + 152 M + 5.2 MB, N_v <= N
    raster_bytes = 244.0 * args.gaussians + 48.0 * args.gaussians + 152.0 * n_isect + 5.2e6

    # ---- device-resident inputs for `value`
    z_dev = torch.from_numpy(np.concatenate([d["z_0_image"] for d in dm.train_data])).to(dev, torch.float16)
    dep_dev = torch.from_numpy(np.concatenate([d["depth_image"] for d in dm.train_data])).to(dev)
    disparity = ops.nhwc_to_nchw(ops.depth_to_disparity(dep_dev.contiguous(), False))
    emb = pipe.prompt_encoder([pipe.negative_prompts, pipe.positive_prompt])
    neg, pos = emb[0:1], emb[1:2]

    dist_ctx, view_ids, mine = None, None, list(range(V))
    if world > 1:
        from gaussctrl_b200 import parallel as par
        dist_ctx = {"world": world, "rank": rank, "gather": par.KVAllGather()}
        view_ids = par.shard_views(V, world, rank, pipe.ref_indices)
        mine = sorted(view_ids + (list(pipe.ref_indices) if rank == 0 else []))

    def device_step():
        lat = pipe.engine.edit_refs_once(z_dev, disparity, pipe.ref_indices, pos, neg, S, GUIDANCE, view_batch=args.view_batch,
                                         view_ids=view_ids, dist_ctx=dist_ctx)
        return pipe.vae.decode_latents(lat[mine])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        l0 = ops.LAUNCHES[0]
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ops.LAUNCHES[0] - l0

    for _ in range(args.warmup):
        device_step()
    clocks = ClockSampler(local)
    clocks.start()
    ms, launches = timed(device_step, args.steps)
    clk = clocks.stop()
    ms_per_step = ms / args.steps
    value = V / (ms_per_step / 1e3)

    # ---- breakdown of one step (untimed extra): denoising loop vs VAE decode
    e0, e1, e2 = ev(), ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    lat_b = pipe.engine.edit_refs_once(z_dev, disparity, pipe.ref_indices, pos, neg, S, GUIDANCE,
                                       view_batch=args.view_batch, view_ids=view_ids, dist_ctx=dist_ctx)
    e1.record()
    pipe.vae.decode_latents(lat_b[mine])
    e2.record()
    torch.cuda.synchronize()
    breakdown = {"denoise_ms": e0.elapsed_time(e1), "vae_decode_ms": e1.elapsed_time(e2)}

    # ---- e2e through the public API: host train_data -> edit_images() -> host images
    e2e = None
    if not args.no_e2e:
        for _ in range(min(args.warmup, 1)):
            pipe.edit_images()
        ms_e, _ = timed(pipe.edit_images, args.steps)
        e2e = {"value": V / (ms_e / args.steps / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes)}

    # ---- stage A through the public API (extra, measured once): render_reverse() = rasterise + VAE encode + DDIM
    #      inversion of every view, results written to host train_data (gc_pipeline.py:122-157)
    stage_a_ms = None
    if not args.no_e2e and world == 1:
        pipe.render_reverse()  # warm-up: captures the inversion graphs
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pipe.render_reverse()
        torch.cuda.synchronize()
        stage_a_ms = (time.perf_counter() - t0) * 1e3

    # ---- the step after the path (extra): 3DGS fine-tune iterations on the edited images
    #      (gc_trainer.py:257-301: training render -> L1+SSIM -> backward -> Adam over 59 floats/Gaussian)
    finetune = None
    if not args.no_e2e and world == 1:
        try:
            import random as _random
            from gaussctrl_b200.finetune import FineTuner
            dm.device = dev
            _random.seed(0)
            tuner = FineTuner(model, dm)
            for it in range(3):
                tuner.train_iteration(30000 + it)
            n_it = 20
            e0, e1 = ev(), ev()
            torch.cuda.synchronize()
            l0 = ops.LAUNCHES[0]
            e0.record()
            for it in range(n_it):
                tuner.train_iteration(30003 + it)
            e1.record()
            torch.cuda.synchronize()
            ms_it = e0.elapsed_time(e1) / n_it
            finetune = {"iterations_per_s": 1e3 / ms_it, "ms_per_iteration": ms_it, "iterations": n_it,
                        "launches_per_iteration": (ops.LAUNCHES[0] - l0) / n_it,
                        "adam_algorithmic_bytes": 28.0 * 59 * args.gaussians,
                        "note": "GaussCtrlModel.get_outputs (training) + get_loss_dict (fused L1+SSIM fwd/bwd) + "
                                "backward + FusedAdam, one random view per iteration, reference lrs"}
        except Exception as exc:  # an extra must never take the headline down
            finetune = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- dominant kernel: cross-view attention at (N=4096, d=40), 5 sources (self + 4 cached refs)
    roof = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        # the launch shape of the view batches: the engine splits the non-reference views of this rank into the fewest
        # equal batches of at most --view-batch views (engine.edit_refs_once)
        n_nonref = len([v for v in (view_ids if view_ids is not None else range(V)) if v not in set(pipe.ref_indices)])
        n_b = max(1, -(-n_nonref // max(1, args.view_batch)))
        vb_eff = max(1, -(-n_nonref // n_b))
        Bq, C = 2 * vb_eff, ATTN_C
        qkv = torch.randn((Bq, ATTN_N, 3 * C), device=dev).half()
        refkv = torch.randn((2 * REFS, ATTN_N, 3 * C), device=dev).half()
        vb = vb_eff
        rows = [[h * vb + f] + [-(h * REFS + r) - 1 for r in range(4)] for h in range(2) for f in range(vb)]
        idx = torch.tensor(rows, dtype=torch.int32, device=dev)
        w = [0.6, 0.1, 0.1, 0.1, 0.1]
        call = lambda: ops.attention(qkv, 0, 3 * C, qkv, C, 2 * C, 3 * C, refkv, C, 2 * C, 3 * C, Bq, ATTN_N, ATTN_N,  # noqa: E731
                                     ATTN_HEADS, ATTN_D, idx, w)
        for _ in range(3):
            call()
        e0, e1 = ev(), ev()
        torch.cuda.synchronize()
        e0.record()
        reps = 20
        for _ in range(reps):
            call()
        e1.record()
        torch.cuda.synchronize()
        t_k = e0.elapsed_time(e1) / reps / 1e3
        flops = Bq * 5 * 4.0 * ATTN_N * ATTN_N * C
        ach = flops / t_k / 1e12
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture of the same shape
            tj = json.load(open(os.path.join(REPO, "profiles", "attn_ncu_traffic.json")))
            ent = tj.get("by_rows", {}).get(str(Bq))   # captures are per launch shape (B = CFG rows of a view batch)
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "multi-source cross-view attention N=4096 d=40 (5 K/V sources)",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "algorithmic_bytes": float(Bq * ATTN_N * C * 2 * 4 + 2 * REFS * ATTN_N * 2 * C * 2),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback",
                "how": f"CUDA events around {reps} back-to-back launches at the workload's shape (B={Bq} rows)"}

    # ---- CPU baseline (oracle port) on a bounded sample
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import sd15
        un, cn, _ = sd15.seeded_models(seed=0, with_vae=False)
        t = reference_step_cpu((un, cn), chunk=1)
        cpu = {"value": 1 / (t * S_STEPS), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 DDIM step of one reference-schedule chunk (R={REFS} refs + c=1 view, CFG batch 10, literal "
                         f"5-pass attention, fp32 oracle): {t:.1f} s; views/s = c / (S x t_step), S={S_STEPS}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": workload_config(args),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "extra": {"raster_ms_per_view_median": raster_ms[len(raster_ms) // 2] if raster_ms else None,
                          "raster": {"intersections": n_isect, "algorithmic_bytes": raster_bytes,
                                     "achieved_gbs": raster_bytes / (raster_ms[len(raster_ms) // 2] / 1e3) / 1e9
                                     if raster_ms else None,
                                     "note": "CUDA events around GaussCtrlModel.get_outputs_for_camera (host syncs of "
                                             "the binning included); roofline bound = HBM"},
                          "schedule": "refs_once (reference views denoised once per DDIM step, K/V recorded)",
                          "view_batch": args.view_batch, "views_per_launch": vb_eff if rank == 0 else None,
                          "breakdown": breakdown,
                          "render_reverse_ms": stage_a_ms, "finetune": finetune,
                          "views_per_s_stage_a_plus_b": (V / ((stage_a_ms + ms_per_step) / 1e3)) if stage_a_ms else None,
                          "views_total": V,
                          "multi_gpu": None if world == 1 else "views sharded round-robin; reference pass sharded over "
                                       "its 2R CFG rows with a per-layer NCCL all-gather of q|k|v"}}
        print(json.dumps(line))
    if world > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
