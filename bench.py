#!/usr/bin/env python
"""bench.py - edited-views/sec of the GaussCtrl hot path (BASELINE.json: 512^2, 20 DDIM steps, chunk=3, R=4) on N B200s.

One "step" = one full pass of `edit_images` over the workload's V views (cross-view-attention DDIM sampling through
ControlNet+UNet for every view, then VAE decode): `value` = V / step time with z_T / depth already in HBM;
`e2e` = the same pass through GaussCtrlPipeline.edit_images() with host (numpy) train_data in and host images out.
Workloads: --config cfg2 (default: the configuration the metric is quoted on) | cfg3 | cfg4 | cfg5 = BASELINE.json
configs[1..4].  At N > 1 the default is STRONG scaling (the config's V views shared by the N GPUs, as the metric says);
--scaling weak gives every GPU the config's V views.
`--impl reference` times the reference's algorithm (oracle restatement: un-fused 5-pass attention, refs recomputed per
chunk, fp32) on the host CPU cores at the metric's chunk size.  See DESIGN.md §6 for what each key means."""
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
HW_LAT, HW_IMG = 64, 512
ATTN_N, ATTN_C, ATTN_D, ATTN_HEADS = 4096, 320, 40, 8
# BASELINE.json configs[1..4]: views, Gaussians, chunk_size, ref_view_num, DDIM steps, guidance, mask, stage A
CONFIGS = {
    "cfg2": dict(views=40, gaussians=1_000_000, chunk=3, refs=4, steps=20, guidance=5.0, mask=False, inversion=False,
                 name="bear-like scene (BASELINE.json configs[1])"),
    "cfg3": dict(views=80, gaussians=3_000_000, chunk=8, refs=4, steps=20, guidance=5.0, mask=False, inversion=False,
                 name="garden-like scene (BASELINE.json configs[2])"),
    "cfg4": dict(views=128, gaussians=2_000_000, chunk=8, refs=8, steps=20, guidance=5.0, mask=False, inversion=True,
                 name="synthetic scene, DDIM inversion + edit (BASELINE.json configs[3])"),
    "cfg5": dict(views=40, gaussians=1_000_000, chunk=4, refs=4, steps=50, guidance=7.5, mask=True, inversion=False,
                 name="face-like scene with object mask (BASELINE.json configs[4])"),
}
CPU_BUDGET_S = 150.0   # the CPU reference arm executes at most this much work per invocation (see run_reference)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--views", type=int, default=None, help="override the config's view count (per GPU when --scaling weak)")
    ap.add_argument("--gaussians", type=int, default=None)
    ap.add_argument("--ddim-steps", type=int, default=None)
    ap.add_argument("--view-batch", type=int, default=40,
                    help="views denoised per launch in the refs-once schedule (results do not depend on it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip stage A / fine-tune / eager-reference extras")
    a = ap.parse_args()
    cfg = dict(CONFIGS[a.config])
    if a.views is not None:
        cfg["views"] = a.views
    if a.gaussians is not None:
        cfg["gaussians"] = a.gaussians
    if a.ddim_steps is not None:
        cfg["steps"] = a.ddim_steps
    a.cfg = cfg
    return a


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


def disc_mask(radius: int = 160) -> np.ndarray:
    yy, xx = np.mgrid[0:HW_IMG, 0:HW_IMG]
    return ((yy - HW_IMG / 2) ** 2 + (xx - HW_IMG / 2) ** 2 < radius ** 2).astype(np.int64)


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


def workload_config(args, world: int, V: int):
    c = args.cfg
    return {"workload": f"{c['name']}: {c['gaussians']} Gaussians, {V} views at 512x512 over {world} GPU(s), "
                        f"chunk_size={c['chunk']}, ref_view_num={c['refs']}, {c['steps']} DDIM steps, guidance {c['guidance']}"
                        f"{', disc object mask (composite epilogue)' if c['mask'] else ''}"
                        f"{', DDIM inversion (render_reverse) inside the timed step' if c['inversion'] else ''}; "
                        f"seeded random-init SD1.5 UNet + ControlNet-depth + VAE",
            "config": args.config, "views_total": V, "views_per_gpu": V / world, "scaling": args.scaling if world > 1 else "n/a",
            "gaussians": c["gaussians"], "chunk_size": c["chunk"], "ref_view_num": c["refs"], "ddim_steps": c["steps"],
            "guidance_scale": c["guidance"], "mask": c["mask"], "inversion_in_step": c["inversion"],
            "l2_policy": "inputs larger than L2 (activations of one denoise step ~ 6 GB > 126 MB)"}


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def reference_inputs(cfg, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    F = cfg["refs"] + cfg["chunk"]
    lat = torch.randn((F, 4, HW_LAT, HW_LAT), generator=g)
    disp = torch.rand((F, 1, HW_IMG, HW_IMG), generator=g).repeat(1, 3, 1, 1)
    pos, neg = torch.randn((1, 77, 768), generator=g), torch.randn((1, 77, 768), generator=g)
    return lat, disp, pos, neg


def reference_step(models, cfg, device="cpu", dtype=torch.float32, n_steps: int = 1):
    """`n_steps` DDIM steps of ONE reference-schedule chunk at the config's own size (R refs + c views, CFG batch
    2(R+c), literal 5-pass attention; gc_pipeline.py:206-219) with the oracle.  Returns seconds per DDIM step."""
    from oracle import pipeline as opipe, sd15
    unet, cnet = models
    lat, disp, pos, neg = (t.to(device=device, dtype=dtype) for t in reference_inputs(cfg))
    sync = torch.cuda.synchronize if str(device).startswith("cuda") else (lambda: None)
    sync()
    t0 = time.perf_counter()
    opipe.edit_chunk(unet, cnet, None, sd15.DDIMTables(), lat, disp, pos, neg, cfg["steps"], cfg["guidance"], cfg["refs"],
                     decode=False, stop_after=n_steps)
    sync()
    return (time.perf_counter() - t0) / n_steps


def reference_views_per_s(cfg, t_step: float, t_decode_per_image: float = 0.0) -> float:
    """A chunk edits c views in S DDIM steps of the (R+c)-frame batch, then decodes all R+c frames and drops the R
    reference images (gc_pipeline.py:209-219)."""
    return cfg["chunk"] / (cfg["steps"] * t_step + (cfg["refs"] + cfg["chunk"]) * t_decode_per_image)


def run_reference(args):
    """`--impl reference`: the oracle port on the host cores, at the metric's own chunk (c views + R references)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import sd15
    cfg = args.cfg
    unet, cnet, vae = sd15.seeded_models(seed=0, with_vae=True)
    cores = torch.get_num_threads()
    # one "step" of this arm = one DDIM step of one chunk (~25 TFLOP at c=3: tens of seconds on the host cores).  The
    # requested warm-up / step counts are honoured up to CPU_BUDGET_S of work; fewer steps are executed beyond that
    # (never a smaller chunk) and the executed counts are stated in `sample`.
    t_first = reference_step((unet, cnet), cfg)
    n_warm = 1
    budget_steps = max(1, int((CPU_BUDGET_S - t_first) // max(t_first, 1e-3)))
    n_timed = max(1, min(args.steps, budget_steps))
    times = [reference_step((unet, cnet), cfg) for _ in range(n_timed)]
    t = sum(times) / len(times)
    # VAE decode of one image (the reference decodes all R+c frames of a chunk)
    with torch.no_grad():
        z = torch.randn((1, 4, HW_LAT, HW_LAT), generator=torch.Generator().manual_seed(3))
        t0 = time.perf_counter()
        vae.decode(z / vae.scaling_factor)
        t_dec = time.perf_counter() - t0
    vps = reference_views_per_s(cfg, t, t_dec)
    F = cfg["refs"] + cfg["chunk"]
    sample = (f"{n_timed} timed DDIM step(s) after {n_warm} warm-up step of ONE reference-schedule chunk at the metric's own "
              f"size (R={cfg['refs']} refs + c={cfg['chunk']} views, CFG batch {2 * F}, literal 5-pass attention, fp32 oracle "
              f"port): {t:.1f} s/step; + VAE decode {t_dec:.1f} s/image x {F} frames per chunk; views/s = c / (S x t_step + "
              f"F x t_decode), S={cfg['steps']}; rasterisation not included (<0.1 % of the path)")
    V = cfg["views"]
    line = {"impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.scaling == "strong" else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, max(1, args.gpus), V if args.scaling == "strong" else V * max(1, args.gpus)),
            "cpu_baseline": {"value": vps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- B200 arm
def multi_gpu_selfcheck(dev, world, rank):
    """Sharded edit (views round-robin, reference pass sharded, K/V exchange over peer memory) == single-GPU edit,
    bit for bit, on a small problem (32x32 latents, 2 DDIM steps, 9 views).  Runs on every rank; rank 0 reports."""
    import torch.distributed as dist
    from gaussctrl_b200 import parallel as par
    from gaussctrl_b200.diffusion import SD15Denoiser
    from gaussctrl_b200.engine import EditEngine

    def check(den, gather):
        g = torch.Generator().manual_seed(7)
        V, S, hw = 9, 2, 32
        lat = torch.randn((V, 4, hw, hw), generator=g).half()
        disp = torch.rand((V, 1, hw * 8, hw * 8), generator=g).repeat(1, 3, 1, 1).half()
        pos, neg = torch.randn((1, 77, 768), generator=g), torch.randn((1, 77, 768), generator=g)
        ref_idx = [0, 3, 5, 8]
        ctx = {"world": world, "rank": rank, "gather": gather}
        mine = par.shard_views(V, world, rank, ref_idx)
        out = EditEngine(den).edit_refs_once(lat, disp, ref_idx, pos, neg, S, 5.0, view_batch=2, view_ids=mine, dist_ctx=ctx)
        full = par.gather_view_results(out[mine].contiguous(), mine, V, world)
        for ri in ref_idx:
            full[ri] = out[ri]
        want = EditEngine(den).edit_refs_once(lat, disp, ref_idx, pos, neg, S, 5.0, view_batch=2)
        diff = torch.tensor([(full.float() - want.float()).abs().max().item()], device=dev)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        return float(diff.item())
    return check


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    cfg = args.cfg
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from gaussctrl_b200 import gsplat_ops as _go, ops, parallel as par
    from gaussctrl_b200._compat import Cameras
    from gaussctrl_b200.gc_model import GaussCtrlModel, GaussCtrlModelConfig
    from gaussctrl_b200.gc_pipeline import GaussCtrlPipeline, GaussCtrlPipelineConfig, SimpleDataManager
    from gaussctrl_b200.sd15_spec import synthetic_weights

    S, R, GUIDANCE = cfg["steps"], cfg["refs"], cfg["guidance"]
    V = cfg["views"] * (world if args.scaling == "weak" else 1)
    n_gauss = cfg["gaussians"]
    # ---- scene + cameras + pipeline (public API objects)
    scene = synthetic_scene(n_gauss, seed=0)
    model = GaussCtrlModel(GaussCtrlModelConfig(), num_points=n_gauss)
    with torch.no_grad():
        for k, v in scene.items():
            getattr(model, k).data = v
    model = model.to(dev)
    model.background_color = torch.zeros(3)
    c2w = torch.stack([orbit_c2w(i, V) for i in range(V)])
    cams = Cameras(c2w, 539.05, 538.17, 258.74, 239.35, HW_IMG, HW_IMG)
    dm = SimpleDataManager(cams)
    pcfg = GaussCtrlPipelineConfig(edit_prompt="a photo of a polar bear in the forest",
                                   reverse_prompt="a photo of a bear statue in the forest", guidance_scale=GUIDANCE,
                                   num_inference_steps=S, chunk_size=cfg["chunk"], ref_view_num=R,
                                   langsam_obj="bear" if cfg["mask"] else "", diffusion_ckpt="synthetic")
    mask = disc_mask() if cfg["mask"] else None
    pipe = GaussCtrlPipeline(pcfg, dev, world_size=world, local_rank=local, datamanager=dm, model=model,
                             weights=synthetic_weights(0), mask_fn=(lambda rgb, text: mask) if cfg["mask"] else None)
    pipe.view_batch = args.view_batch
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

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

    # ---- multi-GPU parity self-check (the 2-GPU pytest cannot run on the driver's 1-GPU test box)
    mg_check = None
    if world > 1:
        pipe._kv_gather = par.make_kv_gather(dev)
        try:
            mg_check = {"max_abs_diff_sharded_vs_single": multi_gpu_selfcheck(dev, world, rank)(pipe.denoiser, pipe._kv_gather),
                        "exchange": type(pipe._kv_gather).__name__, "what": "9 views, 2 DDIM steps, 32x32 latents"}
        except Exception as exc:
            mg_check = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- stage-A products (setup): rasterise rgb + depth for the ControlNet condition through the public model API
    #      (views dealt to the ranks); z_T ~ N(0,1) unless the config runs the inversion
    g = torch.Generator().manual_seed(1)
    mine_a = list(range(rank, V, world))
    torch.cuda.synchronize()
    for _ in range(2):                    # warm-up: per-stream workspaces, allocator pools of the side streams
        outs = pipe.render_views(mine_a)
        _go.check_deferred_overflow()
        del outs
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    outs = pipe.render_views(mine_a)
    e1.record()
    torch.cuda.synchronize()
    _go.check_deferred_overflow()
    raster_ms_per_view = e0.elapsed_time(e1) / max(1, len(mine_a))
    n_isect = int(_go.LAST_M[0])
    rgb_l = torch.stack([o["rgb"].to(torch.float16) for o in outs])
    dep_l = torch.stack([o["depth"].to(torch.float16) for o in outs])
    if world > 1:
        rgb_l = par.gather_view_results(rgb_l, mine_a, V, world)
        dep_l = par.gather_view_results(dep_l, mine_a, V, world)
    rgb_h, dep_h = rgb_l.cpu(), dep_l.permute(0, 3, 1, 2).float().cpu()
    for i in range(V):
        dm.train_data[i]["unedited_image"] = rgb_h[i].clone()
        dm.train_data[i]["depth_image"] = dep_h[i].numpy().copy()
        dm.train_data[i]["z_0_image"] = torch.randn((1, 4, HW_LAT, HW_LAT), generator=g).numpy()
        if mask is not None:
            dm.train_data[i]["mask_image"] = mask
    del outs, rgb_l, dep_l
    # algorithmic bytes of one eval render: 244 N + 48 N_v + 152 M + 5.2 MB, N_v <= N
    raster_bytes = 244.0 * n_gauss + 48.0 * n_gauss + 152.0 * n_isect + 5.2e6

    # ---- device-resident inputs for `value`: exactly what edit_images uploads, already in HBM
    plan = pipe.edit_plan(V)
    need, mine = plan["need"], plan["mine"]
    z_dev = torch.from_numpy(np.concatenate([dm.train_data[i]["z_0_image"] for i in need])).to(dev, torch.float16)
    dep_dev = torch.from_numpy(np.concatenate([dm.train_data[i]["depth_image"] for i in need])).to(dev)
    masks_dev = uned_dev = None
    if mask is not None:
        masks_dev = torch.from_numpy(np.stack([mask.astype(np.float32) for _ in mine])).to(dev)
        uned_dev = torch.stack([dm.train_data[i]["unedited_image"] for i in mine]).to(dev, torch.float16)

    def device_step():
        if cfg["inversion"]:
            pipe.render_reverse()        # cfg4: stage A is part of the measured path (results land in train_data)
        return pipe.edit_on_device(z_dev, dep_dev, plan, masks_dev, uned_dev)

    for _ in range(args.warmup):
        device_step()
    clocks = ClockSampler(local)
    clocks.start()
    ms, launches = timed(device_step, args.steps)
    clk = clocks.stop()
    ms_per_step = ms / args.steps
    value = V / (ms_per_step / 1e3)

    # ---- breakdown of one step (untimed extra): denoising loop vs VAE decode
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    pipe.edit_on_device(z_dev, dep_dev, plan, masks_dev, uned_dev)
    e1.record()
    torch.cuda.synchronize()
    e2, e3 = ev(), ev()
    lat_probe = torch.randn((len(mine), 4, HW_LAT, HW_LAT), device=dev).half()
    e2.record()
    pipe.vae.decode_latents(lat_probe, masks_dev, uned_dev)
    e3.record()
    torch.cuda.synchronize()
    breakdown = {"edit_on_device_ms": e0.elapsed_time(e1), "vae_decode_ms": e2.elapsed_time(e3),
                 "denoise_ms": e0.elapsed_time(e1) - e2.elapsed_time(e3), "views_decoded_on_this_rank": len(mine)}

    # ---- e2e through the public API: host train_data -> edit_images() -> host images
    e2e = None
    if not args.no_e2e:
        def public_step():
            if cfg["inversion"]:
                pipe.render_reverse()
            pipe.edit_images()
        for _ in range(min(args.warmup, 1)):
            public_step()
        ms_e, _ = timed(public_step, args.steps)
        e2e = {"value": V / (ms_e / args.steps / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": int(pipe.h2d_bytes), "d2h_bytes_per_step": int(pipe.d2h_bytes),
               "note": "per rank: uploads its own views + the references; downloads all V edited images (all-gathered)"
               if world > 1 else None}

    extras_on = not (args.no_extras or args.no_e2e)
    # ---- stage A through the public API (extra): render_reverse() = rasterise + VAE encode + DDIM inversion of every
    #      view (dealt to the ranks), results gathered into every rank's host train_data (gc_pipeline.py:122-157)
    stage_a_ms = None
    if extras_on:
        pipe.render_reverse()  # warm-up: captures the inversion graphs
        ms_a, _ = timed(pipe.render_reverse, 1)
        stage_a_ms = ms_a

    # ---- the step after the path (extra): 3DGS fine-tune iterations on the edited images
    finetune = None
    if extras_on and world == 1:
        try:
            import random as _random
            from gaussctrl_b200.finetune import FineTuner
            if "image" not in dm.train_data[0]:
                pipe.edit_images()
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
                        "adam_algorithmic_bytes": 28.0 * 59 * n_gauss}
        except Exception as exc:  # an extra must never take the headline down
            finetune = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- dominant kernel: cross-view attention at (N=4096, d=40), 5 sources (self + 4 cached refs)
    roof = roof_iso = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_sus = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_burst = float(peaks.get("bf16_tflops", 1650.0))
        n_nonref = len([v for v in (plan["view_ids"] if plan["view_ids"] is not None else range(V))
                        if v not in set(pipe.ref_indices)])
        n_b = max(1, -(-n_nonref // max(1, args.view_batch)))
        vb_eff = max(1, -(-n_nonref // n_b))
        Bq, C = 2 * vb_eff, ATTN_C
        traffic, traffic_src = None, None
        try:  # DRAM bytes per launch of this kernel from the committed `ncu --set full` capture of the same shape
            tj = json.load(open(os.path.join(REPO, "profiles", "attn_ncu_traffic.json")))
            ent = tj.get("by_rows", {}).get(str(Bq))
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            pass
        alg_bytes = float(Bq * ATTN_N * C * 2 * 4 + 2 * R * ATTN_N * 2 * C * 2)
        # (1) IN SITU: the engine's own view-batch denoise steps run eagerly (no CUDA graph) with CUDA events around
        #     every attention launch on the launching stream; three consecutive full ControlNet+UNet evaluations, the
        #     (N=4096, d=40, 5-source) UNet launches of the last two are averaged -> power / clock state of the real step
        try:
            from gaussctrl_b200.diffusion import cached_crossview_plan
            from gaussctrl_b200.gc_pipeline import crossview_ref_frames
            key = next(k for k in pipe.engine._steps if k[0] == "refs_once")
            view_step = pipe.engine._steps[key][2][0]
            emb_e = pipe.prompt_encoder([pipe.negative_prompts, pipe.positive_prompt])
            pipe.denoiser.set_prompts(torch.cat([emb_e[0:1], emb_e[1:2]], dim=0))   # stage A left its single prompt set
            ops.ATTN_EVENTS = []
            marks = []
            for _ in range(3):
                marks.append(len(ops.ATTN_EVENTS))
                view_step._body()
            torch.cuda.synchronize()
            evs = [e for e in ops.ATTN_EVENTS[marks[1]:] if e[1] == ATTN_N and e[4] == ATTN_D and e[5] == 5]
            ops.ATTN_EVENTS = None
            t_k = sum(e[6].elapsed_time(e[7]) for e in evs) / len(evs) / 1e3
            rows = evs[0][0]
            flops = rows * 5 * 4.0 * ATTN_N * ATTN_N * C
            ach = flops / t_k / 1e12
            roof = {"bound": "tensor", "kernel": "multi-source cross-view attention N=4096 d=40 (5 K/V sources)",
                    "achieved": ach, "peak": peak_sus, "unit": "TFLOP/s", "frac": ach / peak_sus, "traffic": traffic,
                    "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes, "rows_per_launch": rows,
                    "launches_averaged": len(evs), "avg_launch_ms": t_k * 1e3,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                    if peaks else "fallback",
                    "how": "in situ: CUDA events on the launching stream around each attention launch of three consecutive "
                           "eager ControlNet+UNet evaluations of the view batch (UNet launches of the last two averaged)"}
        except Exception as exc:
            ops.ATTN_EVENTS = None
            roof = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        # (2) ISOLATED: the same launch alone, back to back, against the burst peak
        # the layout SD15Denoiser gives the kernel: fused q|k|v rows whose V heads are padded to 48 columns, column 40 = 1
        vs = 48 if getattr(pipe.denoiser, "ones_column", False) else ATTN_D
        ld = 2 * C + ATTN_HEADS * vs
        qkv = torch.randn((Bq, ATTN_N, ld), device=dev).half()
        refkv = torch.randn((2 * R, ATTN_N, ld), device=dev).half()
        if vs > ATTN_D:
            for t_ in (qkv, refkv):
                vv = t_[..., 2 * C:].reshape(t_.shape[0], ATTN_N, ATTN_HEADS, vs)
                vv[..., ATTN_D:] = 0
                vv[..., ATTN_D] = 1.0
        rows_i = [[h * vb_eff + f] + [-(h * R + r) - 1 for r in range(4)] for h in range(2) for f in range(vb_eff)]
        idx = torch.tensor(rows_i, dtype=torch.int32, device=dev)
        w = [0.6, 0.1, 0.1, 0.1, 0.1]
        call = lambda: ops.attention(qkv, 0, ld, qkv, C, 2 * C, ld, refkv, C, 2 * C, ld, Bq, ATTN_N, ATTN_N,  # noqa: E731
                                     ATTN_HEADS, ATTN_D, idx, w, v_head_stride=vs)
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
        t_i = e0.elapsed_time(e1) / reps / 1e3
        ach_i = Bq * 5 * 4.0 * ATTN_N * ATTN_N * C / t_i / 1e12
        roof_iso = {"achieved": ach_i, "peak": peak_burst, "unit": "TFLOP/s", "frac": ach_i / peak_burst,
                    "rows_per_launch": Bq, "avg_launch_ms": t_i * 1e3,
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)",
                    "how": f"CUDA events around {reps} back-to-back launches"}
        del qkv, refkv

    # ---- CPU baseline (oracle port, metric's own chunk) and the reference schedule in eager fp16 on this GPU
    cpu = ref_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import sd15
        un, cn, vae_o = sd15.seeded_models(seed=0, with_vae=True)
        t = reference_step((un, cn), cfg)
        F = cfg["refs"] + cfg["chunk"]
        cpu = {"value": reference_views_per_s(cfg, t), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"1 DDIM step of ONE reference-schedule chunk at the metric's own size (R={cfg['refs']} refs + "
                         f"c={cfg['chunk']} views, CFG batch {2 * F}, literal 5-pass attention, fp32 oracle port): {t:.1f} s; "
                         f"views/s = c / (S x t_step), S={S}; VAE decode and rasterisation not in this sample "
                         f"(`--impl reference` adds the decode)"}
        if not args.no_extras:
            try:
                # the reference SCHEDULE (refs recomputed per chunk, 5 un-fused attention passes
                # with [B*8,N,N] probabilities in HBM, every frame decoded) in eager fp16 PyTorch (cuBLAS/cuDNN) on this
                # same B200 - the denominator of the north star's ">= 10x the reference GPU pipeline"
                un, cn, vae_o = un.half().to(dev), cn.half().to(dev), vae_o.half().to(dev)
                reference_step((un, cn), cfg, dev, torch.float16, 1)
                t_g = reference_step((un, cn), cfg, dev, torch.float16, 2)
                with torch.no_grad():
                    z = torch.randn((F, 4, HW_LAT, HW_LAT), device=dev).half()
                    vae_o.decode(z / vae_o.scaling_factor)
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    vae_o.decode(z / vae_o.scaling_factor)
                    torch.cuda.synchronize()
                    t_dec = (time.perf_counter() - t0) / F
                ref_gpu = {"value": reference_views_per_s(cfg, t_g, t_dec), "unit": UNIT, "s_per_ddim_step": t_g,
                           "s_decode_per_image": t_dec,
                           "sample": f"oracle (literal reference schedule) in eager fp16 on this GPU: 2 timed DDIM steps of one "
                                     f"chunk (R={cfg['refs']}+c={cfg['chunk']} frames, CFG batch {2 * F}) after 1 warm-up, "
                                     f"+ VAE decode of the {F} frames; views/s = c / (S x t_step + F x t_decode)"}
                del un, cn, vae_o
            except Exception as exc:
                ref_gpu = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak" if (world > 1 and args.scaling == "weak") else "strong",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": workload_config(args, world, V),
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                "extra": {"roofline_isolated": roof_iso,
                          "reference_gpu_eager": ref_gpu,
                          "speedup_vs_reference_gpu_eager": (e2e["value"] / ref_gpu["value"])
                          if (e2e and ref_gpu and "value" in ref_gpu) else None,
                          "raster": {"ms_per_view": raster_ms_per_view, "views_rendered_on_this_rank": len(mine_a),
                                     "intersections": n_isect, "algorithmic_bytes": raster_bytes,
                                     "achieved_gbs": raster_bytes / (raster_ms_per_view / 1e3) / 1e9,
                                     "frac_of_hbm_peak": raster_bytes / (raster_ms_per_view / 1e3) / 1e9 /
                                     float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6552.6))
                                     if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else None,
                                     "note": "CUDA events around GaussCtrlPipeline.render_views of this rank's views "
                                             "(GaussCtrlModel.get_outputs_for_cameras: one gcb_render_eval_batch call per "
                                             "stream, 4 streams, no host synchronisation inside); roofline bound = HBM"},
                          "schedule": "refs_once (reference views denoised once per DDIM step, K/V recorded)",
                          "view_batch": args.view_batch, "breakdown": breakdown,
                          "render_reverse_ms": stage_a_ms, "finetune": finetune,
                          "views_per_s_stage_a_plus_b": (V / ((stage_a_ms + (0 if cfg["inversion"] else ms_per_step)) / 1e3))
                          if stage_a_ms else None,
                          "views_total": V, "multi_gpu_check": mg_check,
                          "multi_gpu": None if world == 1 else
                          "views dealt round-robin (stage A and edit); reference pass sharded over its 2R CFG rows with a "
                          "per-layer exchange of q|k|v (" + type(pipe._kv_gather).__name__ + ")"}}
        print(json.dumps(line))
    if world > 1:
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
