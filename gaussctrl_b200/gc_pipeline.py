"""GaussCtrlPipeline on the B200 kernels: host-side mirror of gaussctrl/gc_pipeline.py:48-291.

Same config fields and defaults, same constructor signature, same methods (`render_reverse`, `edit_images`,
`image2latent`, `depth2disparity`, `depth2disparity_torch`, `update_datasets`) and the same `train_data[i]` dict schema
(`image`, `image_idx`, `unedited_image`, `depth_image`, `z_0_image`, `mask_image`; dtypes/shapes of
gc_pipeline.py:268-274 and :234): the calls the reference's trainer makes (`gc_trainer.py:75-78`, `:272`) are all here.
The constructor builds datamanager and model from `config.datamanager` / `config.model` exactly like nerfstudio's
VanillaPipeline (that contract is exercised against a fake `nerfstudio` tree in tests/test_plugin_seam_cpu.py; the
real package is not installable in this image).
What changed is below the seam: rasterisation, VAE, ControlNet+UNet, cross-view attention and the DDIM updates are
the sm_100a kernels of this package, views are batched, and the reference views are denoised once per step instead
of once per chunk (engine.EditEngine.edit_refs_once)."""
from __future__ import annotations

import hashlib
import os
import random
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence, Type

import numpy as np
import torch

from . import ops
from ._compat import HAVE_NERFSTUDIO, VanillaPipeline, VanillaPipelineConfig
from .gc_datamanager import GaussCtrlDataManagerConfig
from .diffusion import SD15Denoiser
from .engine import EditEngine
from .sd15_spec import DDIMTables, synthetic_weights
from .vae import VaeB200

ADDED_PROMPT = "best quality, extremely detailed"
NEGATIVE_PROMPT = ("longbody, lowres, bad anatomy, bad hands, missing fingers, extra digit, fewer digits, cropped, "
                   "worst quality, low quality")


def select_ref_indices(view_num: int, ref_view_num: int) -> List[int]:
    """gc_pipeline.py:109-114 with the same seed; `random.randint` is inclusive so the reference can return
    `view_num` (out of range, SURVEY §8a gotcha 3): clamped to the last view here."""
    anchors = [(view_num * i) // ref_view_num for i in range(ref_view_num)] + [view_num]
    random.seed(13789)
    idx = [random.randint(a, anchors[i + 1]) for i, a in enumerate(anchors[:-1])]
    return [min(i, view_num - 1) for i in idx]


def crossview_ref_frames(ref_view_num: int) -> tuple:
    """Which rows of the reference block serve as K/V sources of the cross-view attention.  The reference hard-codes
    frames 0,1,2,3 of each CFG half (utils.py:95-109): with R=4 those are the four reference views; with R=8 (BASELINE
    cfg4) only the first four of the eight are sources - kept; with R<4 the reference reads chunk views as if they were
    references, or raises IndexError at R=c=1 (SURVEY §8a gotcha 1) - here the K = R references that exist are used with
    weights (1-coeff)/K (documented deviation)."""
    return tuple(range(min(int(ref_view_num), 4)))


def _pinned(t: torch.Tensor) -> torch.Tensor:
    return t.pin_memory() if torch.cuda.is_available() else t


def synthetic_prompt_embeds(prompts: Sequence[str], seq: int = 77, dim: int = 768) -> torch.Tensor:
    """Stand-in for the CLIP text encoder when no checkpoint is on disk: a deterministic N(0,1) embedding per prompt
    string.  (CLIP itself is outside the hot path: it runs once per `pipe()` call in the reference.)"""
    out = []
    for p in prompts:
        seed = int.from_bytes(hashlib.sha256(p.encode()).digest()[:8], "little") % (2 ** 63)
        out.append(torch.randn((seq, dim), generator=torch.Generator().manual_seed(seed)))
    return torch.stack(out)


@dataclass
class GaussCtrlPipelineConfig(VanillaPipelineConfig):
    """Field names, types and defaults of gaussctrl/gc_pipeline.py:48-73."""
    _target: Type = field(default_factory=lambda: GaussCtrlPipeline)
    datamanager: GaussCtrlDataManagerConfig = field(default_factory=GaussCtrlDataManagerConfig)
    render_rate: int = 500
    edit_prompt: str = ""
    reverse_prompt: str = ""
    langsam_obj: str = ""
    guidance_scale: float = 5
    num_inference_steps: int = 20
    chunk_size: int = 5
    ref_view_num: int = 4
    diffusion_ckpt: str = "CompVis/stable-diffusion-v1-4"
    # --- B200 extensions (defaults keep the reference's results)
    edit_schedule: str = "refs_once"      # or "reference": refs recomputed in every chunk, as gc_pipeline.py:190-219
    view_batch: int = 40                  # views denoised per launch in the refs_once schedule; results do not depend on
                                          # it (batch-invariant kernels), throughput does (12 -> 40: +3 % on one B200)
    synthetic_seed: int = 0               # weights seed for diffusion_ckpt="synthetic" (benchmarks / tests only)


def resolve_checkpoint(ckpt: str) -> str:
    """`from_pretrained(ckpt)` semantics without network (gc_pipeline.py:97-102): a local diffusers-layout folder, or a
    hub id already present in the local Hugging Face cache.  Anything else raises - like the reference does when
    `from_pretrained` cannot resolve the checkpoint - instead of silently editing with random weights."""
    if os.path.isdir(ckpt):
        return ckpt
    try:
        from huggingface_hub import snapshot_download  # local cache only: there is no network on the GPU box
        return snapshot_download(ckpt, local_files_only=True)
    except Exception as exc:
        raise FileNotFoundError(
            f"diffusion_ckpt={ckpt!r} is neither a local diffusers checkpoint folder nor in the local Hugging Face cache "
            f"({type(exc).__name__}); pass a folder, or diffusion_ckpt='synthetic' for seeded random weights "
            f"(benchmarks / tests only)") from exc


class GaussCtrlPipeline(VanillaPipeline):
    config: GaussCtrlPipelineConfig

    def __init__(self, config: GaussCtrlPipelineConfig, device: str, test_mode: str = "val", world_size: int = 1,
                 local_rank: int = 0, grad_scaler: Optional[Any] = None, *, datamanager: Any = None, model: Any = None,
                 weights: Optional[tuple] = None, prompt_encoder: Optional[Callable] = None,
                 mask_fn: Optional[Callable] = None):
        if datamanager is not None or model is not None:
            # B200 extension (bench.py, tests): pre-built datamanager / model objects instead of building them from
            # `config.datamanager` / `config.model` - the pipeline-specific state below is identical
            torch.nn.Module.__init__(self)
            self.config, self.test_mode, self.world_size = config, test_mode, world_size
            self.datamanager, self._model = datamanager, model
        else:
            super().__init__(config, device, test_mode, world_size, local_rank)   # gc_pipeline.py:89
        self.config = config
        self.device_ = torch.device(device)
        self.test_mode = test_mode
        self.world_size, self.local_rank = world_size, local_rank
        self.mask_fn = mask_fn  # LangSAM stays external: any callable (rgb[H,W,3], text) -> mask[H,W]
        self.edit_prompt = config.edit_prompt
        self.reverse_prompt = config.reverse_prompt
        self.pipe_device = self.device_
        ckpt_dir = None
        if weights is None:
            weights, ckpt_dir = self._load_weights(config.diffusion_ckpt, config.synthetic_seed)
        unet_sd, cnet_sd, vae_sd = weights
        self.denoiser = SD15Denoiser(unet_sd, cnet_sd, self.device_)
        self.vae = VaeB200(vae_sd, self.device_) if vae_sd is not None else None
        self.tables = DDIMTables()
        self.engine = EditEngine(self.denoiser, self.tables)
        if prompt_encoder is None and ckpt_dir is not None:
            from .clip_text import make_prompt_encoder   # real tokenizer + CLIP text encoder on the B200 kernels
            prompt_encoder = make_prompt_encoder(ckpt_dir, self.device_)
            if prompt_encoder is None:
                raise FileNotFoundError(f"{ckpt_dir}: tokenizer/ or text_encoder/ is missing - real UNet weights need the "
                                        f"real prompt embeddings (the reference's from_pretrained would fail here too)")
        # synthetic prompt embeddings only ever pair with synthetic / injected weights
        self.prompt_encoder = prompt_encoder or synthetic_prompt_embeds
        self.positive_prompt = self.edit_prompt + ", " + ADDED_PROMPT
        self.positive_reverse_prompt = self.reverse_prompt + ", " + ADDED_PROMPT
        self.negative_prompts = NEGATIVE_PROMPT
        view_num = len(self.datamanager.cameras)
        self.ref_indices = select_ref_indices(view_num, config.ref_view_num)
        self.num_ref_views = len(self.ref_indices)
        self.num_inference_steps = config.num_inference_steps
        self.guidance_scale = config.guidance_scale
        self.controlnet_conditioning_scale = 1.0
        self.eta = 0.0
        self.chunk_size = config.chunk_size

    @staticmethod
    def _load_weights(ckpt: str, seed: int):
        """-> ((unet_sd, controlnet_sd, vae_sd), checkpoint folder or None).  `diffusion_ckpt="synthetic"` (optionally
        "synthetic:<seed>") asks for seeded random-init weights explicitly; every other value must resolve to a
        diffusers-layout folder (checkpoint.py) or raises."""
        if ckpt == "synthetic" or ckpt.startswith("synthetic:"):
            if ":" in ckpt:
                seed = int(ckpt.split(":", 1)[1])
            return synthetic_weights(seed), None
        folder = resolve_checkpoint(ckpt)
        from .checkpoint import load_diffusers_checkpoint
        return load_diffusers_checkpoint(folder), folder

    # ------------------------------------------------------------------------------------------ multi-GPU helpers
    def _dist(self):
        """(world, rank) of the view sharding: one process per GPU under torch.distributed, else (1, 0)."""
        if self.world_size > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
            return torch.distributed.get_world_size(), torch.distributed.get_rank()
        return 1, 0

    def _camera_at(self, idx: int):
        """Camera `idx` with a batch dimension.  It stays where the datamanager keeps it (host memory): the rasteriser
        only reads its intrinsics / pose as kernel arguments, so the reference's `.to(self.device)` (gc_pipeline.py:128)
        would just turn every `.item()` in get_outputs into a device synchronisation."""
        cam = self.datamanager.cameras[idx]
        if len(cam.shape) == 0:           # nerfstudio Cameras[int] drops the batch dimension
            cam = cam.reshape((1,))
        return cam

    # ------------------------------------------------------------------------------------------ stage A
    @torch.no_grad()
    def render_reverse(self):
        """Render rgb + depth of every view, encode, and DDIM-invert to z_T (gc_pipeline.py:122-157).
        The reference loops views at batch 1 on one GPU; here the three stages each run over a rank's views in batches,
        views are dealt round-robin to the ranks (no collective on the data path) and ONE
        all-gather at the end gives every rank the complete `train_data` (z_T 64 KB, render 2 MB per view)."""
        V = len(self.datamanager.cameras)
        world, rank = self._dist()
        mine = list(range(rank, V, world))
        want_mask = self.config.langsam_obj != ""
        if want_mask and self.mask_fn is None:
            raise ValueError(f"langsam_obj={self.config.langsam_obj!r} needs a segmentation callable: pass mask_fn=(rgb, "
                             f"text) -> [H,W] mask (LangSAM in the reference, gc_pipeline.py:147-154)")
        model = self.model
        from . import gsplat_ops
        for attempt in range(2):
            # all of this rank's views are enqueued without a host synchronisation (the intersection counts stay on the
            # device); ONE check afterwards, and a second pass only if a view outgrew the intersection capacity
            outs = self.render_views(mine)
            rgbs = [o["rgb"].to(torch.float16) for o in outs]          # [H,W,3] 0..1   (:132)
            depths = [o["depth"].to(torch.float16) for o in outs]      # [H,W,1]        (:133)
            try:
                gsplat_ops.check_deferred_overflow()
                break
            except gsplat_ops.IsectOverflow:
                if attempt == 1:
                    raise
        H, W = (rgbs[0].shape[0], rgbs[0].shape[1]) if rgbs else (0, 0)
        if mine:
            rgb = torch.stack(rgbs)
            depth = torch.stack(depths)
            eb = self.stage_a_batch
            z0 = torch.cat([self.image2latent_batch(rgb[i:i + eb]) for i in range(0, len(mine), eb)])
            disparity = ops.depth_to_disparity(depth[..., 0].float().contiguous(), True)  # depth2disparity_torch on fp16
            disparity = ops.nhwc_to_nchw(disparity)
            emb = self.prompt_encoder([self.positive_reverse_prompt])
            zT = self.engine.invert(z0, disparity, emb, self.num_inference_steps, batch=self.invert_batch)
        masks = None
        if want_mask:
            masks = [np.asarray(self.mask_fn(rgbs[j].cpu(), self.config.langsam_obj)) * 1 for j in range(len(mine))]
        if world > 1:
            from . import parallel as par
            shapes = par.broadcast_shape((H, W), self.device_)     # ranks without views still take part in the gather
            H, W = shapes
            e = lambda *shp, dt=torch.float16: torch.zeros(shp, dtype=dt, device=self.device_)  # noqa: E731
            if not mine:
                rgb, depth, zT = e(0, H, W, 3), e(0, H, W, 1), e(0, 4, H // 8, W // 8)
            rgb = par.gather_view_results(rgb, mine, V, world)
            depth = par.gather_view_results(depth, mine, V, world)
            zT = par.gather_view_results(zT, mine, V, world)
            if want_mask:
                m_loc = torch.from_numpy(np.stack(masks).astype(np.int32)).to(self.device_) if mine else \
                    e(0, H, W, dt=torch.int32)
                m_all = par.gather_view_results(m_loc, mine, V, world).cpu().numpy()
                masks = [m_all[i].astype(np.int64) for i in range(V)]
            ids = list(range(V))
        else:
            ids = mine
        # one device->host copy per product instead of the reference's three `.cpu()` syncs per view (:268-274)
        rgb_h, depth_h, z_h = rgb.cpu(), depth.permute(0, 3, 1, 2).to(torch.float32).cpu(), zT.to(torch.float32).cpu()
        for j, cam_idx in enumerate(ids):
            self._store_view(cam_idx, rgb_h[j].clone(), depth_h[j].numpy().copy(), z_h[j:j + 1].numpy().copy(),
                             None if masks is None else masks[j])

    def render_views(self, view_ids: Sequence[int]) -> List[Dict[str, torch.Tensor]]:
        """Eval renders (`model.get_outputs_for_camera`) of the given views, enqueued without any host synchronisation
        and spread over `raster_streams` CUDA streams (the binning kernels of one view leave most of the GPU idle).
        The caller runs `gsplat_ops.check_deferred_overflow()` afterwards."""
        from . import gsplat_ops
        model = self.model
        streams = None
        if self.device_.type == "cuda" and self.raster_streams > 1:
            if getattr(self, "_raster_streams", None) is None:
                self._raster_streams = [torch.cuda.Stream(self.device_) for _ in range(self.raster_streams)]
            streams = self._raster_streams
        if hasattr(model, "get_outputs_for_cameras"):
            # one batched call per stream (gcb_render_eval_batch): the per-view host loop runs in C
            return model.get_outputs_for_cameras([self._camera_at(ci) for ci in view_ids], streams=streams)
        model.defer_isect_check = True
        try:
            if streams is not None:
                return gsplat_ops.render_views_multistream(
                    lambda ci: model.get_outputs_for_camera(self._camera_at(ci)), list(view_ids), streams)
            return [model.get_outputs_for_camera(self._camera_at(ci)) for ci in view_ids]
        finally:
            model.defer_isect_check = False

    # batch sizes of stage A (B200 extension; results do not depend on them: batch-invariant kernels)
    raster_streams = 4   # concurrent eval renders (measured on B200, 1 M Gaussians: 0.51 -> 0.36 ms per view)
    stage_a_batch = 8    # views per VAE-encode launch set (512^2 x 128-channel activations: 67 MB per view and layer)
    invert_batch = 40    # views per DDIM-inversion launch set (the edit stage's view batch)

    def _store_view(self, cam_idx, unedited_image, depth_np, latent_np, mask):
        td = self.datamanager.train_data[cam_idx]
        td["unedited_image"] = unedited_image
        td["depth_image"] = depth_np
        td["z_0_image"] = latent_np
        if mask is not None:
            td["mask_image"] = mask

    # ------------------------------------------------------------------------------------------ stage B
    @torch.no_grad()
    def edit_images(self):
        """Edit every view with ControlNet + cross-view attention and write the result into
        `train_data[i]["image"]` as [H,W,3] fp32 on the CPU (gc_pipeline.py:159-237).
        With world_size > 1 the non-reference views are dealt round-robin to the ranks, every rank uploads only its own
        views + the references, the reference pass is sharded over its CFG rows (parallel.py), each reference view is
        decoded by one rank, and the decoded images are all-gathered so EVERY rank's train_data is complete (the
        fine-tune that follows samples any view on any rank)."""
        td = self.datamanager.train_data
        dev = self.device_
        plan = self.edit_plan(len(td))
        need, mine = plan["need"], plan["mine"]
        z = _pinned(torch.from_numpy(np.concatenate([td[i]["z_0_image"] for i in need], axis=0)))
        dep = _pinned(torch.from_numpy(np.concatenate([td[i]["depth_image"] for i in need], axis=0)))
        z_dev = z.to(dev, non_blocking=True).to(torch.float16)
        dep_dev = dep.to(dev, non_blocking=True)
        self.h2d_bytes = z.numel() * 4 + dep.numel() * 4
        masks = uned = None
        if all("mask_image" in td[i] for i in mine):
            masks = torch.from_numpy(np.stack([np.asarray(td[i]["mask_image"], dtype=np.float32) for i in mine])).to(dev)
            uned = torch.stack([td[i]["unedited_image"] for i in mine]).to(dev, torch.float16)
            self.h2d_bytes += masks.numel() * 4 + uned.numel() * 2
        imgs, ids = self.edit_on_device(z_dev, dep_dev, plan, masks, uned)
        host = torch.empty(imgs.shape, dtype=torch.float32, pin_memory=imgs.is_cuda)
        host.copy_(imgs, non_blocking=True)
        if imgs.is_cuda:
            torch.cuda.current_stream().synchronize()
        self.d2h_bytes = host.numel() * 4
        for j, i in enumerate(ids):
            # global_idx = image_idx (gc_pipeline.py:224,234); own storage per view: the datamanager deep-copies entries
            td[int(td[i].get("image_idx", i))]["image"] = host[j].clone()

    def edit_plan(self, V: int) -> dict:
        """Who edits what: `mine` = views this rank decodes, `view_ids` = its non-reference views (None on one GPU),
        `need` = views whose stage-A products it uploads (its own + the references), `pos_of` = view id -> row in `need`."""
        world, rank = self._dist()
        refs = list(self.ref_indices)
        from . import parallel as par
        if world > 1:
            view_ids = par.shard_views(V, world, rank, refs)
            mine = sorted(view_ids + [refs[i] for i in par.ref_decode_owner(self.num_ref_views, world, rank)])
            need = sorted(set(mine) | set(refs))
        else:
            view_ids, mine, need = None, list(range(V)), list(range(V))
        return {"V": V, "world": world, "rank": rank, "view_ids": view_ids, "mine": mine, "need": need,
                "pos_of": {v: j for j, v in enumerate(need)}}

    @torch.no_grad()
    def edit_on_device(self, z_dev: torch.Tensor, dep_dev: torch.Tensor, plan: dict, masks: Optional[torch.Tensor] = None,
                       uned: Optional[torch.Tensor] = None):
        """The device part of edit_images: z_dev [len(need),4,h,w] fp16 latents and dep_dev [len(need),H,W] fp32 depth of
        the views in `plan["need"]` -> (edited images [n,H,W,3] fp32 on the device, their view ids).  On several GPUs the
        images of all ranks are all-gathered, so n = V on every rank."""
        S, g = self.num_inference_steps, float(self.guidance_scale)
        if g <= 1.0:
            # diffusers does not double the batch for CFG then, and CrossViewAttnProcessor's `video_length = B // 2`
            # (utils.py:94) silently mixes views: every shipped script uses g in {3, 5, 7.5} (SURVEY §8a gotcha 2)
            raise ValueError(f"guidance_scale={g} <= 1: the cross-view attention layout needs classifier-free guidance")
        V, world, rank = plan["V"], plan["world"], plan["rank"]
        view_ids, mine, pos_of = plan["view_ids"], plan["mine"], plan["pos_of"]
        refs = list(self.ref_indices)
        R = self.num_ref_views
        from . import parallel as par
        # depth2disparity (numpy fp32, then .to(float16): gc_pipeline.py:182-186, 198-200), per view
        disparity = ops.nhwc_to_nchw(ops.depth_to_disparity(dep_dev.contiguous(), False))
        emb = self.prompt_encoder([self.negative_prompts, self.positive_prompt])
        neg, pos = emb[0:1], emb[1:2]
        if self.config.edit_schedule == "reference":
            if world > 1:
                raise ValueError("edit_schedule='reference' is the single-GPU literal schedule; multi-GPU runs use "
                                 "'refs_once'")
            outs = []
            for i in range(0, V, self.chunk_size):
                sel = refs + list(range(i, min(V, i + self.chunk_size)))
                outs.append(self.engine.edit_reference_schedule(z_dev[sel], disparity[sel], pos, neg, S, g, R,
                                                                ref_frames=crossview_ref_frames(R)))
            lat = torch.cat(outs)
        else:
            dist_ctx = None
            if world > 1:
                if getattr(self, "_kv_gather", None) is None:
                    self._kv_gather = par.make_kv_gather(self.device_)
                dist_ctx = {"world": world, "rank": rank, "gather": self._kv_gather}
            vb = max(1, getattr(self, "view_batch", getattr(self.config, "view_batch", self.chunk_size)))
            lat = self.engine.edit_refs_once(z_dev, disparity, [pos_of[r] for r in refs], pos, neg, S, g, view_batch=vb,
                                             view_ids=None if view_ids is None else [pos_of[v] for v in view_ids],
                                             dist_ctx=dist_ctx, ref_frames=crossview_ref_frames(R))
        imgs = self.vae.decode_latents(lat[[pos_of[i] for i in mine]], masks, uned)         # [n,H,W,3] fp32
        ids = mine
        if world > 1:
            imgs = par.gather_view_results(imgs, mine, V, world)    # every rank ends up with all V edited images
            ids = list(range(V))
        return imgs, ids

    # ------------------------------------------------------------------------------------------ helpers (same names)
    @torch.no_grad()
    def image2latent(self, image: torch.Tensor) -> torch.Tensor:
        """[H,W,3] in 0..1 -> [1,4,h,w] (gc_pipeline.py:239-246)."""
        return self.image2latent_batch(image[None])

    @torch.no_grad()
    def image2latent_batch(self, images: torch.Tensor) -> torch.Tensor:
        return self.vae.encode_mean(images.to(self.device_))

    def depth2disparity(self, depth):
        """numpy [1,H,W] -> [1,3,H,W] (gc_pipeline.py:248-256)."""
        d = torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32)).to(self.device_)
        out = ops.nhwc_to_nchw(ops.depth_to_disparity(d, False))
        return out.float().cpu().numpy()

    def depth2disparity_torch(self, depth: torch.Tensor) -> torch.Tensor:
        """torch [1,H,W] -> [1,3,H,W] (gc_pipeline.py:258-266), fp16 arithmetic when the input is fp16."""
        d = depth.to(self.device_)
        out = ops.nhwc_to_nchw(ops.depth_to_disparity(d.float().contiguous(), d.dtype == torch.float16))
        return out.to(depth.dtype)

    def update_datasets(self, cam_idx, unedited_image, depth, latent, mask):
        """gc_pipeline.py:268-274 (same signature: depth [H,W,1] tensor, latent [1,4,h,w] tensor)."""
        self._store_view(cam_idx, unedited_image, depth.permute(2, 0, 1).cpu().to(torch.float32).numpy(),
                         latent.cpu().to(torch.float32).numpy(), mask)

    # ---- B200 extension: stage-A products on disk in the reference's folder layout (store.py)
    def save_stage_a(self, root: str) -> None:
        from . import store
        store.save_train_data(root, self.datamanager.train_data)

    def load_stage_a(self, root: str) -> bool:
        """Fill train_data from `root` (depth_npy/, z_0/, mask_npy/, unedited/); True when edit_images can run without
        render_reverse."""
        from . import store
        store.load_train_data(root, len(self.datamanager.train_data), self.datamanager.train_data,
                              load_mask=self.config.langsam_obj != "")
        return store.has_stage_a(root)

    def get_train_loss_dict(self, step: int):
        ray_bundle, batch = self.datamanager.next_train(step)
        model_outputs = self._model(ray_bundle)  # gc_pipeline.py:276-287
        metrics_dict = self.model.get_metrics_dict(model_outputs, batch)
        loss_dict = self.model.get_loss_dict(model_outputs, batch, metrics_dict)
        return model_outputs, loss_dict, metrics_dict

    def forward(self):
        raise NotImplementedError


class SimpleDataManager:
    """Minimal datamanager carrying what the hot path touches: `cameras` and the `train_data` list of dicts
    (the full GaussCtrlDataManager - image loading, undistortion, view sub-sampling, gc_datamanager.py - is I/O
    outside the hot path)."""

    def __init__(self, cameras, train_data: Optional[List[Dict]] = None):
        self.cameras = cameras
        n = len(cameras)
        self.train_data = train_data if train_data is not None else [{"image_idx": i} for i in range(n)]
        self.train_unseen_cameras = list(range(n))
        self.device = "cuda"

    def next_train(self, step: int):
        """gc_datamanager.py:213-235: a random not-yet-seen view and a copy of its train_data entry."""
        from .finetune import next_train_view
        idx = next_train_view(self.train_unseen_cameras, len(self.train_data))
        data = dict(self.train_data[idx])
        data["image"] = torch.as_tensor(data["image"]).to(self.device)
        return self.cameras[idx:idx + 1].to(self.device), data
