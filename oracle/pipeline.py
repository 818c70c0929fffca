"""Oracle: the two hot loops of GaussCtrlPipeline (TEST INFRASTRUCTURE – see oracle/__init__.py).

Restates, with the reference's batch layout and schedule:
    gc_pipeline.py:109-114  reference-view selection            -> `select_ref_indices`
    gc_pipeline.py:122-157  render_reverse (per-view inversion) -> `invert_view`
    gc_pipeline.py:159-237  edit_images  (per-chunk CFG sampling with refs recomputed in every chunk) -> `edit_chunk`
    gc_pipeline.py:239-266  image2latent / depth2disparity / depth2disparity_torch
and the body of diffusers' StableDiffusionControlNetPipeline.__call__ (0.26.0; PARITY UNPINNED) as used there."""
from __future__ import annotations

import random
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import sd15


def select_ref_indices(view_num: int, ref_view_num: int, clamp: bool = False) -> List[int]:
    """gc_pipeline.py:109-114 verbatim semantics (random.randint is inclusive: may return view_num)."""
    anchors = [(view_num * i) // ref_view_num for i in range(ref_view_num)] + [view_num]
    random.seed(13789)
    idx = [random.randint(a, anchors[i + 1]) for i, a in enumerate(anchors[:-1])]
    return [min(i, view_num - 1) for i in idx] if clamp else idx


def depth2disparity(depth: np.ndarray) -> np.ndarray:
    """gc_pipeline.py:248-256. depth [1,H,W] -> [1,3,H,W]."""
    disparity = 1 / (depth + 1e-5)
    disparity_map = disparity / np.max(disparity)
    return np.concatenate([disparity_map] * 3, axis=0)[None]


def depth2disparity_torch(depth: torch.Tensor) -> torch.Tensor:
    """gc_pipeline.py:258-266."""
    disparity = 1 / (depth + 1e-5)
    disparity_map = disparity / torch.max(disparity)
    return torch.concatenate([disparity_map] * 3, dim=0)[None]


def image2latent(vae: sd15.AutoencoderKL, image: torch.Tensor) -> torch.Tensor:
    """gc_pipeline.py:239-246. image [H,W,3] in 0..1 -> [1,4,H/8,W/8]."""
    x = (image * 2 - 1).permute(2, 0, 1).unsqueeze(0)
    return vae.encode_mean(x) * 0.18215


@torch.no_grad()
def invert_view(unet, cnet, tables: sd15.DDIMTables, z0, disparity, prompt_embed, S: int):
    """The `pipe(...)` call of render_reverse (gc_pipeline.py:142-145): guidance_scale=0 => no CFG, batch 1,
    vanilla attention, DDIMInverseScheduler, ascending timesteps."""
    unet.set_attn_processor(sd15.vanilla_processor)
    cnet.set_attn_processor(sd15.vanilla_processor)
    x = z0
    for t in tables.inverse_timesteps(S):
        down, mid = cnet(x, int(t), prompt_embed, disparity, 1.0)
        eps = unet(x, int(t), prompt_embed, down, mid)
        x = tables.inverse_step(eps, int(t), x, S)
    return x


@torch.no_grad()
def edit_chunk(unet, cnet, vae, tables: sd15.DDIMTables, latents, disparity, pos_embed, neg_embed, S: int,
               guidance_scale: float, num_ref: int, ref_frames: Sequence[int] = (0, 1, 2, 3),
               decode: bool = True, return_latents: bool = False, stop_after: Optional[int] = None):
    """One `pipe(...)` call of edit_images (gc_pipeline.py:209-219) on F = R + c frames.

    latents [F,4,h,w], disparity [F,3,H,W]; pos/neg_embed [1,77,D].  CFG batch = cat([uncond, cond]) (2F rows),
    CrossViewAttnProcessor(0.6) in the UNet and (0.0) in the ControlNet, DDIM eta=0, descending timesteps.
    Returns decoded images [c,3,H,W] (refs dropped, `.images[num_ref:]`)."""
    assert guidance_scale > 1.0, "reference assumes CFG doubling (utils.py:94); SURVEY §8a gotcha 2"
    F_ = latents.shape[0]
    unet.set_attn_processor(sd15.CrossViewProcessor(0.6, 2, ref_frames))
    cnet.set_attn_processor(sd15.CrossViewProcessor(0.0, 2, ref_frames))
    ehs = torch.cat([neg_embed.expand(F_, -1, -1), pos_embed.expand(F_, -1, -1)], dim=0)
    cond = torch.cat([disparity] * 2, dim=0)
    x = latents
    for t in list(tables.timesteps(S))[:stop_after]:  # stop_after: test-only truncation of a long schedule (cfg5: S=50)
        xin = torch.cat([x] * 2, dim=0)
        down, mid = cnet(xin, int(t), ehs, cond, 1.0)
        eps = unet(xin, int(t), ehs, down, mid)
        eps_u, eps_c = eps.chunk(2)
        eps = eps_u + guidance_scale * (eps_c - eps_u)
        x = tables.step(eps, int(t), x, S)
    if return_latents or not decode or vae is None:
        return x[num_ref:]
    img = vae.decode(x / vae.scaling_factor)
    img = (img / 2 + 0.5).clamp(0, 1)
    return img[num_ref:]


def composite_mask(edited: torch.Tensor, unedited_hw3: torch.Tensor, mask: Optional[np.ndarray]) -> torch.Tensor:
    """gc_pipeline.py:223-234: edited [3,H,W]; returns [H,W,3] float32."""
    out = edited
    if mask is not None:
        m = torch.from_numpy(np.asarray(mask))
        out = edited * m[None] + unedited_hw3.permute(2, 0, 1) * (1 - m)[None]
    return out.permute(1, 2, 0).to(torch.float32)
