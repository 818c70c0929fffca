"""Loading real diffusers checkpoints into the parameter layout of sd15_spec (the key names ARE diffusers': SURVEY §8c).

Replaces `StableDiffusionControlNetPipeline.from_pretrained(diffusion_ckpt, controlnet=ControlNetModel.from_pretrained(
"lllyasviel/sd-controlnet-depth"))` (gaussctrl/gc_pipeline.py:97-102) for checkpoints that are already on disk in the
diffusers folder layout (there is no network here, hub names cannot be resolved):

    <ckpt>/unet/diffusion_pytorch_model[.fp16].safetensors
    <ckpt>/vae/diffusion_pytorch_model[.fp16].safetensors
    <controlnet>/diffusion_pytorch_model[.fp16].safetensors        (default <ckpt>/controlnet, or $GCB_CONTROLNET_CKPT)

Every tensor is checked against sd15_spec's shape table; a missing, extra or mis-shaped tensor raises.  The VAE
mid-block attention of SD1.x checkpoints uses the pre-0.18 names (query/key/value/proj_attn, sometimes stored as
1x1-conv-shaped linears): they are renamed/reshaped exactly as diffusers' loader does."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch

from . import sd15_spec as sp

_VAE_ATTN_RENAME = {"query": "to_q", "key": "to_k", "value": "to_v", "proj_attn": "to_out.0"}
_FILE_CANDIDATES = ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors")


def _find(folder: str) -> str:
    for name in _FILE_CANDIDATES:
        path = os.path.join(folder, name)
        if os.path.isfile(path):
            return path
    raise FileNotFoundError(f"no {' / '.join(_FILE_CANDIDATES)} under {folder}")


def _convert_vae_keys(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        parts = k.split(".")
        if "attentions" in parts and len(parts) >= 2 and parts[-2] in _VAE_ATTN_RENAME:
            k = ".".join(parts[:-2] + [_VAE_ATTN_RENAME[parts[-2]], parts[-1]])
        if "attentions" in parts and k.endswith(".weight") and v.dim() == 4:   # conv-shaped linear [C,C,1,1]
            v = v[:, :, 0, 0]
        out[k] = v
    return out


def check_state_dict(sd: Dict[str, torch.Tensor], shapes: Dict[str, Tuple[int, ...]], what: str) -> Dict[str, torch.Tensor]:
    missing = sorted(set(shapes) - set(sd))
    extra = sorted(set(sd) - set(shapes))
    if missing or extra:
        raise ValueError(f"{what}: {len(missing)} missing / {len(extra)} unexpected tensors "
                         f"(first missing {missing[:3]}, first unexpected {extra[:3]})")
    for k, shp in shapes.items():
        if tuple(sd[k].shape) != tuple(shp):
            raise ValueError(f"{what}: {k} has shape {tuple(sd[k].shape)}, expected {tuple(shp)}")
    return {k: sd[k].to(torch.float32) for k in shapes}


def load_component(folder: str, shapes: Dict[str, Tuple[int, ...]], what: str, vae: bool = False) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file
    sd = load_file(_find(folder))
    if vae:
        sd = _convert_vae_keys(sd)
    return check_state_dict(sd, shapes, what)


def load_diffusers_checkpoint(ckpt: str, controlnet: Optional[str] = None, with_vae: bool = True):
    """-> (unet_sd, controlnet_sd, vae_sd) fp32 state dicts under diffusers key names."""
    controlnet = controlnet or os.environ.get("GCB_CONTROLNET_CKPT") or os.path.join(ckpt, "controlnet")
    unet = load_component(os.path.join(ckpt, "unet"), sp.unet_shapes(), "unet")
    cnet = load_component(controlnet, sp.controlnet_shapes(), "controlnet")
    vae = load_component(os.path.join(ckpt, "vae"), sp.vae_shapes(), "vae", vae=True) if with_vae else None
    return unet, cnet, vae


def save_diffusers_checkpoint(ckpt: str, unet, cnet, vae=None, dtype=torch.float16) -> None:
    """Write state dicts in the same folder layout (used to snapshot the synthetic weights of a benchmark run)."""
    from safetensors.torch import save_file
    for sub, sd in (("unet", unet), ("controlnet", cnet), ("vae", vae)):
        if sd is None:
            continue
        os.makedirs(os.path.join(ckpt, sub), exist_ok=True)
        save_file({k: v.to(dtype).contiguous() for k, v in sd.items()}, os.path.join(ckpt, sub, _FILE_CANDIDATES[0]))
