"""Kernel-native cross-view attention cases shared by the golden generator (make_golden.py, which runs the REFERENCE's
utils.py on them) and the GPU parity test (test_processor_gpu.py, which runs gaussctrl_b200.utils on them).

Inputs are regenerated from the seed (the tensors are too large to commit: up to 26 MB per case) and rounded to fp16
so the reference's fp32 run and the fp16 CUDA run see identical numbers; a fingerprint of the inputs is stored with
the golden output so RNG drift between torch builds fails loudly instead of comparing different problems."""
from __future__ import annotations

import hashlib

import numpy as np
import torch

# name: (heads, dim_head, N tokens, frames per CFG half F, self_attn_coeff, cross_attention_dim or None, text length)
NATIVE_CASES = {
    "d40_F7_unet": (8, 40, 256, 7, 0.6, None, 0),      # R=4 + c=3 (BASELINE cfg2 chunk), UNet coefficient
    "d40_F7_cnet": (8, 40, 256, 7, 0.0, None, 0),      # ControlNet coefficient: self pass weighted 0
    "d80_F7_unet": (8, 80, 256, 7, 0.6, None, 0),
    "d160_F5_unet": (8, 160, 256, 5, 0.6, None, 0),    # mma.sync path (16x16 level head dim)
    "d40_F12_unet": (8, 40, 256, 12, 0.6, None, 0),    # R=8 + c=4: only frames 0..3 of the 8 references are sources
    "d160_N64_F7": (8, 160, 64, 7, 0.6, None, 0),      # 8x8 level
    "text_d40": (8, 40, 256, 7, 0.6, 768, 77),         # attn2: 77 CLIP tokens
    "text_d160": (8, 160, 64, 7, 0.6, 768, 77),
}
TOKEN_STRIDE = 16  # the committed golden keeps every 16th token row (+ the mean over all tokens)


def native_case_inputs(name: str):
    """-> (attn state_dict (fp16-representable fp32), hidden [2F,N,C], encoder_hidden_states or None, meta tuple)"""
    from oracle.crossview_attn import AttentionStub
    heads, dh, n, f, coeff, cross, ntext = NATIVE_CASES[name]
    seed = 20000 + sorted(NATIVE_CASES).index(name)
    torch.manual_seed(seed)
    c = heads * dh
    attn = AttentionStub(c, heads, dh, cross_attention_dim=cross)
    sd = {k: v.detach().half().float() for k, v in attn.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    hs = torch.randn((2 * f, n, c), generator=g).half().float()
    ehs = torch.randn((2 * f, ntext, cross), generator=g).half().float() if cross is not None else None
    return sd, hs, ehs, NATIVE_CASES[name]


def fingerprint(sd, hs, ehs) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(sd[k].numpy().tobytes())
    h.update(hs.numpy().tobytes())
    if ehs is not None:
        h.update(ehs.numpy().tobytes())
    return h.hexdigest()


def subsample(out: torch.Tensor):
    """-> (every TOKEN_STRIDE-th token row fp32, mean over all tokens fp64)"""
    o = out.detach()
    return o[:, ::TOKEN_STRIDE].contiguous().numpy().astype(np.float32), o.double().mean(dim=1).numpy()
