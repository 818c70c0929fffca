"""Oracle: cross-view attention (TEST INFRASTRUCTURE – see oracle/__init__.py).

Restates `gaussctrl/utils.py:25-133` (`compute_attn`, `CrossViewAttnProcessor.__call__`) and the helper
methods of `diffusers.models.attention_processor.Attention` (diffusers==0.26.0, requirements.txt:2) that the
processor calls: `head_to_batch_dim`, `batch_to_head_dim`, `get_attention_scores` (= baddbmm(alpha=scale) ->
softmax(dim=-1) -> cast), `prepare_attention_mask` (None passthrough).

Two formulations are provided:
  * `crossview_attention_literal`  – the reference's 5-pass structure, pass by pass (utils.py:88-117);
  * `multi_source_attention`       – the fused formulation the CUDA kernel implements:
        out = sum_s w_s * softmax(q K_s^T * scale) V_s          (independent softmax per source)
    which equals the literal one for sources = [self, ref0..ref3], w = [c, (1-c)/4 x4].
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch


class AttentionStub(torch.nn.Module):
    """Minimal stand-in for diffusers' `Attention` module (0.26.0) with exactly the attributes
    `CrossViewAttnProcessor.__call__` touches (utils.py:56-131).  SD1.x transformer blocks use:
    no spatial_norm, no group_norm, no norm_cross, residual_connection=False, rescale_output_factor=1,
    to_q/k/v without bias, to_out = [Linear(with bias), Dropout(0)], upcast_softmax=False, scale=d^-0.5."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int] = None,
                 bias: bool = False, upcast_softmax: bool = False):
        super().__init__()
        inner = heads * dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.scale = dim_head ** -0.5
        self.upcast_softmax = upcast_softmax
        self.upcast_attention = False
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.to_q = torch.nn.Linear(query_dim, inner, bias=bias)
        self.to_k = torch.nn.Linear(kv_dim, inner, bias=bias)
        self.to_v = torch.nn.Linear(kv_dim, inner, bias=bias)
        self.to_out = torch.nn.ModuleList([torch.nn.Linear(inner, query_dim), torch.nn.Dropout(0.0)])

    # diffusers Attention.prepare_attention_mask: returns None when the mask is None
    def prepare_attention_mask(self, attention_mask, target_length, batch_size, out_dim=3):
        if attention_mask is None:
            return None
        raise NotImplementedError("oracle only models the mask=None path used by the reference")

    # [B, N, h*d] -> [B*h, N, d]
    def head_to_batch_dim(self, tensor, out_dim=3):
        b, n, c = tensor.shape
        h = self.heads
        tensor = tensor.reshape(b, n, h, c // h).permute(0, 2, 1, 3)
        return tensor.reshape(b * h, n, c // h)

    # [B*h, N, d] -> [B, N, h*d]
    def batch_to_head_dim(self, tensor):
        bh, n, d = tensor.shape
        h = self.heads
        tensor = tensor.reshape(bh // h, h, n, d).permute(0, 2, 1, 3)
        return tensor.reshape(bh // h, n, h * d)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        if attention_mask is None:
            baddbmm_input = torch.empty(query.shape[0], query.shape[1], key.shape[1], dtype=query.dtype,
                                        device=query.device)
            beta = 0
        else:
            baddbmm_input, beta = attention_mask, 1
        scores = torch.baddbmm(baddbmm_input, query, key.transpose(-1, -2), beta=beta, alpha=self.scale)
        if self.upcast_softmax:
            scores = scores.float()
        probs = scores.softmax(dim=-1)
        return probs.to(dtype)


def _gather_frame(x: torch.Tensor, video_length: int, frame: int) -> torch.Tensor:
    """utils.py:26-31: rearrange '(b f) n c -> b f n c', index [:, [frame]*f], flatten back."""
    bf, n, c = x.shape
    b = bf // video_length
    x4 = x.reshape(b, video_length, n, c)
    x4 = x4[:, [frame] * video_length]
    return x4.reshape(bf, n, c)


def crossview_attention_literal(attn: AttentionStub, hidden_states: torch.Tensor,
                                encoder_hidden_states: Optional[torch.Tensor], self_attn_coeff: float,
                                unet_chunk_size: int = 2, ref_frames: Sequence[int] = (0, 1, 2, 3)) -> torch.Tensor:
    """Pass-by-pass restatement of CrossViewAttnProcessor.__call__ (utils.py:44-133) for 3-D inputs.

    `ref_frames=(0,1,2,3)` is the reference's hard-coded behaviour (utils.py:95-98).  Other tuples give the
    generalised-R semantics (SURVEY §8a gotcha 1): mean over the listed frames."""
    query = attn.to_q(hidden_states)
    is_cross = encoder_hidden_states is not None
    ehs = hidden_states if not is_cross else encoder_hidden_states
    key = attn.to_k(ehs)
    value = attn.to_v(ehs)
    query = attn.head_to_batch_dim(query)
    if is_cross:
        probs = attn.get_attention_scores(query, attn.head_to_batch_dim(key), None)
        out = torch.bmm(probs, attn.head_to_batch_dim(value))
    else:
        probs = attn.get_attention_scores(query, attn.head_to_batch_dim(key), None)
        h_self = torch.bmm(probs, attn.head_to_batch_dim(value))
        video_length = key.shape[0] // unet_chunk_size
        h_refs = []
        for r in ref_frames:
            k_r = attn.head_to_batch_dim(_gather_frame(key, video_length, r))
            v_r = attn.head_to_batch_dim(_gather_frame(value, video_length, r))
            p_r = attn.get_attention_scores(query, k_r, None)
            h_refs.append(torch.bmm(p_r, v_r))
        out = self_attn_coeff * h_self + (1 - self_attn_coeff) * torch.mean(torch.stack(h_refs), dim=0)
    out = attn.batch_to_head_dim(out)
    out = attn.to_out[0](out)
    out = attn.to_out[1](out)
    if attn.residual_connection:
        out = out + hidden_states
    return out / attn.rescale_output_factor


def multi_source_attention(q: torch.Tensor, k_sources: List[torch.Tensor], v_sources: List[torch.Tensor],
                           weights: Sequence[float], heads: int, scale: Optional[float] = None) -> torch.Tensor:
    """Fused formulation (what `gcb_attn_multi_fwd` computes), fp32 math.

    q: [B, N, h*d]; k_sources[s], v_sources[s]: [B, M_s, h*d]; returns [B, N, h*d] =
    sum_s weights[s] * softmax(q k_s^T * scale) v_s, softmax independent per source and per head."""
    b, n, c = q.shape
    d = c // heads
    scale = d ** -0.5 if scale is None else scale
    qh = q.float().reshape(b, n, heads, d).permute(0, 2, 1, 3)
    out = torch.zeros_like(qh)
    for ks, vs, w in zip(k_sources, v_sources, weights):
        if w == 0.0:
            continue
        kh = ks.float().reshape(b, -1, heads, d).permute(0, 2, 1, 3)
        vh = vs.float().reshape(b, -1, heads, d).permute(0, 2, 1, 3)
        p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
        out = out + w * (p @ vh)
    return out.permute(0, 2, 1, 3).reshape(b, n, c)


def crossview_sources(k: torch.Tensor, v: torch.Tensor, video_length: int, ref_frames: Sequence[int],
                      self_attn_coeff: float):
    """Build the (k_sources, v_sources, weights) triple equivalent to the reference's 5 passes."""
    ks, vs, ws = [k], [v], [float(self_attn_coeff)]
    for r in ref_frames:
        ks.append(_gather_frame(k, video_length, r))
        vs.append(_gather_frame(v, video_length, r))
        ws.append((1.0 - float(self_attn_coeff)) / len(ref_frames))
    return ks, vs, ws
