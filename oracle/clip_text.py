"""ORACLE (test infrastructure only): CPU restatement of the CLIP text encoder behind the reference's prompts
(SURVEY §8f row 4).

Reference call sites: `self.pipe(prompt=[positive]*F, negative_prompt=[negative]*F, ...)` (gaussctrl/gc_pipeline.py:
142-145 and :209-219) -> diffusers 0.26.0 `encode_prompt` -> `text_encoder(input_ids, attention_mask=None)[0]`, the
`CLIPTextModel` of the SD1.x checkpoint (`transformers>=4.38.0`, requirements.txt:1): 12 pre-LN transformer layers,
width 768, 12 heads, MLP 3072 with quick-GELU, learned positions (77), CAUSAL self-attention, final LayerNorm;
the pipeline conditions on `last_hidden_state` [B,77,768].

PINNED: transformers IS installed in the build container, so this restatement is checked against the real
`transformers.CLIPTextModel` (seeded random init, SD1.x config) in tests/test_clip_cpu.py to fp32 round-off, with the
state_dict key names of that module."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

HIDDEN, HEADS, LAYERS, MLP, VOCAB, MAX_POS, LN_EPS = 768, 12, 12, 3072, 49408, 77, 1e-5


def quick_gelu(x: torch.Tensor) -> torch.Tensor:
    return x * torch.sigmoid(1.702 * x)


def clip_text_forward(sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, layers: int = LAYERS, heads: int = HEADS):
    """sd: CLIPTextModel state_dict (keys `text_model.*`); input_ids [B,T] int64 -> last_hidden_state [B,T,C]."""
    p = "text_model."
    B, T = input_ids.shape
    x = sd[p + "embeddings.token_embedding.weight"][input_ids] + sd[p + "embeddings.position_embedding.weight"][:T]
    C = x.shape[-1]
    d = C // heads
    causal = torch.full((T, T), float("-inf"), dtype=x.dtype).triu(1)
    for i in range(layers):
        q = f"{p}encoder.layers.{i}."
        h = F.layer_norm(x, (C,), sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"], LN_EPS)
        qq = F.linear(h, sd[q + "self_attn.q_proj.weight"], sd[q + "self_attn.q_proj.bias"]) * d ** -0.5
        kk = F.linear(h, sd[q + "self_attn.k_proj.weight"], sd[q + "self_attn.k_proj.bias"])
        vv = F.linear(h, sd[q + "self_attn.v_proj.weight"], sd[q + "self_attn.v_proj.bias"])
        qq, kk, vv = (t.reshape(B, T, heads, d).transpose(1, 2) for t in (qq, kk, vv))
        att = torch.softmax(qq @ kk.transpose(-1, -2) + causal, dim=-1) @ vv
        att = att.transpose(1, 2).reshape(B, T, C)
        x = x + F.linear(att, sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (C,), sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"], LN_EPS)
        h = quick_gelu(F.linear(h, sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"]))
        x = x + F.linear(h, sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"])
    return F.layer_norm(x, (C,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], LN_EPS)


def seeded_state_dict(seed: int = 0, layers: int = LAYERS, scale: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random CLIPTextModel-shaped weights (no checkpoint on disk): N(0, 0.02^2) matrices, LayerNorm ~ (1, 0) + noise."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    p = "text_model."

    def w(*shape, std=0.02):
        return torch.randn(shape, generator=g) * std * scale

    sd[p + "embeddings.token_embedding.weight"] = w(VOCAB, HIDDEN)
    sd[p + "embeddings.position_embedding.weight"] = w(MAX_POS, HIDDEN, std=0.01)
    for i in range(layers):
        q = f"{p}encoder.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[q + f"self_attn.{n}.weight"] = w(HIDDEN, HIDDEN, std=0.03)
            sd[q + f"self_attn.{n}.bias"] = w(HIDDEN)
        for n in ("layer_norm1", "layer_norm2"):
            sd[q + n + ".weight"] = 1.0 + w(HIDDEN)
            sd[q + n + ".bias"] = w(HIDDEN)
        sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = w(MLP, HIDDEN, std=0.03), w(MLP)
        sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = w(HIDDEN, MLP, std=0.02), w(HIDDEN)
    sd[p + "final_layer_norm.weight"] = 1.0 + w(HIDDEN)
    sd[p + "final_layer_norm.bias"] = w(HIDDEN)
    return sd
