"""Parameter tables of the SD1.x networks the hot path runs (UNet2DConditionModel, ControlNetModel(depth),
AutoencoderKL) with diffusers-0.26 state_dict key names, and the synthetic seeded initialisation used when no
checkpoint is available.  The reference builds these networks with
`StableDiffusionControlNetPipeline.from_pretrained(...)` / `ControlNetModel.from_pretrained(...)`
(gaussctrl/gc_pipeline.py:100-102); real safetensors weights load into the same tables unchanged."""
from __future__ import annotations

import math
from typing import Dict, Sequence, Tuple

import torch

BLOCK_OUT = (320, 640, 1280, 1280)
TEMB = 1280
CROSS_DIM = 768
COND_CH = (16, 32, 96, 256)
VAE_CH = (128, 256, 512, 512)

Shapes = Dict[str, Tuple[int, ...]]


def _conv(s: Shapes, name: str, cin: int, cout: int, k: int) -> None:
    s[name + ".weight"] = (cout, cin, k, k)
    s[name + ".bias"] = (cout,)


def _lin(s: Shapes, name: str, cin: int, cout: int, bias: bool = True) -> None:
    s[name + ".weight"] = (cout, cin)
    if bias:
        s[name + ".bias"] = (cout,)


def _norm(s: Shapes, name: str, c: int) -> None:
    s[name + ".weight"] = (c,)
    s[name + ".bias"] = (c,)


def _resnet(s: Shapes, p: str, cin: int, cout: int, temb: int | None) -> None:
    _norm(s, p + ".norm1", cin)
    _conv(s, p + ".conv1", cin, cout, 3)
    if temb is not None:
        _lin(s, p + ".time_emb_proj", temb, cout)
    _norm(s, p + ".norm2", cout)
    _conv(s, p + ".conv2", cout, cout, 3)
    if cin != cout:
        _conv(s, p + ".conv_shortcut", cin, cout, 1)


def _attn(s: Shapes, p: str, dim: int, kv_dim: int, bias: bool = False) -> None:
    _lin(s, p + ".to_q", dim, dim, bias)
    _lin(s, p + ".to_k", kv_dim, dim, bias)
    _lin(s, p + ".to_v", kv_dim, dim, bias)
    _lin(s, p + ".to_out.0", dim, dim, True)


def _transformer(s: Shapes, p: str, dim: int) -> None:
    _norm(s, p + ".norm", dim)
    _conv(s, p + ".proj_in", dim, dim, 1)
    b = p + ".transformer_blocks.0"
    _norm(s, b + ".norm1", dim)
    _attn(s, b + ".attn1", dim, dim)
    _norm(s, b + ".norm2", dim)
    _attn(s, b + ".attn2", dim, CROSS_DIM)
    _norm(s, b + ".norm3", dim)
    _lin(s, b + ".ff.net.0.proj", dim, dim * 8)
    _lin(s, b + ".ff.net.2", dim * 4, dim)
    _conv(s, p + ".proj_out", dim, dim, 1)


def _encoder(s: Shapes, ch: Sequence[int]) -> None:
    """conv_in, time embedding, 4 down blocks, mid block (shared by UNet and ControlNet)."""
    _conv(s, "conv_in", 4, ch[0], 3)
    _lin(s, "time_embedding.linear_1", ch[0], TEMB)
    _lin(s, "time_embedding.linear_2", TEMB, TEMB)
    cin = ch[0]
    for i, co in enumerate(ch):
        for j in range(2):
            _resnet(s, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else co, co, TEMB)
            if i < 3:
                _transformer(s, f"down_blocks.{i}.attentions.{j}", co)
        if i < 3:
            _conv(s, f"down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        cin = co
    _resnet(s, "mid_block.resnets.0", ch[3], ch[3], TEMB)
    _transformer(s, "mid_block.attentions.0", ch[3])
    _resnet(s, "mid_block.resnets.1", ch[3], ch[3], TEMB)


def skip_channels(ch: Sequence[int] = BLOCK_OUT):
    sk = [ch[0]]
    for i, co in enumerate(ch):
        sk += [co, co] + ([co] if i < 3 else [])
    return sk


def unet_shapes(ch: Sequence[int] = BLOCK_OUT) -> Shapes:
    s: Shapes = {}
    _encoder(s, ch)
    sk = skip_channels(ch)
    prev = ch[3]
    for i, co in enumerate(reversed(ch)):
        for j in range(3):
            cin = (prev if j == 0 else co) + sk.pop()
            _resnet(s, f"up_blocks.{i}.resnets.{j}", cin, co, TEMB)
            if i > 0:
                _transformer(s, f"up_blocks.{i}.attentions.{j}", co)
        if i < 3:
            _conv(s, f"up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        prev = co
    _norm(s, "conv_norm_out", ch[0])
    _conv(s, "conv_out", ch[0], 4, 3)
    return s


def controlnet_shapes(ch: Sequence[int] = BLOCK_OUT) -> Shapes:
    s: Shapes = {}
    _encoder(s, ch)
    e = "controlnet_cond_embedding"
    _conv(s, e + ".conv_in", 3, COND_CH[0], 3)
    for i in range(len(COND_CH) - 1):
        _conv(s, f"{e}.blocks.{2 * i}", COND_CH[i], COND_CH[i], 3)
        _conv(s, f"{e}.blocks.{2 * i + 1}", COND_CH[i], COND_CH[i + 1], 3)
    _conv(s, e + ".conv_out", COND_CH[-1], ch[0], 3)
    for i, c in enumerate(skip_channels(ch)):
        _conv(s, f"controlnet_down_blocks.{i}", c, c, 1)
    _conv(s, "controlnet_mid_block", ch[3], ch[3], 1)
    return s


def vae_shapes(ch: Sequence[int] = VAE_CH) -> Shapes:
    s: Shapes = {}

    def mid(p: str, c: int) -> None:
        _resnet(s, p + ".resnets.0", c, c, None)
        a = p + ".attentions.0"
        _norm(s, a + ".group_norm", c)
        _attn(s, a, c, c, bias=True)
        _resnet(s, p + ".resnets.1", c, c, None)

    _conv(s, "encoder.conv_in", 3, ch[0], 3)
    cin = ch[0]
    for i, co in enumerate(ch):
        for j in range(2):
            _resnet(s, f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else co, co, None)
        if i < len(ch) - 1:
            _conv(s, f"encoder.down_blocks.{i}.downsamplers.0.conv", co, co, 3)
        cin = co
    mid("encoder.mid_block", ch[-1])
    _norm(s, "encoder.conv_norm_out", ch[-1])
    _conv(s, "encoder.conv_out", ch[-1], 8, 3)
    rev = list(reversed(ch))
    _conv(s, "decoder.conv_in", 4, rev[0], 3)
    mid("decoder.mid_block", rev[0])
    cin = rev[0]
    for i, co in enumerate(rev):
        for j in range(3):
            _resnet(s, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else co, co, None)
        if i < len(rev) - 1:
            _conv(s, f"decoder.up_blocks.{i}.upsamplers.0.conv", co, co, 3)
        cin = co
    _norm(s, "decoder.conv_norm_out", rev[-1])
    _conv(s, "decoder.conv_out", rev[-1], 3, 3)
    _conv(s, "quant_conv", 8, 8, 1)
    _conv(s, "post_quant_conv", 4, 4, 1)
    return s


def random_state_dict(shapes: Shapes, seed: int, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """PyTorch-default-like init: U(+-1/sqrt(fan_in)) for conv/linear weights and biases, ones/zeros for norms."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shp in shapes.items():
        if name.endswith(".weight") and len(shp) == 1:
            out[name] = torch.ones(shp, dtype=dtype)
        elif name.endswith(".bias") and (name[:-5] + ".weight") in shapes and len(shapes[name[:-5] + ".weight"]) == 1:
            out[name] = torch.zeros(shp, dtype=dtype)
        else:
            wshape = shapes[name[:-5] + ".weight"] if name.endswith(".bias") else shp
            fan_in = math.prod(wshape[1:])
            bound = 1.0 / math.sqrt(fan_in)
            out[name] = ((torch.rand(shp, generator=g, dtype=torch.float32) * 2 - 1) * bound).to(dtype)
    return out


def synthetic_weights(seed: int = 0, with_vae: bool = True):
    """Synthetic-weight recipe of SURVEY §8d (ControlNet 'zero' convs N(0, 0.02^2); UNet conv_out x4)."""
    unet = random_state_dict(unet_shapes(), seed)
    cnet = random_state_dict(controlnet_shapes(), seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    for k in list(cnet):
        if (k.startswith("controlnet_down_blocks") or k.startswith("controlnet_mid_block")
                or k.startswith("controlnet_cond_embedding.conv_out")):
            if k.endswith(".weight"):
                cnet[k] = torch.randn(cnet[k].shape, generator=g) * 0.02
            else:
                cnet[k] = torch.zeros_like(cnet[k])
    unet["conv_out.weight"] = unet["conv_out.weight"] * 4.0
    vae = random_state_dict(vae_shapes(), seed + 3) if with_vae else None
    return unet, cnet, vae


class DDIMTables:
    """SD scheduler tables (DDIMScheduler / DDIMInverseScheduler from_pretrained(ckpt, subfolder="scheduler"),
    gc_pipeline.py:97-98): scaled-linear betas 0.00085..0.012 over 1000 steps, steps_offset 1, leading spacing,
    set_alpha_to_one False, epsilon prediction, eta 0."""

    def __init__(self, num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train = num_train
        self.steps_offset = steps_offset

    def timesteps(self, S: int):
        r = self.num_train // S
        return [int(i * r) + self.steps_offset for i in range(S)][::-1]

    def inverse_timesteps(self, S: int):
        r = self.num_train // S
        return [int(i * r) + self.steps_offset for i in range(S)]

    def step_coefs(self, t: int, S: int):
        """(sqrt a_t, sqrt(1-a_t), sqrt a_prev, sqrt(1-a_prev)) of DDIMScheduler.step."""
        prev = t - self.num_train // S
        a_t = float(self.alphas_cumprod[t])
        a_p = float(self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod)
        return (a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5)

    def inverse_step_coefs(self, t: int, S: int):
        """DDIMInverseScheduler.step: from alpha[t - ratio] (alpha[0] when negative) to alpha[t]."""
        cur = t - self.num_train // S
        a_t = float(self.alphas_cumprod[cur] if cur >= 0 else self.final_alpha_cumprod)
        a_p = float(self.alphas_cumprod[t])
        return (a_t ** 0.5, (1 - a_t) ** 0.5, a_p ** 0.5, (1 - a_p) ** 0.5)
