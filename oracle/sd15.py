"""Oracle: SD1.x UNet2DConditionModel / ControlNetModel(depth) / AutoencoderKL / DDIM schedulers
(TEST INFRASTRUCTURE – see oracle/__init__.py).

PARITY UNPINNED.  The arithmetic lives in diffusers==0.26.0 (requirements.txt:2), un-vendored and not
installed here, and no checkpoints exist on disk.  This file restates the published SD1.x architecture
 as plain torch modules whose `state_dict()` keys are the diffusers keys, so real
safetensors weights would load unchanged.  Reference call sites it is anchored on:
    gc_pipeline.py:97-102   schedulers + ControlNetModel + StableDiffusionControlNetPipeline (fp16)
    gc_pipeline.py:142-145  inversion call  (guidance_scale=0, output_type='latent')
    gc_pipeline.py:209-219  edit call       (CFG, controlnet_conditioning_scale=1.0, eta=0, output_type='pt')
    gc_pipeline.py:239-246  image2latent    (vae.encode(...).latent_dist.mean * 0.18215)
Attention modules call a pluggable processor with the diffusers signature `proc(attn, hidden, encoder_hidden_states)`
so `oracle.crossview_attn.crossview_attention_literal` (= utils.py:44-133) can be plugged in exactly like
`set_attn_processor` does (gc_pipeline.py:136-137,163-168)."""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .crossview_attn import AttentionStub, crossview_attention_literal


# ----------------------------------------------------------------------------- attention processors
def vanilla_processor(attn: AttentionStub, hidden, encoder_hidden_states=None):
    """diffusers AttnProcessor (set at gc_pipeline.py:136-137): plain softmax(QK^T)V."""
    ehs = hidden if encoder_hidden_states is None else encoder_hidden_states
    q = attn.head_to_batch_dim(attn.to_q(hidden))
    k = attn.head_to_batch_dim(attn.to_k(ehs))
    v = attn.head_to_batch_dim(attn.to_v(ehs))
    p = attn.get_attention_scores(q, k, None)
    out = attn.batch_to_head_dim(torch.bmm(p, v))
    return attn.to_out[1](attn.to_out[0](out))


class CrossViewProcessor:
    """Same ctor as the reference's CrossViewAttnProcessor (utils.py:40-42)."""

    def __init__(self, self_attn_coeff, unet_chunk_size=2, ref_frames: Sequence[int] = (0, 1, 2, 3)):
        self.self_attn_coeff = self_attn_coeff
        self.unet_chunk_size = unet_chunk_size
        self.ref_frames = tuple(ref_frames)

    def __call__(self, attn, hidden, encoder_hidden_states=None):
        return crossview_attention_literal(attn, hidden, encoder_hidden_states, self.self_attn_coeff,
                                           self.unet_chunk_size, self.ref_frames)


# ----------------------------------------------------------------------------- building blocks
def timestep_embedding(t: torch.Tensor, dim: int = 320) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch: Optional[int] = 1280, eps=1e-5, groups=32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout) if temb_ch is not None else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class GEGLU(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.proj = nn.Linear(cin, cout * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, cross_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = AttentionStub(dim, heads, dim // heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = AttentionStub(dim, heads, dim // heads, cross_attention_dim=cross_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)
        self.processor: Callable = vanilla_processor

    def forward(self, x, ehs):
        x = self.processor(self.attn1, self.norm1(x), None) + x
        x = self.processor(self.attn2, self.norm2(x), ehs) + x
        return self.ff(self.norm3(x)) + x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads=8, cross_dim=768):
        super().__init__()
        self.norm = nn.GroupNorm(32, dim, eps=1e-6)
        self.proj_in = nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, cross_dim)])
        self.proj_out = nn.Conv2d(dim, dim, 1)

    def forward(self, x, ehs):
        b, c, h, w = x.shape
        res = x
        y = self.proj_in(self.norm(x))
        y = y.permute(0, 2, 3, 1).reshape(b, h * w, c)
        for blk in self.transformer_blocks:
            y = blk(y, ehs)
        y = y.reshape(b, h, w, c).permute(0, 3, 1, 2)
        return self.proj_out(y) + res


class Downsample2D(nn.Module):
    def __init__(self, ch, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cin, cout, has_attn, add_down, heads=8, cross_dim=768):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin, cout), ResnetBlock2D(cout, cout)])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in range(2)])
        self.has_attn = has_attn
        if add_down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.add_down = add_down

    def forward(self, x, temb, ehs):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ehs)
            outs.append(x)
        if self.add_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, heads=8, cross_dim=768):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch), ResnetBlock2D(ch, ch)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, cross_dim)])

    def forward(self, x, temb, ehs):
        x = self.resnets[0](x, temb)
        x = self.attentions[0](x, ehs)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, in_chs: Sequence[int], cout, has_attn, add_up, heads=8, cross_dim=768):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, cout) for c in in_chs])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, cross_dim) for _ in in_chs])
        self.has_attn = has_attn
        if add_up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.add_up = add_up

    def forward(self, x, skips: List[torch.Tensor], temb, ehs):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, temb)
            if self.has_attn:
                x = self.attentions[i](x, ehs)
        if self.add_up:
            x = self.upsamplers[0](x)
        return x


BLOCK_OUT = (320, 640, 1280, 1280)
UP_IN = ((2560, 2560, 2560), (2560, 2560, 1920), (1920, 1280, 960), (960, 640, 640))
UP_OUT = (1280, 1280, 640, 320)


def _scaled(chs, width_div):
    return tuple(max(32, c // width_div) for c in chs)


class UNet2DConditionModel(nn.Module):
    """SD1.x UNet.  `width_div` > 1 shrinks channel widths for fast CPU tests (architecture unchanged)."""

    def __init__(self, width_div: int = 1, cross_dim: int = 768, heads: int = 8):
        super().__init__()
        ch = _scaled(BLOCK_OUT, width_div)
        self.ch = ch
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(4, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.down_blocks = nn.ModuleList()
        cin = ch[0]
        for i, co in enumerate(ch):
            blk = DownBlock(cin, co, has_attn=(i < 3), add_down=(i < 3), heads=heads, cross_dim=cross_dim)
            self.down_blocks.append(blk)
            cin = co
        self.mid_block = MidBlock(ch[3], heads, cross_dim)
        # skip-channel bookkeeping identical to diffusers (pop from the end of the residual list)
        skip_ch = [ch[0]]
        for i, co in enumerate(ch):
            skip_ch += [co, co] + ([co] if i < 3 else [])
        self.up_blocks = nn.ModuleList()
        prev = ch[3]
        rev = list(reversed(ch))
        for i, co in enumerate(rev):
            ins = []
            for j in range(3):
                ins.append((prev if j == 0 else co) + skip_ch.pop())
            self.up_blocks.append(UpBlock(ins, co, has_attn=(i > 0), add_up=(i < 3), heads=heads, cross_dim=cross_dim))
            prev = co
        self.conv_norm_out = nn.GroupNorm(32, ch[0], eps=1e-5)
        self.conv_out = nn.Conv2d(ch[0], 4, 3, padding=1)
        for m in self.modules():
            if isinstance(m, ResnetBlock2D) and m.time_emb_proj is not None and m.time_emb_proj.in_features != temb:
                m.time_emb_proj = nn.Linear(temb, m.time_emb_proj.out_features)

    def set_attn_processor(self, proc):
        for m in self.modules():
            if isinstance(m, BasicTransformerBlock):
                m.processor = proc

    def forward(self, sample, t, ehs, down_res: Optional[Sequence[torch.Tensor]] = None, mid_res=None):
        tt = torch.as_tensor(t).reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(tt.cpu(), self.ch[0]).to(device=sample.device, dtype=sample.dtype))
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, ehs)
            skips += outs
        if down_res is not None:
            skips = [s + r for s, r in zip(skips, down_res)]
        x = self.mid_block(x, temb, ehs)
        if mid_res is not None:
            x = x + mid_res
        for blk in self.up_blocks:
            x = blk(x, skips, temb, ehs)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class ControlNetConditioningEmbedding(nn.Module):
    def __init__(self, cout, chs=(16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(3, chs[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        for i in range(len(chs) - 1):
            self.blocks.append(nn.Conv2d(chs[i], chs[i], 3, padding=1))
            self.blocks.append(nn.Conv2d(chs[i], chs[i + 1], 3, padding=1, stride=2))
        self.conv_out = nn.Conv2d(chs[-1], cout, 3, padding=1)

    def forward(self, c):
        x = F.silu(self.conv_in(c))
        for b in self.blocks:
            x = F.silu(b(x))
        return self.conv_out(x)


class ControlNetModel(nn.Module):
    def __init__(self, width_div: int = 1, cross_dim: int = 768, heads: int = 8):
        super().__init__()
        ch = _scaled(BLOCK_OUT, width_div)
        self.ch = ch
        temb = ch[0] * 4
        self.conv_in = nn.Conv2d(4, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb)
        self.controlnet_cond_embedding = ControlNetConditioningEmbedding(ch[0])
        self.down_blocks = nn.ModuleList()
        self.controlnet_down_blocks = nn.ModuleList([nn.Conv2d(ch[0], ch[0], 1)])
        cin = ch[0]
        for i, co in enumerate(ch):
            self.down_blocks.append(DownBlock(cin, co, has_attn=(i < 3), add_down=(i < 3), heads=heads,
                                              cross_dim=cross_dim))
            for _ in range(2 + (1 if i < 3 else 0)):
                self.controlnet_down_blocks.append(nn.Conv2d(co, co, 1))
            cin = co
        self.mid_block = MidBlock(ch[3], heads, cross_dim)
        self.controlnet_mid_block = nn.Conv2d(ch[3], ch[3], 1)
        for m in self.modules():
            if isinstance(m, ResnetBlock2D) and m.time_emb_proj is not None and m.time_emb_proj.in_features != temb:
                m.time_emb_proj = nn.Linear(temb, m.time_emb_proj.out_features)

    def set_attn_processor(self, proc):
        for m in self.modules():
            if isinstance(m, BasicTransformerBlock):
                m.processor = proc

    def forward(self, sample, t, ehs, cond, conditioning_scale: float = 1.0):
        tt = torch.as_tensor(t).reshape(-1).expand(sample.shape[0])
        temb = self.time_embedding(timestep_embedding(tt.cpu(), self.ch[0]).to(device=sample.device, dtype=sample.dtype))
        x = self.conv_in(sample) + self.controlnet_cond_embedding(cond)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, temb, ehs)
            skips += outs
        x = self.mid_block(x, temb, ehs)
        down = [conv(s) * conditioning_scale for conv, s in zip(self.controlnet_down_blocks, skips)]
        mid = self.controlnet_mid_block(x) * conditioning_scale
        return down, mid


# ----------------------------------------------------------------------------- VAE
class VaeAttention(nn.Module):
    """diffusers Attention as used in the VAE mid block: 1 head, GroupNorm(32, eps 1e-6), biased q/k/v,
    residual connection, 4-D input."""

    def __init__(self, ch):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x).reshape(b, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)
        p = torch.softmax(q @ k.transpose(1, 2) * (c ** -0.5), dim=-1)
        o = self.to_out[0](p @ v)
        return o.transpose(1, 2).reshape(b, c, h, w) + x


class VaeMid(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, None, eps=1e-6), ResnetBlock2D(ch, ch, None, eps=1e-6)])
        self.attentions = nn.ModuleList([VaeAttention(ch)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class VaeEncBlock(nn.Module):
    def __init__(self, cin, cout, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin, cout, None, eps=1e-6), ResnetBlock2D(cout, cout, None, eps=1e-6)])
        if add_down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)])
        self.add_down = add_down

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.downsamplers[0](x) if self.add_down else x


class VaeDecBlock(nn.Module):
    def __init__(self, cin, cout, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, None, eps=1e-6) for i in range(3)])
        if add_up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.add_up = add_up

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.upsamplers[0](x) if self.add_up else x


class VaeEncoder(nn.Module):
    def __init__(self, chs=(128, 256, 512, 512)):
        super().__init__()
        self.conv_in = nn.Conv2d(3, chs[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cin = chs[0]
        for i, co in enumerate(chs):
            self.down_blocks.append(VaeEncBlock(cin, co, add_down=(i < len(chs) - 1)))
            cin = co
        self.mid_block = VaeMid(chs[-1])
        self.conv_norm_out = nn.GroupNorm(32, chs[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(chs[-1], 8, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class VaeDecoder(nn.Module):
    def __init__(self, chs=(128, 256, 512, 512)):
        super().__init__()
        rev = list(reversed(chs))
        self.conv_in = nn.Conv2d(4, rev[0], 3, padding=1)
        self.mid_block = VaeMid(rev[0])
        self.up_blocks = nn.ModuleList()
        cin = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(VaeDecBlock(cin, co, add_up=(i < len(rev) - 1)))
            cin = co
        self.conv_norm_out = nn.GroupNorm(32, rev[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(rev[-1], 3, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class AutoencoderKL(nn.Module):
    scaling_factor = 0.18215

    def __init__(self, chs=(128, 256, 512, 512)):
        super().__init__()
        self.encoder = VaeEncoder(chs)
        self.decoder = VaeDecoder(chs)
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)

    def encode_mean(self, x):
        return self.quant_conv(self.encoder(x))[:, :4]

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))


# ----------------------------------------------------------------------------- schedulers
class DDIMTables:
    """SD scheduler config: scaled_linear betas 0.00085->0.012, 1000 steps, steps_offset=1, leading spacing,
    set_alpha_to_one=False, clip_sample=False, epsilon prediction."""

    def __init__(self, num_train=1000, beta_start=0.00085, beta_end=0.012, steps_offset=1):
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.num_train = num_train
        self.steps_offset = steps_offset

    def timesteps(self, S: int) -> np.ndarray:
        ratio = self.num_train // S
        return (np.arange(0, S) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset

    def inverse_timesteps(self, S: int) -> np.ndarray:
        ratio = self.num_train // S
        return (np.arange(0, S) * ratio).round().copy().astype(np.int64) + self.steps_offset

    def step(self, eps, t: int, x, S: int):
        """DDIMScheduler.step, eta=0."""
        prev = t - self.num_train // S
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps

    def inverse_step(self, eps, t: int, x, S: int):
        """DDIMInverseScheduler.step: from alpha[t - ratio] (alpha[0] when negative) to alpha[t]."""
        cur = t - self.num_train // S
        a_t = self.alphas_cumprod[cur] if cur >= 0 else self.final_alpha_cumprod
        a_p = self.alphas_cumprod[t]
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps


def seeded_models(seed: int = 0, width_div: int = 1, cross_dim: int = 768, vae_chs=(128, 256, 512, 512),
                  with_vae: bool = True):
    """Synthetic-weight recipe of SURVEY §8d: PyTorch-default seeded init; ControlNet 'zero' convs re-initialised
    N(0, 0.02^2) so the residual path is exercised; conv_out scaled so eps has roughly unit variance."""
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    unet = UNet2DConditionModel(width_div, cross_dim)
    cnet = ControlNetModel(width_div, cross_dim)
    vae = AutoencoderKL(vae_chs) if with_vae else None
    with torch.no_grad():
        for conv in list(cnet.controlnet_down_blocks) + [cnet.controlnet_mid_block, cnet.controlnet_cond_embedding.conv_out]:
            conv.weight.normal_(0.0, 0.02, generator=g)
            conv.bias.zero_()
        unet.conv_out.weight.mul_(4.0)
    for m in [unet, cnet] + ([vae] if vae is not None else []):
        m.eval()
        for p in m.parameters():
            p.requires_grad_(False)
    return unet, cnet, vae
