"""SD1.x UNet + ControlNet(depth) denoiser on the sm_100a kernels (channels-last fp16, fp32 accumulate).

Replaces what `self.pipe(...)` runs per DDIM step in the reference (gaussctrl/gc_pipeline.py:142-145 inversion,
:209-219 editing): `ControlNetModel.forward` -> `UNet2DConditionModel.forward` with every `Attention` module going
through `CrossViewAttnProcessor` (gaussctrl/utils.py:44-133).  Differences of schedule that keep the arithmetic:
  * q/k/v projections of a self-attention are one GEMM (to_q|to_k|to_v have no bias and share their input);
  * the 5 attention passes of utils.py:88-117 are ONE multi-source kernel launch, probabilities stay on chip;
  * text K/V (attn2.to_k/to_v of the prompt embeddings) and the ControlNet conditioning embedding do not depend on
    the timestep, so they are computed once instead of every step;
  * all `time_emb_proj` linears of a network are one GEMM per step;
  * ControlNet residual adds (`down_block_res_samples + controlnet residuals`) are the residual epilogue of the
    ControlNet 1x1 "zero" convolutions;
  * skip-connection `torch.cat`s are never materialised (GroupNorm reads two tensors, conv_shortcut is two GEMMs).
Weights arrive as a diffusers-keyed state_dict (sd15_spec)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import GCB_ACT_GEGLU, GCB_ACT_NONE, GCB_ACT_SILU, GCB_GEMM_TCGEN05
from .sd15_spec import skip_channels

import os as _os

_FUSED_GATHER = _os.environ.get("GCB_FUSED_GATHER", "1") != "0"   # A/B switch: projection GEMM + exchange as one kernel


@dataclass
class AttnPlan:
    """How self-attention picks its K/V sources for one network call.

    src_index [B, n_src] int32 (device): row of (k, v) for values >= 0 - the current batch - or row -(v+1) of the
    recorded reference K/V (`ref_kv[layer]`, a fused qkv tensor [R2, N, 3C]) for negative values.
    weights: per-source blend weights; UNet and ControlNet use different ones (self_attn_coeff 0.6 / 0,
    gc_pipeline.py:163-168)."""
    src_index: torch.Tensor
    weights_unet: Sequence[float]
    weights_cnet: Sequence[float]
    ref_kv: Optional[Dict[str, torch.Tensor]] = None       # layer name -> qkv tensor of the reference rows
    record_kv: Optional[Dict[str, torch.Tensor]] = None    # if set, every self-attn layer stores its qkv tensor here
    text_index: Optional[torch.Tensor] = None              # [B,1] int32: which prompt embedding each row uses
    gather: Optional[object] = None                        # callable(layer, qkv_local) -> qkv of ALL reference rows


def vanilla_plan(B: int, device, text_index: Optional[torch.Tensor] = None) -> AttnPlan:
    idx = torch.arange(B, dtype=torch.int32, device=device).reshape(B, 1)
    if text_index is None:
        text_index = torch.zeros((B, 1), dtype=torch.int32, device=device)
    return AttnPlan(idx, [1.0], [1.0], text_index=text_index)


def literal_crossview_plan(F: int, device, ref_frames: Sequence[int] = (0, 1, 2, 3), coeff_unet: float = 0.6,
                           coeff_cnet: float = 0.0, record_kv=None) -> AttnPlan:
    """The reference's batch layout (gc_pipeline.py:206-219, utils.py:94-109): 2F rows = [uncond x F | cond x F], every
    row attends to itself and to frames `ref_frames` of its own CFG half."""
    rows = []
    for half in range(2):
        for f in range(F):
            rows.append([half * F + f] + [half * F + r for r in ref_frames])
    idx = torch.tensor(rows, dtype=torch.int32, device=device)
    K = len(ref_frames)
    text_index = torch.tensor([[0]] * F + [[1]] * F, dtype=torch.int32, device=device)
    return AttnPlan(idx, [coeff_unet] + [(1 - coeff_unet) / K] * K, [coeff_cnet] + [(1 - coeff_cnet) / K] * K,
                    record_kv=record_kv, text_index=text_index)


def cached_crossview_plan(Bv: int, R: int, device, ref_kv, ref_frames: Sequence[int] = (0, 1, 2, 3),
                          coeff_unet: float = 0.6, coeff_cnet: float = 0.0) -> AttnPlan:
    """Views-only batch [uncond x Bv | cond x Bv]; reference K/V come from the recorded reference pass whose rows are
    [uncond x R | cond x R]."""
    rows = []
    for half in range(2):
        for _ in range(Bv):
            rows.append([-1] + [-(half * R + r) - 1 for r in ref_frames])
    idx = torch.tensor(rows, dtype=torch.int32, device=device)
    idx[:, 0] = torch.arange(2 * Bv, dtype=torch.int32, device=device)
    K = len(ref_frames)
    text_index = torch.tensor([[0]] * Bv + [[1]] * Bv, dtype=torch.int32, device=device)
    return AttnPlan(idx, [coeff_unet] + [(1 - coeff_unet) / K] * K, [coeff_cnet] + [(1 - coeff_cnet) / K] * K,
                    ref_kv=ref_kv, text_index=text_index)


class PackedNet:
    """Device-resident fp16 weights of one network in the layouts the kernels want."""

    def __init__(self, sd: Dict[str, torch.Tensor], device, heads: int = 8):
        self.sd = sd
        self.dev = device
        self.heads = heads
        self._cache: Dict[str, object] = {}

    def _t(self, name: str) -> torch.Tensor:
        return self.sd[name].to(device=self.dev, dtype=torch.float16).contiguous()

    def has(self, name: str) -> bool:
        return (name + ".weight") in self.sd

    def vec(self, name: str) -> torch.Tensor:
        if name not in self._cache:
            self._cache[name] = self._t(name)
        return self._cache[name]

    def conv(self, name: str):
        """-> (w [Cout, k*k*Cin] OHWI, bias, ksize)"""
        key = "conv:" + name
        if key not in self._cache:
            w = self.sd[name + ".weight"]
            k = w.shape[-1]
            wp = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(device=self.dev, dtype=torch.float16).contiguous()
            self._cache[key] = (wp, self._t(name + ".bias"), k)
        return self._cache[key]

    def conv_pad_out(self, name: str, cout: int):
        """conv weights with the output channels zero-padded to `cout` -> (w [cout, k*k*Cin], bias [cout])"""
        key = "convpad:" + name
        if key not in self._cache:
            w, b, _ = self.conv(name)
            wp = torch.zeros((cout, w.shape[1]), dtype=torch.float16, device=self.dev)
            bp = torch.zeros((cout,), dtype=torch.float16, device=self.dev)
            wp[: w.shape[0]] = w
            bp[: b.shape[0]] = b
            self._cache[key] = (wp, bp)
        return self._cache[key]

    def conv_c4(self, name: str):
        """3x3 conv over 4 input channels as a K = 40 GEMM (ops.conv3x3_c4) -> (w [Cout, 40], bias)"""
        key = "convc4:" + name
        if key not in self._cache:
            w, b, _ = self.conv(name)
            assert w.shape[1] == 36, w.shape
            wp = torch.zeros((w.shape[0], 40), dtype=torch.float16, device=self.dev)
            wp[:, :36] = w
            self._cache[key] = (wp, b)
        return self._cache[key]

    def conv_split(self, name: str, c1: int):
        """1x1 conv over a channel concat, split into the two K ranges -> (w1, w2, bias)"""
        key = "convsplit:" + name
        if key not in self._cache:
            w = self.sd[name + ".weight"].reshape(self.sd[name + ".weight"].shape[0], -1)
            w1 = w[:, :c1].to(device=self.dev, dtype=torch.float16).contiguous()
            w2 = w[:, c1:].to(device=self.dev, dtype=torch.float16).contiguous()
            self._cache[key] = (w1, w2, self._t(name + ".bias"))
        return self._cache[key]

    def lin(self, name: str):
        key = "lin:" + name
        if key not in self._cache:
            b = self._t(name + ".bias") if (name + ".bias") in self.sd else None
            self._cache[key] = (self._t(name + ".weight"), b)
        return self._cache[key]

    def cat_lin(self, names: Sequence[str]):
        key = "cat:" + "|".join(names)
        if key not in self._cache:
            w = torch.cat([self.sd[n + ".weight"] for n in names], dim=0)
            b = None
            if (names[0] + ".bias") in self.sd:
                b = torch.cat([self.sd[n + ".bias"] for n in names], dim=0).to(device=self.dev, dtype=torch.float16)
            self._cache[key] = (w.to(device=self.dev, dtype=torch.float16).contiguous(), b)
        return self._cache[key]

    def qkv_ones_padded(self, attn: str, heads: int, pad_to: int):
        """Fused [to_q | to_k | to_v] projection whose V block gives every head `pad_to` columns: the head's d value
        columns, then a column that is identically 1.0 (zero weight row + bias 1), then zeros.  The tcgen05 attention
        kernel reads the softmax row sums out of that ones column (gcb_attn_multi_fwd, v_head_stride)."""
        key = f"qkvpad:{attn}:{pad_to}"
        if key not in self._cache:
            wq, wk, wv = (self.sd[f"{attn}.{n}.weight"] for n in ("to_q", "to_k", "to_v"))
            C = wq.shape[0]
            d = C // heads
            wvp = torch.zeros((heads, pad_to, wv.shape[1]), dtype=wv.dtype)
            wvp[:, :d] = wv.reshape(heads, d, -1)
            w = torch.cat([wq, wk, wvp.reshape(heads * pad_to, -1)], dim=0)
            b = torch.zeros((w.shape[0],), dtype=torch.float32)
            b[2 * C + d::pad_to] = 1.0
            self._cache[key] = (w.to(device=self.dev, dtype=torch.float16).contiguous(),
                                b.to(device=self.dev, dtype=torch.float16))
        return self._cache[key]

    def geglu(self, name: str):
        """GEGLU projection with rows interleaved per N tile for the fused epilogue."""
        key = "geglu:" + name
        if key not in self._cache:
            w, b = self.lin(name)
            perm = ops.geglu_perm(w.shape[0], self.dev)
            self._cache[key] = (w[perm].contiguous(), b[perm].contiguous())
        return self._cache[key]


class SD15Denoiser:
    """ControlNet + UNet noise prediction for a batch of latents (one DDIM step's network evaluation)."""

    def __init__(self, unet_sd: Dict[str, torch.Tensor], cnet_sd: Dict[str, torch.Tensor], device, heads: int = 8,
                 fuse_geglu: bool = True):
        self.dev = torch.device(device)
        self.unet = PackedNet(unet_sd, self.dev, heads)
        self.cnet = PackedNet(cnet_sd, self.dev, heads)
        self.heads = heads
        self.fuse_geglu = fuse_geglu
        # head dim 40: V heads padded to 48 columns with a ones column, so the softmax row sums come out of the P V
        # product and the attention kernel may evaluate part of its exponentials as packed-half polynomials
        # (attn_tc.cu ex2_hpoly: 530 -> 620-634 TFLOP/s on B200, profiles/r3_attn.md)
        self.ones_column = True
        self.ch = [unet_sd[f"down_blocks.{i}.resnets.0.conv1.weight"].shape[0] for i in range(4)]
        self._temb_layout: Dict[int, Tuple[List[str], Dict[str, int], int]] = {}
        self.text_kv: Dict[Tuple[int, str, int], torch.Tensor] = {}

    # ---------------------------------------------------------------------------------------- per-run constants
    def set_prompts(self, embeds: torch.Tensor) -> None:
        """embeds [P,77,768] fp16: computes attn2 K/V of every transformer block once (they are step-invariant)."""
        embeds = embeds.to(device=self.dev, dtype=torch.float16).contiguous()
        self.n_prompts = embeds.shape[0]
        self.text_len = embeds.shape[1]
        for net_id, net in ((0, self.unet), (1, self.cnet)):
            for name in sorted({k[: k.index(".attn2.") + 6] for k in net.sd if ".attn2.to_k.weight" in k}):
                w, _ = net.cat_lin([name + ".to_k", name + ".to_v"])
                kv = ops.linear(embeds, w)  # [P,77,2C]
                key = (net_id, name, self.n_prompts)
                old = self.text_kv.get(key)
                if old is not None and old.shape == kv.shape:
                    old.copy_(kv)  # captured CUDA graphs hold the address of this buffer
                else:
                    self.text_kv[key] = kv

    def controlnet_cond(self, cond_nhwc: torch.Tensor) -> torch.Tensor:
        """controlnet_cond_embedding (3->16->16->32->32->96->96->256->C0 convs with SiLU): [B,512,512,3] -> [B,64,64,C0].
        Step-invariant, so callers evaluate it once per view."""
        n = self.cnet
        e = "controlnet_cond_embedding"
        w, b, _ = n.conv(e + ".conv_in")
        x = ops.conv2d_direct(cond_nhwc, w, b, 3, 1, (1, 1), GCB_ACT_SILU)
        for i in range(6):
            w, b, _ = n.conv(f"{e}.blocks.{i}")
            x = ops.conv2d_direct(x, w, b, 3, 2 if i % 2 == 1 else 1, (1, 1), GCB_ACT_SILU)
        w, b, _ = n.conv(e + ".conv_out")
        return ops.conv2d_direct(x, w, b, 3, 1, (1, 1), GCB_ACT_NONE)

    # ---------------------------------------------------------------------------------------- building blocks
    def _temb(self, net: PackedNet, net_id: int, t_dev: torch.Tensor):
        """time_embedding MLP + every resnet's time_emb_proj(SiLU(temb)) in one GEMM -> (proj [B, total], offsets)."""
        if net_id not in self._temb_layout:
            names = [k[: -len(".time_emb_proj.weight")] for k in net.sd if k.endswith(".time_emb_proj.weight")]
            offs, o = {}, 0
            for nme in names:
                offs[nme] = o
                o += net.sd[nme + ".time_emb_proj.weight"].shape[0]
            self._temb_layout[net_id] = (names, offs, o)
        names, offs, total = self._temb_layout[net_id]
        e = ops.timestep_embedding(t_dev, self.ch[0])
        w1, b1 = net.lin("time_embedding.linear_1")
        w2, b2 = net.lin("time_embedding.linear_2")
        h = ops.linear(e, w1, b1, act=GCB_ACT_SILU)
        # SiLU(temb) is what every ResnetBlock2D feeds to time_emb_proj: fuse it in this GEMM's epilogue
        temb_act = ops.linear(h, w2, b2, act=GCB_ACT_SILU)
        wp, bp = net.cat_lin([nme + ".time_emb_proj" for nme in names])
        return ops.linear(temb_act, wp, bp), offs, total

    def _resnet(self, net: PackedNet, p: str, x: torch.Tensor, x2: Optional[torch.Tensor], tproj) -> torch.Tensor:
        proj, offs, total = tproj
        h = ops.groupnorm(x, x2, net.vec(p + ".norm1.weight"), net.vec(p + ".norm1.bias"), 32, 1e-5, True)
        w, b, _ = net.conv(p + ".conv1")
        h = ops.conv2d(h, w, b, 3, rowvec=proj, rowvec_off=offs[p], rowvec_ld=total)
        h = ops.groupnorm(h, None, net.vec(p + ".norm2.weight"), net.vec(p + ".norm2.bias"), 32, 1e-5, True)
        if net.has(p + ".conv_shortcut"):
            if x2 is None:
                ws, bs, _ = net.conv(p + ".conv_shortcut")
                sc = ops.conv2d(x, ws, bs, 1)
            else:
                w1, w2, bs = net.conv_split(p + ".conv_shortcut", x.shape[-1])
                sc = ops.conv2d(x, w1, bs, 1)
                sc = ops.conv2d(x2, w2, None, 1, residual=sc)
        else:
            assert x2 is None
            sc = x
        w, b, _ = net.conv(p + ".conv2")
        return ops.conv2d(h, w, b, 3, residual=sc)

    def _transformer(self, net: PackedNet, net_id: int, p: str, x: torch.Tensor, plan: AttnPlan) -> torch.Tensor:
        B, H, W, C = x.shape
        N = H * W
        heads, d = self.heads, C // self.heads
        h = ops.groupnorm(x, None, net.vec(p + ".norm.weight"), net.vec(p + ".norm.bias"), 32, 1e-6, False)
        w, b, _ = net.conv(p + ".proj_in")
        h = ops.conv2d(h, w, b, 1).reshape(B, N, C)
        blk = p + ".transformer_blocks.0"
        # --- attn1: (cross-view) self-attention
        n1 = ops.layernorm(h, net.vec(blk + ".norm1.weight"), net.vec(blk + ".norm1.bias"))
        if d == 40 and self.ones_column:
            # head dim 40: V heads padded to 48 columns with a ones column -> row sums come out of the P V product
            wqkv, bqkv = net.qkv_ones_padded(blk + ".attn1", heads, 48)   # [2C + heads*48, C]
            ld, vstride = 2 * C + heads * 48, 48
        else:
            wqkv, _ = net.cat_lin([blk + ".attn1.to_q", blk + ".attn1.to_k", blk + ".attn1.to_v"])
            bqkv = None
            ld, vstride = 3 * C, d
        qkv = None
        layer = f"{net_id}:{blk}.attn1"
        kv2 = None
        if plan.gather is not None:
            # sharded reference pass (parallel.py): every rank needs the q|k|v rows of ALL reference rows
            fused = None
            if qkv is None and _FUSED_GATHER and hasattr(plan.gather, "linear_gather"):
                # one kernel: the GEMM's epilogue stores every output tile into all ranks' K/V buffers over NVLink
                # (gcb_linear_allgather_fwd); the local rows are read back out of the gathered buffer
                fused = plan.gather.linear_gather(layer, n1, wqkv, bqkv, C)   # the Q third (C columns) stays local
            if fused is not None:
                kv2, qkv = fused
                ops.LAUNCHES[0] += 2
            else:
                if qkv is None:
                    qkv = ops.linear(n1, wqkv, bqkv)  # [B,N,ld]
                kv2 = plan.gather(layer, qkv)   # GEMM, then push + flag kernels (or NCCL)
            if plan.record_kv is not None:
                plan.record_kv[layer] = kv2
        else:
            if qkv is None:
                qkv = ops.linear(n1, wqkv, bqkv)  # [B,N,ld]
            if plan.record_kv is not None:
                plan.record_kv[layer] = qkv
            kv2 = plan.ref_kv[layer] if plan.ref_kv is not None else None
        weights = plan.weights_unet if net_id == 0 else plan.weights_cnet
        a = ops.attention(qkv, 0, ld, qkv, C, 2 * C, ld, kv2, C, 2 * C, ld, B, N, N, heads, d, plan.src_index, weights,
                          v_head_stride=vstride)
        wo, bo = net.lin(blk + ".attn1.to_out.0")
        h = ops.linear(a, wo, bo, residual=h)
        # --- attn2: text cross-attention (K/V precomputed by set_prompts)
        n2 = ops.layernorm(h, net.vec(blk + ".norm2.weight"), net.vec(blk + ".norm2.bias"))
        wq, _ = net.lin(blk + ".attn2.to_q")
        q = ops.linear(n2, wq)
        tkv = self.text_kv[(net_id, blk + ".attn2", self.n_prompts)]
        a = ops.attention(q, 0, C, tkv, 0, C, 2 * C, None, 0, 0, 0, B, N, self.text_len, heads, d, plan.text_index, [1.0])
        wo, bo = net.lin(blk + ".attn2.to_out.0")
        h = ops.linear(a, wo, bo, residual=h)
        # --- GEGLU feed-forward
        n3 = ops.layernorm(h, net.vec(blk + ".norm3.weight"), net.vec(blk + ".norm3.bias"))
        if self.fuse_geglu and ops._GEMM_IMPL[0] == GCB_GEMM_TCGEN05:
            wg, bg = net.geglu(blk + ".ff.net.0.proj")
            ff = ops.linear(n3, wg, bg, act=GCB_ACT_GEGLU)
        else:
            wg, bg = net.lin(blk + ".ff.net.0.proj")
            ff = ops.geglu(ops.linear(n3, wg, bg))
        w2, b2 = net.lin(blk + ".ff.net.2")
        h = ops.linear(ff, w2, b2, residual=h)
        w, b, _ = net.conv(p + ".proj_out")
        return ops.conv2d(h.reshape(B, H, W, C), w, b, 1, residual=x)

    def _encoder(self, net: PackedNet, net_id: int, x: torch.Tensor, tproj, plan: AttnPlan):
        """conv_in output -> (mid output, skip list): the 4 down blocks + mid block shared by UNet and ControlNet."""
        skips = [x]
        for i in range(4):
            for j in range(2):
                x = self._resnet(net, f"down_blocks.{i}.resnets.{j}", x, None, tproj)
                if i < 3:
                    x = self._transformer(net, net_id, f"down_blocks.{i}.attentions.{j}", x, plan)
                skips.append(x)
            if i < 3:
                w, b, _ = net.conv(f"down_blocks.{i}.downsamplers.0.conv")
                x = ops.conv3x3_s2(x, w, b, (1, 1))
                skips.append(x)
        x = self._resnet(net, "mid_block.resnets.0", x, None, tproj)
        x = self._transformer(net, net_id, "mid_block.attentions.0", x, plan)
        x = self._resnet(net, "mid_block.resnets.1", x, None, tproj)
        return x, skips

    # ---------------------------------------------------------------------------------------- one network evaluation
    def eps(self, x: torch.Tensor, t_dev: torch.Tensor, cond_emb: torch.Tensor, plan: AttnPlan) -> torch.Tensor:
        """x [B,h,w,4] fp16 NHWC latents, t_dev [B] fp32 device timesteps, cond_emb [B,h,w,C0] = controlnet_cond()
        -> eps [B,h,w,4].  controlnet_conditioning_scale = 1.0 (gc_pipeline.py:118,216)."""
        cn, un = self.cnet, self.unet
        # ---- ControlNet
        tp_c = self._temb(cn, 1, t_dev)
        col = ops.im2col3x3_c4(x)   # the 3x3 patches of the latents, shared by both conv_in
        w, b = cn.conv_c4("conv_in")
        xc = ops.conv3x3_c4(x, w, b, residual=cond_emb, col=col)
        c_mid, c_skips = self._encoder(cn, 1, xc, tp_c, plan)
        # ---- UNet down + mid
        tp_u = self._temb(un, 0, t_dev)
        w, b = un.conv_c4("conv_in")
        xu = ops.conv3x3_c4(x, w, b, col=col)
        u_mid, u_skips = self._encoder(un, 0, xu, tp_u, plan)
        # ---- add the ControlNet residuals: zero-conv GEMM with the UNet tensor as residual epilogue
        skips = []
        for i, (cs, us) in enumerate(zip(c_skips, u_skips)):
            w, b, _ = cn.conv(f"controlnet_down_blocks.{i}")
            skips.append(ops.conv2d(cs, w, b, 1, residual=us))
        w, b, _ = cn.conv("controlnet_mid_block")
        h = ops.conv2d(c_mid, w, b, 1, residual=u_mid)
        # ---- UNet up
        for i in range(4):
            for j in range(3):
                h = self._resnet(un, f"up_blocks.{i}.resnets.{j}", h, skips.pop(), tp_u)
                if i > 0:
                    h = self._transformer(un, 0, f"up_blocks.{i}.attentions.{j}", h, plan)
            if i < 3:
                w, b, _ = un.conv(f"up_blocks.{i}.upsamplers.0.conv")
                h = ops.conv2d(ops.upsample_nearest2x(h), w, b, 3)
        h = ops.groupnorm(h, None, un.vec("conv_norm_out.weight"), un.vec("conv_norm_out.bias"), 32, 1e-5, True)
        # conv_out (C0 -> 4): run on the tensor-core GEMM with the 4 output channels zero-padded to 8
        w8, b8 = un.conv_pad_out("conv_out", 8)
        return ops.conv2d(h, w8, b8, 3)[..., :4].contiguous()
