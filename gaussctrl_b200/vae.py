"""AutoencoderKL (SD1.x VAE) encode-mean / decode on the sm_100a kernels, channels-last fp16.

Replaces `self.pipe.vae.encode(image).latent_dist.mean * 0.18215` (gaussctrl/gc_pipeline.py:239-246) and the
`vae.decode(latents / scaling_factor)` + `(x/2+0.5).clamp(0,1)` that `pipe(..., output_type='pt')` runs at the end
of every edit call (gc_pipeline.py:209-219).  The mid-block attention (1 head, d=512, N=(H/8)^2 tokens) is two
tcgen05 GEMMs around a row-softmax kernel."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from ._lib import GCB_ACT_NONE
from .diffusion import PackedNet

SCALING = 0.18215


class VaeB200:
    def __init__(self, sd: Dict[str, torch.Tensor], device):
        self.dev = torch.device(device)
        self.net = PackedNet(sd, self.dev, heads=1)

    # ------------------------------------------------------------------------------------------ blocks
    def _resnet(self, p: str, x: torch.Tensor) -> torch.Tensor:
        n = self.net
        h = ops.groupnorm(x, None, n.vec(p + ".norm1.weight"), n.vec(p + ".norm1.bias"), 32, 1e-6, True)
        w, b, _ = n.conv(p + ".conv1")
        h = ops.conv2d(h, w, b, 3)
        h = ops.groupnorm(h, None, n.vec(p + ".norm2.weight"), n.vec(p + ".norm2.bias"), 32, 1e-6, True)
        if n.has(p + ".conv_shortcut"):
            ws, bs, _ = n.conv(p + ".conv_shortcut")
            x = ops.conv2d(x, ws, bs, 1)
        w, b, _ = n.conv(p + ".conv2")
        return ops.conv2d(h, w, b, 3, residual=x)

    def _attention(self, p: str, x: torch.Tensor) -> torch.Tensor:
        n = self.net
        B, H, W, C = x.shape
        N = H * W
        y = ops.groupnorm(x, None, n.vec(p + ".group_norm.weight"), n.vec(p + ".group_norm.bias"), 32, 1e-6, False)
        wqkv, bqkv = n.cat_lin([p + ".to_q", p + ".to_k", p + ".to_v"])
        qkv = ops.linear(y.reshape(B, N, C), wqkv, bqkv)  # [B,N,3C]
        wo, bo = n.lin(p + ".to_out.0")
        out = torch.empty((B, N, C), dtype=torch.float16, device=x.device)
        xr = x.reshape(B, N, C)
        for i in range(B):
            q = qkv[i, :, :C].contiguous()
            k = qkv[i, :, C:2 * C].contiguous()
            vt = ops.transpose(qkv[i, :, 2 * C:].contiguous().reshape(1, N, C))[0]  # [C,N]
            s = ops.linear(q, k)                       # [N,N] = q k^T
            pr = ops.softmax_rows(s, C ** -0.5)
            o = ops.linear(pr, vt)                     # [N,C]
            out[i] = ops.linear(o, wo, bo, residual=xr[i].contiguous())
        return out.reshape(B, H, W, C)

    def _mid(self, p: str, x: torch.Tensor) -> torch.Tensor:
        x = self._resnet(p + ".resnets.0", x)
        x = self._attention(p + ".attentions.0", x)
        return self._resnet(p + ".resnets.1", x)

    # ------------------------------------------------------------------------------------------ public
    @torch.no_grad()
    def decode(self, z_nhwc: torch.Tensor) -> torch.Tensor:
        """z [B,h,w,4] fp16 (already divided by the scaling factor) -> decoder output [B,8h,8w,3] fp16 (pre-clamp)."""
        n = self.net
        w, b, _ = n.conv("post_quant_conv")
        x = ops.conv2d_direct(z_nhwc, w, b, 1, 1, (0, 0))
        w, b, _ = n.conv("decoder.conv_in")
        x = ops.conv2d_direct(x, w, b, 3, 1, (1, 1))
        x = self._mid("decoder.mid_block", x)
        for i in range(4):
            for j in range(3):
                x = self._resnet(f"decoder.up_blocks.{i}.resnets.{j}", x)
            if i < 3:
                w, b, _ = n.conv(f"decoder.up_blocks.{i}.upsamplers.0.conv")
                x = ops.conv2d(ops.upsample_nearest2x(x), w, b, 3)
        x = ops.groupnorm(x, None, n.vec("decoder.conv_norm_out.weight"), n.vec("decoder.conv_norm_out.bias"), 32, 1e-6,
                          True)
        # conv_out (128 -> 3 at full resolution): on the tensor-core GEMM with the output channels zero-padded to 8
        # (the SIMT direct convolution took 1.4 ms per 512^2 view: 30 % of the decode, profiles/r3l_vae_launches.csv)
        w8, b8 = n.conv_pad_out("decoder.conv_out", 8)
        return ops.conv2d(x, w8, b8, 3)[..., :3].contiguous()

    @torch.no_grad()
    def decode_latents(self, latents_nchw: torch.Tensor, mask: Optional[torch.Tensor] = None,
                       unedited: Optional[torch.Tensor] = None, batch: int = 4) -> torch.Tensor:
        """Final latents [V,4,h,w] -> edited images [V,H,W,3] fp32 in [0,1] (+ optional mask composite,
        gc_pipeline.py:223-234)."""
        outs = []
        for i in range(0, latents_nchw.shape[0], batch):
            z = latents_nchw[i:i + batch].to(self.dev, torch.float16)
            z = ops.nchw_to_nhwc((z.float() / SCALING).half().contiguous())
            img = self.decode(z)
            m = None if mask is None else mask[i:i + batch].to(self.dev, torch.float32).contiguous()
            u = None if unedited is None else unedited[i:i + batch].to(self.dev, torch.float16).contiguous()
            outs.append(ops.postprocess_composite(img, m, u))
        return torch.cat(outs, dim=0)

    @torch.no_grad()
    def encode_mean(self, img_nhwc: torch.Tensor) -> torch.Tensor:
        """image2latent (gc_pipeline.py:239-246): img [B,H,W,3] in 0..1 -> latents [B,4,h,w] fp16 = mean * 0.18215."""
        n = self.net
        x = (img_nhwc.to(self.dev, torch.float16) * 2 - 1).contiguous()
        w, b, _ = n.conv("encoder.conv_in")
        x = ops.conv2d_direct(x, w, b, 3, 1, (1, 1))
        for i in range(4):
            for j in range(2):
                x = self._resnet(f"encoder.down_blocks.{i}.resnets.{j}", x)
            if i < 3:
                w, b, _ = n.conv(f"encoder.down_blocks.{i}.downsamplers.0.conv")
                x = ops.conv3x3_s2(x, w, b, (0, 1))
        x = self._mid("encoder.mid_block", x)
        x = ops.groupnorm(x, None, n.vec("encoder.conv_norm_out.weight"), n.vec("encoder.conv_norm_out.bias"), 32, 1e-6,
                          True)
        w, b, _ = n.conv("encoder.conv_out")   # 512 -> 8: tensor-core GEMM
        x = ops.conv2d(x, w, b, 3)
        w, b, _ = n.conv("quant_conv")
        x = ops.conv2d_direct(x, w, b, 1, 1, (0, 0))
        mean = ops.nhwc_to_nchw(x)[:, :4]
        return (mean.float() * SCALING).half()
