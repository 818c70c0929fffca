"""ORACLE (test infrastructure only - never imported by gaussctrl_b200/): CPU restatement of the 3DGS fine-tune step
that follows the edit.

Follows, in the reference:
  * gaussctrl/gc_trainer.py:257-301  `train_iteration`: zero grads -> `pipeline.get_train_loss_dict(step)` ->
    `loss = sum(loss_dict.values())` -> `loss.backward()` -> Adam step of every parameter group -> scheduler step;
  * gaussctrl/gc_pipeline.py:276-287 `get_train_loss_dict`: `datamanager.next_train` -> model(camera) ->
    `model.get_loss_dict`;
  * gaussctrl/gc_config.py:57-89     the Adam groups (lr per group, eps 1e-15) and the exponential-decay scheduler of
    `xyz` (lr_final 1.6e-6, max_steps 30000).
Third-party pieces restated (not vendored in /root/reference, not installed here):
  * nerfstudio 1.0.0 `SplatfactoModel.get_loss_dict`: `main_loss = (1-l)*|gt-pred|.mean() + l*(1 - SSIM(gt, pred))`,
    l = ssim_lambda = 0.2, `SSIM(data_range=1.0, size_average=True, channel=3)`; scale regularisation is off by default;
  * pytorch_msssim (the `SSIM` module nerfstudio imports), published algorithm: 11-tap Gaussian window sigma 1.5
    normalised to sum 1, separable VALID (unpadded) depth-wise convolution along H then W, K = (0.01, 0.03),
    ssim_map = (2 mu1 mu2 + C1)/(mu1^2 + mu2^2 + C1) * (2 s12 + C2)/(s1 + s2 + C2), mean over pixels then channels;
  * nerfstudio `ExponentialDecayScheduler` without warm-up: lr(t) = exp(log(lr0)(1-t) + log(lr1) t), t = clip(step/max);
  * torch.optim.Adam is used as is (torch IS installed: that half of the oracle is the real implementation).
PARITY UNPINNED for the SSIM restatement (pytorch_msssim is not installable here); it is cross-checked in
tests/test_finetune_cpu.py against an independent dense 2-D window formulation and known answers (SSIM(x,x) = 1)."""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

WIN_SIZE, WIN_SIGMA = 11, 1.5
K1, K2 = 0.01, 0.03
SSIM_LAMBDA = 0.2

# gc_config.py:57-89 (camera_opt is not a Gaussian parameter group)
REFERENCE_LRS: Dict[str, float] = {"xyz": 1.6e-4, "features_dc": 0.0025, "features_rest": 0.0025 / 20, "opacity": 0.05,
                                   "scaling": 0.005, "rotation": 0.001}
GROUP_TO_PARAM = {"xyz": "means", "features_dc": "features_dc", "features_rest": "features_rest", "opacity": "opacities",
                  "scaling": "scales", "rotation": "quats"}
ADAM_EPS = 1e-15


def gauss_window() -> torch.Tensor:
    """pytorch_msssim._fspecial_gauss_1d(11, 1.5) in fp32."""
    coords = torch.arange(WIN_SIZE, dtype=torch.float32) - WIN_SIZE // 2
    g = torch.exp(-(coords ** 2) / (2 * WIN_SIGMA ** 2))
    return g / g.sum()


def gaussian_filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    """x [B,C,H,W]; valid separable depth-wise blur, H first then W (pytorch_msssim.gaussian_filter)."""
    C = x.shape[1]
    out = F.conv2d(x, win.reshape(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)
    return F.conv2d(out, win.reshape(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)


def ssim(X: torch.Tensor, Y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """X, Y [B,C,H,W] -> scalar (size_average=True)."""
    win = gauss_window().to(X.dtype)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    mu1, mu2 = gaussian_filter(X, win), gaussian_filter(Y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    sigma1_sq = gaussian_filter(X * X, win) - mu1_sq
    sigma2_sq = gaussian_filter(Y * Y, win) - mu2_sq
    sigma12 = gaussian_filter(X * Y, win) - mu1_mu2
    cs_map = (2 * sigma12 + C2) / (sigma1_sq + sigma2_sq + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1).mean()


def l1_ssim_loss(pred: torch.Tensor, gt: torch.Tensor, ssim_lambda: float = SSIM_LAMBDA):
    """pred, gt [H,W,3] in 0..1 -> (main_loss, L1, ssim) as nerfstudio's SplatfactoModel.get_loss_dict."""
    l1 = torch.abs(gt - pred).mean()
    s = ssim(gt.permute(2, 0, 1)[None], pred.permute(2, 0, 1)[None])
    return (1 - ssim_lambda) * l1 + ssim_lambda * (1 - s), l1, s


def exponential_decay_lr(step: int, lr_init: float, lr_final: float, max_steps: int) -> float:
    t = min(max(step / max_steps, 0.0), 1.0)
    return math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


def make_optimizers(params: Dict[str, torch.nn.Parameter], lrs: Dict[str, float] = REFERENCE_LRS):
    """One torch.optim.Adam per group, as nerfstudio's Optimizers does."""
    return {g: torch.optim.Adam([params[GROUP_TO_PARAM[g]]], lr=lr, eps=ADAM_EPS) for g, lr in lrs.items()}


def train_iteration(params: Dict[str, torch.nn.Parameter], optimizers, c2w, intr, H: int, W: int, gt: torch.Tensor,
                    background: torch.Tensor, step: int, sh_degree: int = 3) -> List[float]:
    """gc_trainer.py:257-301 for one camera with the oracle rasteriser (training-mode get_outputs: rgb + alpha only)."""
    from . import gsplat_ref as gr
    for o in optimizers.values():
        o.zero_grad()
    fx, fy, cx, cy = intr
    out = gr.get_outputs(params, c2w, fx, fy, cx, cy, H, W, sh_degree, background, training=True)
    loss, l1, s = l1_ssim_loss(out["rgb"], gt)
    loss.backward()
    optimizers["xyz"].param_groups[0]["lr"] = exponential_decay_lr(step, REFERENCE_LRS["xyz"], 1.6e-6, 30000)
    for o in optimizers.values():
        o.step()
    return [float(loss.detach()), float(l1.detach()), float(s.detach())]
