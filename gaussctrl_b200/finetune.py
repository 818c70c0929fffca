"""3DGS fine-tune step after the edit, on the sm_100a kernels.

Host-side mirror of `GaussCtrlTrainer.train_iteration` (gaussctrl/gc_trainer.py:257-301) for the Gaussian parameter
groups of gaussctrl/gc_config.py:57-89:

    zero grads -> pipeline.get_train_loss_dict(step)  (gc_pipeline.py:276-287: datamanager.next_train -> model(camera)
    -> model.get_loss_dict) -> loss.backward() -> Adam step per group -> scheduler step

Below the seam: the training-mode render and its backward are `gsplat_ops` (raster.cu / raster_bwd.cu), the
`(1-l)*L1 + l*(1-SSIM)` loss and its gradient are one C-ABI call (`gcb_l1_ssim_loss_fwd_bwd`, finetune.cu), and all
six Adam groups update in one launch (`gcb_adam_step`).  No CPU path: CPU tensors raise."""
from __future__ import annotations

import ctypes
import math
import random
from typing import Dict, Iterable, List, Optional, Tuple

import torch

from . import ops
from ._lib import GcbError, check, lib

SSIM_LAMBDA = 0.2  # nerfstudio SplatfactoModelConfig.ssim_lambda
# gc_config.py:57-89: nerfstudio optimizer group -> (parameter name on the model, lr); Adam eps 1e-15 everywhere
REFERENCE_GROUPS: Dict[str, Tuple[str, float]] = {
    "xyz": ("means", 1.6e-4), "features_dc": ("features_dc", 0.0025), "features_rest": ("features_rest", 0.0025 / 20),
    "opacity": ("opacities", 0.05), "scaling": ("scales", 0.005), "rotation": ("quats", 0.001)}
ADAM_EPS = 1e-15
XYZ_LR_FINAL, XYZ_MAX_STEPS = 1.6e-6, 30000  # ExponentialDecaySchedulerConfig of "xyz" (gc_config.py:61-64)


def exponential_decay_lr(step: int, lr_init: float, lr_final: float, max_steps: int) -> float:
    """nerfstudio ExponentialDecayScheduler without warm-up."""
    t = min(max(step / max_steps, 0.0), 1.0)
    return math.exp(math.log(lr_init) * (1 - t) + math.log(lr_final) * t)


# ---------------------------------------------------------------------------------------------------------- loss
class _L1SSIMLoss(torch.autograd.Function):
    """Forward computes the loss AND d loss / d pred (one pass set); backward only scales it."""

    @staticmethod
    def forward(ctx, pred: torch.Tensor, gt: torch.Tensor, ssim_lambda: float):
        if not pred.is_cuda or not gt.is_cuda:
            raise GcbError("gaussctrl_b200 l1_ssim_loss needs CUDA tensors (there is no CPU path)")
        assert pred.shape == gt.shape and pred.dim() == 3, (pred.shape, gt.shape)
        p = pred.detach().to(torch.float32).contiguous()
        g = gt.detach().to(torch.float32).contiguous()
        H, W, C = p.shape
        nbytes = lib.gcb_l1_ssim_workspace_bytes(H, W, C)
        ws = torch.empty((max(nbytes, 4) // 4,), dtype=torch.float32, device=p.device)
        out = torch.empty((3,), dtype=torch.float32, device=p.device)
        v_pred = torch.empty_like(p)
        check(lib.gcb_l1_ssim_loss_fwd_bwd(ops._p(p), ops._p(g), H, W, C, float(ssim_lambda), ops._p(out), ops._p(v_pred),
                                           ops._p(ws), nbytes, ops._stream()))
        ops.LAUNCHES[0] += 5
        ctx.save_for_backward(v_pred)
        ctx.mark_non_differentiable(out)
        return out[0].clone(), out

    @staticmethod
    def backward(ctx, grad_loss, _grad_parts):
        (v_pred,) = ctx.saved_tensors
        return v_pred * grad_loss, None, None


def l1_ssim_loss(pred: torch.Tensor, gt: torch.Tensor, ssim_lambda: float = SSIM_LAMBDA):
    """pred, gt [H,W,3] fp32 in 0..1 -> (main_loss, parts) with parts = device float[3] (main_loss, L1, ssim);
    nerfstudio SplatfactoModel.get_loss_dict's `main_loss`.  Differentiable w.r.t. pred.  No host sync."""
    return _L1SSIMLoss.apply(pred, gt, ssim_lambda)


# ---------------------------------------------------------------------------------------------------------- Adam
class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) with every tensor of every group updated by
    `gcb_adam_step` - one launch per 8 tensors instead of ~10 eager kernels per tensor.  State keys (`step`, `exp_avg`,
    `exp_avg_sq`) are torch.optim.Adam's, so optimizer state_dicts of a nerfstudio checkpoint load unchanged."""

    def __init__(self, params, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        assert closure is None, "closures are not used by nerfstudio's Optimizers"
        batches: Dict[Tuple[int, float, float, float], List] = {}
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise GcbError("gaussctrl_b200 FusedAdam needs CUDA parameters (there is no CPU path)")
                assert p.dtype == torch.float32 and p.is_contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                key = (int(st["step"].item()), float(b1), float(b2), float(group["eps"]))
                batches.setdefault(key, []).append((p, g, st["exp_avg"], st["exp_avg_sq"], float(group["lr"])))
        for (step, b1, b2, eps), items in batches.items():
            n = len(items)
            VP = ctypes.c_void_p * n
            check(lib.gcb_adam_step(n, VP(*[i[0].data_ptr() for i in items]), VP(*[i[1].data_ptr() for i in items]),
                                    VP(*[i[2].data_ptr() for i in items]), VP(*[i[3].data_ptr() for i in items]),
                                    (ctypes.c_longlong * n)(*[i[0].numel() for i in items]),
                                    (ctypes.c_double * n)(*[i[4] for i in items]), b1, b2, eps, step, ops._stream()))
            ops.LAUNCHES[0] += (n + 7) // 8
        return None


def make_optimizer(model: torch.nn.Module, groups: Dict[str, Tuple[str, float]] = REFERENCE_GROUPS) -> FusedAdam:
    """One FusedAdam holding the reference's six groups (named, so the scheduler can find "xyz")."""
    pg = [{"params": [getattr(model, pname)], "lr": lr, "name": g} for g, (pname, lr) in groups.items()]
    return FusedAdam(pg, eps=ADAM_EPS)


# ---------------------------------------------------------------------------------------------------- train step
class FineTuner:
    """`train_iteration` of the reference trainer for a GaussCtrlModel + a datamanager with `next_train(step)`."""

    def __init__(self, model: torch.nn.Module, datamanager, groups: Dict[str, Tuple[str, float]] = REFERENCE_GROUPS):
        self.model = model
        self.datamanager = datamanager
        self.groups = groups
        self.optimizer = make_optimizer(model, groups)

    def get_train_loss_dict(self, step: int):
        """gc_pipeline.py:276-287."""
        camera, batch = self.datamanager.next_train(step)
        self.model.train()
        outputs = self.model.get_outputs(camera)
        metrics_dict = self.model.get_metrics_dict(outputs, batch)
        loss_dict = self.model.get_loss_dict(outputs, batch, metrics_dict)
        return outputs, loss_dict, metrics_dict

    def train_iteration(self, step: int):
        """gc_trainer.py:257-301 (mixed_precision=False: no autocast, grad scaler is the identity)."""
        self.optimizer.zero_grad(set_to_none=True)
        _, loss_dict, metrics_dict = self.get_train_loss_dict(step)
        loss = sum(v for v in loss_dict.values() if torch.is_tensor(v))
        loss.backward()
        for group in self.optimizer.param_groups:
            if group.get("name") == "xyz":
                group["lr"] = exponential_decay_lr(step, self.groups["xyz"][1], XYZ_LR_FINAL, XYZ_MAX_STEPS)
        self.optimizer.step()
        return loss, loss_dict, metrics_dict


def next_train_view(unseen: List[int], n_views: int) -> int:
    """View choice of GaussCtrlDataManager.next_train (gc_datamanager.py:213-223): pop a random unseen camera (python
    `random`, inclusive randint), refill when exhausted.  `unseen` is mutated in place."""
    idx = unseen.pop(random.randint(0, len(unseen) - 1))
    if len(unseen) == 0:
        unseen.extend(range(n_views))
    return idx
