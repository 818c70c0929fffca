"""CPU tests of the fine-tune-step oracle (oracle/finetune.py) and of the pass decomposition the CUDA loss kernel uses
(gaussctrl_b200/csrc/finetune.cu): the SSIM restatement against an independent dense 2-D formulation and known
answers, the analytic partials of the four-pass backward against autograd, the learning-rate schedule and the
view-sampling of `next_train`."""
import math
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import finetune as oft


def _images(H, W, seed):
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand((H, W, 3), generator=g, dtype=torch.float64)
    gt = F.avg_pool2d(gt.permute(2, 0, 1)[None], 5, stride=1, padding=2)[0].permute(1, 2, 0).contiguous()
    pred = (gt + 0.1 * torch.randn((H, W, 3), generator=g, dtype=torch.float64)).clamp(0, 1)
    return pred, gt


def test_window_sums_to_one_and_is_symmetric():
    w = oft.gauss_window()
    assert w.shape == (11,) and abs(float(w.sum()) - 1.0) < 1e-6
    assert torch.allclose(w, w.flip(0)) and int(w.argmax()) == 5
    assert abs(float(w[5] / w[4]) - math.exp(1 / 4.5)) < 1e-5


def test_ssim_known_answers():
    pred, gt = _images(32, 40, 0)
    x = gt.permute(2, 0, 1)[None]
    assert abs(float(oft.ssim(x, x)) - 1.0) < 1e-12
    y = pred.permute(2, 0, 1)[None]
    s = float(oft.ssim(x, y))
    assert 0.0 < s < 1.0 and abs(s - float(oft.ssim(y, x))) < 1e-12  # symmetric
    # constant images: mu = c, sigma = 0 -> ssim = (2 c1 c2 + C1) / (c1^2 + c2^2 + C1)
    a, b = torch.full_like(x, 0.3), torch.full_like(x, 0.5)
    want = (2 * 0.3 * 0.5 + 1e-4) / (0.09 + 0.25 + 1e-4)
    assert abs(float(oft.ssim(a, b)) - want) < 1e-5  # the fp32 window sums to 1 only within 1e-7


def test_ssim_equals_dense_window_formulation():
    """Independent restatement: one dense 11x11 window (outer product), unfold instead of conv2d."""
    pred, gt = _images(24, 29, 1)
    X, Y = gt.permute(2, 0, 1), pred.permute(2, 0, 1)
    w1 = oft.gauss_window().double()
    w2 = (w1[:, None] * w1[None, :]).reshape(-1)

    def blur(img):  # [3,H,W] -> [3,Ho,Wo]
        patches = img.unfold(1, 11, 1).unfold(2, 11, 1).reshape(3, img.shape[1] - 10, img.shape[2] - 10, 121)
        return (patches * w2).sum(-1)

    mu1, mu2 = blur(X), blur(Y)
    s1, s2, s12 = blur(X * X) - mu1 ** 2, blur(Y * Y) - mu2 ** 2, blur(X * Y) - mu1 * mu2
    m = ((2 * mu1 * mu2 + 1e-4) / (mu1 ** 2 + mu2 ** 2 + 1e-4)) * ((2 * s12 + 9e-4) / (s1 + s2 + 9e-4))
    assert abs(float(m.mean()) - float(oft.ssim(X[None], Y[None]))) < 1e-12


def _kernel_plan_numpy(pred, gt, lam):
    """Literal numpy emulation of the five passes of finetune.cu (same index arithmetic, fp64)."""
    H, W, C = pred.shape
    w = oft.gauss_window().double().numpy()
    Wo, Ho = W - 10, H - 10
    prods = [pred, gt, pred * pred, gt * gt, pred * gt]
    hb = [sum(w[k] * a[:, k:k + Wo, :] for k in range(11)) for a in prods]                       # pass 1
    mp, mg, epp, egg, epg = [sum(w[k] * a[k:k + Ho] for k in range(11)) for a in hb]             # pass 2
    sp, sg, spg = epp - mp * mp, egg - mg * mg, epg - mp * mg
    A1, A2 = 2 * mp * mg + 1e-4, 2 * spg + 9e-4
    B1, B2 = mp * mp + mg * mg + 1e-4, sp + sg + 9e-4
    s = A1 * A2 / (B1 * B2)
    dm = [2 * mg * (A2 - A1) / (B1 * B2) - 2 * mp * s * (1 / B1 - 1 / B2), -s / B2, 2 * A1 / (B1 * B2)]
    tb = []
    for d in dm:                                                                                  # pass 3
        t = np.zeros((H, Wo, C))
        for y in range(H):
            for k in range(11):
                if 0 <= y - k < Ho:
                    t[y] += w[k] * d[y - k]
        tb.append(t)
    a = []
    for t in tb:                                                                                  # pass 4
        o = np.zeros((H, W, C))
        for x in range(W):
            for k in range(11):
                if 0 <= x - k < Wo:
                    o[:, x] += w[k] * t[:, x - k]
        a.append(o)
    n_map, n_img = Ho * Wo * C, H * W * C
    grad = (-lam / n_map) * (a[0] + 2 * pred * a[1] + gt * a[2]) + ((1 - lam) / n_img) * np.sign(pred - gt)
    l1, ssim = np.abs(pred - gt).mean(), s.mean()
    return (1 - lam) * l1 + lam * (1 - ssim), l1, ssim, grad


@pytest.mark.parametrize("H,W", [(16, 23), (11, 11), (31, 12)])
def test_kernel_pass_decomposition_matches_autograd(H, W):
    pred, gt = _images(H, W, 2)
    p = pred.clone().requires_grad_(True)
    loss, l1, s = oft.l1_ssim_loss(p, gt, 0.2)
    loss.backward()
    got_loss, got_l1, got_s, got_grad = _kernel_plan_numpy(pred.numpy(), gt.numpy(), 0.2)
    assert abs(got_loss - float(loss)) < 1e-12 and abs(got_l1 - float(l1)) < 1e-12 and abs(got_s - float(s)) < 1e-12
    assert np.abs(got_grad - p.grad.numpy()).max() < 1e-12 * max(1.0, np.abs(got_grad).max() * 1e3)


def test_exponential_decay_schedule():
    f = oft.exponential_decay_lr
    assert abs(f(0, 1.6e-4, 1.6e-6, 30000) - 1.6e-4) < 1e-12
    assert abs(f(15000, 1.6e-4, 1.6e-6, 30000) - 1.6e-5) < 1e-12   # geometric midpoint
    assert abs(f(30000, 1.6e-4, 1.6e-6, 30000) - 1.6e-6) < 1e-15   # the fine-tune starts here (ckpt step 29999 + 1)
    assert abs(f(30499, 1.6e-4, 1.6e-6, 30000) - 1.6e-6) < 1e-15
    from gaussctrl_b200 import finetune as ft
    for s in (0, 1234, 30000, 31000):
        assert ft.exponential_decay_lr(s, 1.6e-4, 1.6e-6, 30000) == f(s, 1.6e-4, 1.6e-6, 30000)
    assert {g: lr for g, (_, lr) in ft.REFERENCE_GROUPS.items()} == oft.REFERENCE_LRS
    assert {g: p for g, (p, _) in ft.REFERENCE_GROUPS.items()} == oft.GROUP_TO_PARAM


def test_next_train_view_visits_every_view_once_per_epoch():
    from gaussctrl_b200.finetune import next_train_view
    random.seed(3)
    unseen = list(range(7))
    first = [next_train_view(unseen, 7) for _ in range(7)]
    assert sorted(first) == list(range(7)) and unseen == list(range(7))   # refilled after the last pop
    second = [next_train_view(unseen, 7) for _ in range(7)]
    assert sorted(second) == list(range(7))


def test_fused_adam_refuses_cpu_parameters():
    from gaussctrl_b200 import _lib
    from gaussctrl_b200.finetune import FusedAdam, l1_ssim_loss
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(_lib.GcbError):
        FusedAdam([p]).step()
    with pytest.raises(_lib.GcbError):
        l1_ssim_loss(torch.zeros(16, 16, 3), torch.zeros(16, 16, 3))
    lib = _lib.lib
    assert lib.gcb_l1_ssim_workspace_bytes(512, 512, 3) == 4 * (5 * 512 * 502 * 3 + 3 * 502 * 502 * 3 + 3 * 512 * 502 * 3 +
                                                                 -(-502 * 502 * 3 // 256) + -(-512 * 512 * 3 // 256))
    assert lib.gcb_l1_ssim_workspace_bytes(8, 512, 3) == 0
    assert lib.gcb_l1_ssim_loss_fwd_bwd(None, None, 8, 8, 3, 0.2, None, None, None, 0, None) == -1
