"""CPU checks of the rasteriser oracle's helpers (test infrastructure for the -m gpu parity tests)."""
import math

import numpy as np
import torch

from oracle import gsplat_ref as gr


def _small_scene(n=3000, seed=2, H=96, W=128):
    g = torch.Generator().manual_seed(seed)
    means = torch.rand((n, 3), generator=g) * 2 - 1
    scales = torch.exp(torch.randn((n, 3), generator=g) * 0.4 + math.log(0.04))
    quats = torch.randn((n, 4), generator=g)
    quats = quats / quats.norm(dim=-1, keepdim=True)
    c2w = torch.eye(4)
    c2w[:3, 3] = torch.tensor([0.1, -0.2, 2.5])
    fx = fy = 110.0
    vm = gr.viewmat_from_c2w(c2w)
    pm = gr.projection_matrix(0.001, 1000, 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fy)))
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    out = gr.project_gaussians(means, scales, 1, quats, vm[:3], pm @ vm, fx, fy, W / 2, H / 2, H, W, tb)
    return out, tb, (H, W), g


def test_vectorized_binning_equals_loop():
    (xys, depths, radii, conics, nth, _), tb, _, _ = _small_scene()
    k1, g1, b1 = gr.bin_and_sort(xys, depths, radii, nth, tb)
    k2, g2, b2 = gr.bin_and_sort_vectorized(xys, depths, radii, nth, tb)
    assert len(k1) > 1000
    assert np.array_equal(k1, k2) and np.array_equal(g1, g2) and np.array_equal(b1, b2)


def test_rasterize_sorted_accepts_tensor_bins_and_background_device():
    (xys, depths, radii, conics, nth, _), tb, (H, W), g = _small_scene(n=800)
    _, gids, bins = gr.bin_and_sort(xys, depths, radii, nth, tb)
    col = torch.rand((xys.shape[0], 3), generator=g)
    op = torch.rand((xys.shape[0],), generator=g)
    a = gr.rasterize_sorted(xys, conics, col, op, gids, bins, H, W, torch.tensor([0.1, 0.2, 0.3]))
    b = gr.rasterize_sorted(xys, conics, col, op, gids, torch.from_numpy(bins), H, W, torch.tensor([0.1, 0.2, 0.3]))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and np.array_equal(a[2], b[2])
