"""ELEMENT-WISE parity at BASELINE.json configs[1] sizes: the oracle's own code (oracle/crossview_attn.py,
oracle/gsplat_ref.py) and a plain fp32 torch convolution are executed ON THE GPU in fp32 (TF32 off) - where they
finish in seconds - and compared element by element with the sm_100a kernels called through the C ABI.
(Properties at the same sizes: test_z_fullsize_properties_gpu.py; small-size CPU-oracle parity: test_kernels_gpu.py.)"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from gaussctrl_b200 import ops as _ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).half().cuda()


def _errs(got, want):
    got, want = got.float(), want.float()
    return ((got - want).norm() / (want.norm() + 1e-12)).item(), (got - want).abs().max().item()


@pytest.mark.parametrize("N,d,Bv", [(4096, 40, 3), (1024, 80, 3), (256, 160, 3), (64, 160, 3)])
def test_crossview_attention_elementwise_at_full_size(ops, N, d, Bv):
    """One cfg2 chunk's self-attention layer at each UNet level (B = 2 x 3 view rows, self + 4 cached reference sources,
    UNet weights 0.6 / 0.1 x4 and ControlNet weights 0 / 0.25 x4) against oracle.multi_source_attention in fp32."""
    from oracle import crossview_attn as cva
    heads, R = 8, 4
    C = heads * d
    qkv = _rand((2 * Bv, N, 3 * C), 11)
    ref = _rand((2 * R, N, 3 * C), 12)
    rows = [[h * Bv + f] + [-(h * R + r) - 1 for r in range(4)] for h in range(2) for f in range(Bv)]
    idx = torch.tensor(rows, dtype=torch.int32).cuda()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    ks = [k] + [torch.stack([ref[h * R + r, :, C:2 * C] for h in range(2) for _ in range(Bv)]) for r in range(4)]
    vs = [v] + [torch.stack([ref[h * R + r, :, 2 * C:] for h in range(2) for _ in range(Bv)]) for r in range(4)]
    for w in ([0.6, 0.1, 0.1, 0.1, 0.1], [0.0, 0.25, 0.25, 0.25, 0.25]):
        got = ops.attention(qkv, 0, 3 * C, qkv, C, 2 * C, 3 * C, ref, C, 2 * C, 3 * C, 2 * Bv, N, N, heads, d, idx, w)
        with torch.no_grad():
            want = cva.multi_source_attention(q, ks, vs, w, heads)
        torch.cuda.synchronize()
        rel, mx = _errs(got, want)
        # fp16 probabilities (P is rounded to fp16 before P V on the tensor core) and fp16 output: rel-RMS ~3e-4
        assert rel < 2e-3 and mx < 2e-3, (N, d, w[0], rel, mx)


@pytest.mark.parametrize("case", [(24, 64, 64, 320, 320, 3), (24, 64, 64, 320, 320, 1), (24, 32, 32, 640, 640, 3),
                                  (24, 16, 16, 1280, 1280, 3), (24, 8, 8, 2560, 1280, 3)])
def test_conv_elementwise_at_full_size(ops, case):
    """The convolutions of a 12-view batch (24 CFG rows; M = 98 304 at the 64x64 level) against fp32 F.conv2d."""
    B, H, W, Cin, Cout, k = case
    x = _rand((B, H, W, Cin), 1)
    w = _rand((Cout, k * k * Cin), 2, scale=1.0 / math.sqrt(k * k * Cin))
    bias = _rand((Cout,), 3)
    res = _rand((B, H, W, Cout), 4)
    y = ops.conv2d(x, w, bias, k, residual=res)
    wr = w.float().reshape(Cout, k, k, Cin).permute(0, 3, 1, 2)
    want = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wr, bias.float(), padding=k // 2)
    want = want.permute(0, 2, 3, 1) + res.float()
    torch.cuda.synchronize()
    rel, mx = _errs(y, want)
    assert rel < 1e-3 and mx < 2e-2, (case, rel, mx)   # one fp16 rounding of an fp32-accumulated value (|y| up to ~8)


def test_groupnorm_layernorm_elementwise_at_full_size(ops):
    B, H, W, C = 24, 64, 64, 320
    x = _rand((B, H, W, C), 5)
    gam, bet = _rand((C,), 6), _rand((C,), 7)
    y = ops.groupnorm(x, None, gam, bet, 32, 1e-5, True)
    want = torch.nn.functional.silu(torch.nn.functional.group_norm(x.float().permute(0, 3, 1, 2), 32, gam.float(),
                                                                   bet.float(), 1e-5)).permute(0, 2, 3, 1)
    rel, mx = _errs(y, want)
    assert rel < 1e-3, (rel, mx)
    t = x.reshape(B, H * W, C)
    y = ops.layernorm(t, gam, bet)
    want = torch.nn.functional.layer_norm(t.float(), (C,), gam.float(), bet.float(), 1e-5)
    rel, mx = _errs(y, want)
    assert rel < 1e-3, (rel, mx)


def test_raster_elementwise_at_one_million_gaussians():
    """The 1 M-Gaussian bench scene at 512x512, stage by stage against oracle/gsplat_ref.py: projection bit-exact (oracle
    on the CPU: its expression trees define the bits), tile binning exactly equal (integer work), compositing of
    (r,g,b,depth) and the get_outputs epilogue within 2e-5 (oracle's rasterize_sorted in fp32 on the GPU)."""
    from bench import orbit_c2w, synthetic_scene
    from gaussctrl_b200 import gsplat_ops as go
    from gaussctrl_b200.gc_model import render_gaussians
    from oracle import gsplat_ref as gr
    P = synthetic_scene(1_000_000, seed=0)
    c2w = torch.eye(4)
    c2w[:3] = orbit_c2w(3, 40)
    fx, fy, cx, cy, H, W = 539.05, 538.17, 258.74, 239.35, 512, 512
    vm = gr.viewmat_from_c2w(c2w)
    pm = gr.projection_matrix(0.001, 1000, 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fy)))
    tb = ((W + 15) // 16, (H + 15) // 16, 1)
    scales = torch.exp(P["scales"])
    quats = P["quats"] / P["quats"].norm(dim=-1, keepdim=True)
    with torch.no_grad():
        want = gr.project_gaussians(P["means"], scales, 1, quats, vm[:3], pm @ vm, fx, fy, cx, cy, H, W, tb)
        got = go.project_gaussians(P["means"].cuda(), scales.cuda(), 1, quats.cuda(), vm[:3], pm @ vm, fx, fy, cx, cy, H,
                                   W, tb)
    for name, a, b in zip(("xys", "depths", "radii", "conics", "num_tiles_hit"), got, want):
        assert torch.equal(a.cpu(), b), f"project_gaussians {name} not bit-exact at 1 M Gaussians"
    xys, depths, radii, conics, nth, _ = got
    # ---- binning: exact
    keys_w, gids_w, bins_w = gr.bin_and_sort_vectorized(*want[:3], want[4], tb)
    gids, bins, keys, M = go.bin_and_sort(xys, depths, radii, nth, tb, want_keys=True)
    assert M == len(gids_w) and M > 1_000_000
    assert np.array_equal(keys.cpu().numpy(), keys_w)
    assert np.array_equal(gids.cpu().numpy(), gids_w)
    assert np.array_equal(bins.cpu().numpy(), bins_w)
    # ---- compositing: oracle code on the GPU in fp32
    viewdirs = P["means"] - c2w[:3, 3]
    viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
    colors = torch.cat((P["features_dc"][:, None, :], P["features_rest"]), dim=1)
    rgbs = torch.clamp(gr.spherical_harmonics(3, viewdirs, colors) + 0.5, min=0.0)
    col4 = torch.cat([rgbs, want[1][:, None]], dim=1).cuda()
    opac = torch.sigmoid(P["opacities"]).cuda()
    bg4 = torch.zeros(4)
    with torch.no_grad():
        img_w, alpha_w, _ = gr.rasterize_sorted(want[0].cuda(), want[3].cuda(), col4, opac, gids_w, bins_w, H, W, bg4)
        img, fT, _ = go.rasterize_sorted(xys, conics, col4, opac, gids, bins, H, W, bg4)
    torch.cuda.synchronize()
    assert (img - img_w).abs()[..., :3].max().item() < 2e-5
    assert ((img - img_w).abs()[..., 3] / (1.0 + img_w[..., 3].abs())).max().item() < 2e-5   # depth channel, relative
    assert ((1.0 - fT) - alpha_w).abs().max().item() < 2e-5
    # ---- the fused eval path of GaussCtrlModel.get_outputs against the same oracle products
    out = render_gaussians({k_: v_.cuda() for k_, v_ in P.items()}, c2w, fx, fy, cx, cy, H, W, 3,
                           torch.zeros(3, device="cuda"))
    rgb_w = torch.clamp(img_w[..., :3], max=1.0)
    a_w = alpha_w[..., None]
    depth_w = torch.where(a_w > 0, img_w[..., 3:4] / torch.where(a_w > 0, a_w, torch.ones_like(a_w)),
                          torch.full_like(a_w, 1000.0))
    # the fused front end evaluates exp(scale) / SH with device intrinsics (1-2 ulp from torch's CPU results), so a
    # Gaussian's integer radius can differ by one pixel and with it one tile of a faint tail: bound the bulk tightly
    # and the worst pixel by the largest contribution such a tail can make (alpha at 3 sigma = exp(-4.5) ~ 0.011)
    d_rgb = (out["rgb"] - rgb_w).abs().amax(dim=-1)
    assert (d_rgb > 5e-5).float().mean().item() < 1e-3 and d_rgb.max().item() < 2e-2, (d_rgb.max().item(),)
    d_a = (out["accumulation"] - a_w).abs()
    assert (d_a > 2e-5).float().mean().item() < 1e-3 and d_a.max().item() < 2e-2
    hit = a_w[..., 0] > 1e-3
    d_dep = (out["depth"] - depth_w).abs()[..., 0][hit] / depth_w[..., 0][hit]
    assert (d_dep > 1e-4).float().mean().item() < 1e-3 and d_dep.max().item() < 5e-2
