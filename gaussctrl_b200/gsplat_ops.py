"""The gsplat operator seam on the sm_100a rasteriser kernels.

Same call signatures as gsplat 0.1.3's `project_gaussians`, `spherical_harmonics`, `rasterize_gaussians` as the
reference calls them (gaussctrl/gc_model.py:140-154, :166, :174-186, :191-202), plus `rasterize_rgbd`, the fused
rgb+depth+alpha pass `GaussCtrlModel.get_outputs` uses here instead of two rasterize calls."""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import ops
from ._lib import check, lib

BLOCK = 16
LAST_M = [0]        # intersections of the most recent binning (bench.py reports it with the raster roofline)
FUSED_EVAL = True  # GaussCtrlModel eval renders use the fused project+SH front end (gc_model.render_gaussians)
_p = ops._p
_stream = ops._stream


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("gaussctrl_b200 rasteriser needs CUDA tensors (there is no CPU path)")
    return t.detach().to(torch.float32).contiguous()


def _host16(m: torch.Tensor):
    m = m.detach().to("cpu", torch.float32)
    if m.shape[0] == 3:
        m = torch.cat([m, torch.tensor([[0.0, 0.0, 0.0, 1.0]])], dim=0)
    return (ctypes.c_float * 16)(*m.reshape(-1).tolist())


def _project_gaussians_fwd(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
                           tile_bounds, clip_thresh: float = 0.01):
    """-> (xys [N,2], depths [N], radii [N] i32, conics [N,3], num_tiles_hit [N] i32, cov3d [N,6])."""
    means3d, scales, quats = _f32(means3d), _f32(scales), _f32(quats)
    N, dev = means3d.shape[0], means3d.device
    xys = torch.empty((N, 2), dtype=torch.float32, device=dev)
    depths = torch.empty((N,), dtype=torch.float32, device=dev)
    radii = torch.empty((N,), dtype=torch.int32, device=dev)
    conics = torch.empty((N, 3), dtype=torch.float32, device=dev)
    nth = torch.empty((N,), dtype=torch.int32, device=dev)
    cov3d = torch.empty((N, 6), dtype=torch.float32, device=dev)
    check(lib.gcb_project_gaussians_fwd(_p(means3d), _p(scales), float(glob_scale), _p(quats), _host16(viewmat),
                                        _host16(projmat), float(fx), float(fy), float(cx), float(cy), int(img_height),
                                        int(img_width), int(tile_bounds[0]), int(tile_bounds[1]), float(clip_thresh), N,
                                        _p(xys), _p(depths), _p(radii), _p(conics), _p(nth), _p(cov3d), _stream()))
    ops.LAUNCHES[0] += 1
    return xys, depths, radii, conics, nth, cov3d


def _spherical_harmonics_fwd(degrees_to_use: int, viewdirs, coeffs):
    viewdirs, coeffs = _f32(viewdirs), _f32(coeffs)
    N, K = coeffs.shape[0], coeffs.shape[1]
    colors = torch.empty((N, 3), dtype=torch.float32, device=coeffs.device)
    check(lib.gcb_sh_fwd(int(degrees_to_use), K, _p(viewdirs), _p(coeffs), _p(colors), N, _stream()))
    ops.LAUNCHES[0] += 1
    return colors


def bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds, want_keys: bool = False):
    """Depth-order the Gaussians, emit tile intersections, group by tile.
    -> (gaussian_ids [M] i32, tile_bins [T,2] i32, isect_keys [M] i64 or None, M)."""
    N, dev = depths.shape[0], depths.device
    tbx, tby = int(tile_bounds[0]), int(tile_bounds[1])
    sorted_ids = torch.empty((N,), dtype=torch.int32, device=dev)
    cum = torch.empty((N,), dtype=torch.int32, device=dev)
    nb = lib.gcb_depth_order_workspace_bytes(N)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    check(lib.gcb_depth_order(_p(depths), _p(num_tiles_hit), N, _p(sorted_ids), _p(cum), _p(ws), nb, _stream()))
    ops.LAUNCHES[0] += 16
    M = int(cum[-1].item())  # the one host sync gsplat's rasterize_gaussians also has
    LAST_M[0] = M
    gids = torch.empty((max(M, 1),), dtype=torch.int32, device=dev)
    bins = torch.empty((tbx * tby, 2), dtype=torch.int32, device=dev)
    keys = torch.empty((max(M, 1),), dtype=torch.int64, device=dev) if want_keys else None
    nb2 = lib.gcb_bin_tiles_workspace_bytes(N, M, tbx, tby)
    ws2 = torch.empty(max(nb2, 16), dtype=torch.uint8, device=dev)
    check(lib.gcb_bin_tiles(_p(xys), _p(depths), _p(radii), _p(sorted_ids), _p(cum), N, M, tbx, tby, _p(gids), _p(bins),
                            _p(keys), _p(ws2), ws2.numel(), _stream()))
    ops.LAUNCHES[0] += 6
    return gids[:M], bins, (keys[:M] if keys is not None else None), M


def rasterize_sorted(xys, conics, colors, opacity, gids, bins, img_height, img_width, background):
    """-> (img [H,W,C], final_T [H,W], final_idx [H,W])"""
    colors = _f32(colors)
    C = colors.shape[1]
    dev = colors.device
    H, W = int(img_height), int(img_width)
    out = torch.empty((H, W, C), dtype=torch.float32, device=dev)
    fT = torch.empty((H, W), dtype=torch.float32, device=dev)
    fidx = torch.empty((H, W), dtype=torch.int32, device=dev)
    bg = (ctypes.c_float * C)(*[float(v) for v in background.detach().cpu().reshape(-1).tolist()])
    check(lib.gcb_rasterize_fwd(_p(_f32(xys)), _p(_f32(conics)), _p(colors), _p(_f32(opacity).reshape(-1)), _p(gids),
                                _p(bins), H, W, C, bg, _p(out), _p(fT), _p(fidx), _stream()))
    ops.LAUNCHES[0] += 1
    return out, fT, fidx


class _ProjectGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
                tile_bounds, clip_thresh):
        out = _project_gaussians_fwd(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height,
                                     img_width, tile_bounds, clip_thresh)
        xys, depths, radii, conics, nth, cov3d = out
        ctx.save_for_backward(_f32(means3d), _f32(scales), _f32(quats), radii)
        ctx.consts = (float(glob_scale), _host16(viewmat), _host16(projmat), float(fx), float(fy), float(cx), float(cy),
                      int(img_height), int(img_width))
        ctx.mark_non_differentiable(radii, nth, cov3d)
        return out

    @staticmethod
    def backward(ctx, v_xys, v_depths, v_radii, v_conics, v_nth, v_cov3d):
        means3d, scales, quats, radii = ctx.saved_tensors
        gs, vm, pm, fx, fy, cx, cy, H, W = ctx.consts
        N, dev = means3d.shape[0], means3d.device
        z = lambda t, shp: torch.zeros(shp, dtype=torch.float32, device=dev) if t is None else _f32(t)  # noqa: E731
        v_xys, v_depths, v_conics = z(v_xys, (N, 2)), z(v_depths, (N,)), z(v_conics, (N, 3))
        v_means = torch.empty((N, 3), dtype=torch.float32, device=dev)
        v_scales = torch.empty((N, 3), dtype=torch.float32, device=dev)
        v_quats = torch.empty((N, 4), dtype=torch.float32, device=dev)
        check(lib.gcb_project_gaussians_bwd(_p(means3d), _p(scales), gs, _p(quats), vm, pm, fx, fy, cx, cy, H, W,
                                            _p(radii), _p(v_xys), _p(v_depths), _p(v_conics), N, _p(v_means),
                                            _p(v_scales), _p(v_quats), _stream()))
        ops.LAUNCHES[0] += 1
        return (v_means, v_scales, None, v_quats) + (None,) * 10


class _SphericalHarmonics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, degree, viewdirs, coeffs):
        ctx.degree, ctx.K = int(degree), coeffs.shape[1]
        ctx.save_for_backward(_f32(viewdirs))
        return _spherical_harmonics_fwd(degree, viewdirs, coeffs)

    @staticmethod
    def backward(ctx, v_colors):
        (viewdirs,) = ctx.saved_tensors
        N = viewdirs.shape[0]
        v_coeffs = torch.empty((N, ctx.K, 3), dtype=torch.float32, device=viewdirs.device)
        check(lib.gcb_sh_bwd(ctx.degree, ctx.K, _p(viewdirs), _p(_f32(v_colors)), _p(v_coeffs), N, _stream()))
        ops.LAUNCHES[0] += 1
        return None, None, v_coeffs  # view directions come from detached means (gc_model.py:163)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width, background):
        H, W = int(img_height), int(img_width)
        tile_bounds = ((W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK, 1)
        gids, bins, _, M = bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds)
        img, fT, fidx = rasterize_sorted(xys, conics, colors, opacity, gids, bins, H, W, background)
        ctx.save_for_backward(_f32(xys), _f32(conics), _f32(colors), _f32(opacity).reshape(-1), gids, bins, fT, fidx,
                              background.detach().to("cpu", torch.float32))
        ctx.size = (H, W)
        return img, 1.0 - fT

    @staticmethod
    def backward(ctx, v_img, v_alpha):
        xys, conics, colors, opac, gids, bins, fT, fidx, bg = ctx.saved_tensors
        H, W = ctx.size
        N, C, dev = xys.shape[0], colors.shape[1], xys.device
        v_xy = torch.zeros((N, 2), dtype=torch.float32, device=dev)
        v_conic = torch.zeros((N, 3), dtype=torch.float32, device=dev)
        v_colors = torch.zeros((N, C), dtype=torch.float32, device=dev)
        v_opac = torch.zeros((N,), dtype=torch.float32, device=dev)
        bgc = (ctypes.c_float * C)(*[float(v) for v in bg.reshape(-1).tolist()])
        check(lib.gcb_rasterize_bwd(_p(xys), _p(conics), _p(colors), _p(opac), _p(gids), _p(bins), H, W, C, bgc, _p(fT),
                                    _p(fidx), _p(_f32(v_img)), _p(None if v_alpha is None else _f32(v_alpha)), _p(v_xy),
                                    _p(v_conic), _p(v_colors), _p(v_opac), _stream()))
        ops.LAUNCHES[0] += 1
        return v_xy, None, None, v_conic, None, v_colors, v_opac[:, None], None, None, None


def project_gaussians(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
                      tile_bounds, clip_thresh: float = 0.01):
    """gsplat 0.1.3 `project_gaussians` (differentiable w.r.t. means3d, scales, quats)."""
    return _ProjectGaussians.apply(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height,
                                   img_width, tile_bounds, clip_thresh)


def spherical_harmonics(degrees_to_use: int, viewdirs, coeffs):
    """gsplat 0.1.3 `spherical_harmonics` (differentiable w.r.t. coeffs)."""
    return _SphericalHarmonics.apply(degrees_to_use, viewdirs, coeffs)


def rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                        background: Optional[torch.Tensor] = None, return_alpha: bool = False):
    """gsplat 0.1.3 signature; differentiable w.r.t. xys, conics, colors, opacity.  colors [N,C], C in {1,3,4}."""
    C = colors.shape[-1]
    dev = colors.device
    if background is None:
        background = torch.ones(C, dtype=torch.float32, device=dev)
    if int(num_tiles_hit.sum().item()) < 1:
        img = torch.ones(img_height, img_width, C, device=dev) * background.to(dev)
        return (img, torch.zeros(img_height, img_width, device=dev)) if return_alpha else img
    opacity = opacity if opacity.dim() == 2 else opacity[:, None]
    img, alpha = _RasterizeGaussians.apply(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height,
                                           img_width, background)
    return (img, alpha) if return_alpha else img


def render_eval_fused(params, viewmat, projmat, cam_origin, fx, fy, cx, cy, img_height, img_width, sh_degree,
                      background):
    """Eval-mode get_outputs in 3 stages on raw splatfacto parameters: fused project+SH+activations (one pass over the
    236 B/Gaussian record), binning, fused rgb+depth composite + epilogue.
    -> (rgb [H,W,3], depth [H,W,1], alpha [H,W,1], xys, radii) or None when nothing is visible."""
    means = _f32(params["means"])
    N, dev = means.shape[0], means.device
    H, W = int(img_height), int(img_width)
    tbx, tby = (W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK
    xys = torch.empty((N, 2), dtype=torch.float32, device=dev)
    depths = torch.empty((N,), dtype=torch.float32, device=dev)
    radii = torch.empty((N,), dtype=torch.int32, device=dev)
    conics = torch.empty((N, 3), dtype=torch.float32, device=dev)
    nth = torch.empty((N,), dtype=torch.int32, device=dev)
    rgbd = torch.empty((N, 4), dtype=torch.float32, device=dev)
    opac = torch.empty((N,), dtype=torch.float32, device=dev)
    org = (ctypes.c_float * 3)(*[float(v) for v in cam_origin])
    rest = params.get("features_rest")
    check(lib.gcb_project_sh_fused_fwd(_p(means), _p(_f32(params["scales"])), _p(_f32(params["quats"])),
                                       _p(_f32(params["features_dc"])), _p(None if rest is None else _f32(rest)),
                                       _p(_f32(params["opacities"]).reshape(-1)), _host16(viewmat), _host16(projmat), org,
                                       float(fx), float(fy), float(cx), float(cy), H, W, tbx, tby, int(sh_degree), N,
                                       _p(xys), _p(depths), _p(radii), _p(conics), _p(nth), _p(rgbd), _p(opac), _stream()))
    ops.LAUNCHES[0] += 1
    gids, bins, _, M = bin_and_sort(xys, depths, radii, nth, (tbx, tby, 1))
    if M < 1:
        return None
    bg4 = torch.cat([background.detach().to("cpu", torch.float32).reshape(3), torch.zeros(1)])
    img4, fT, _ = rasterize_sorted(xys, conics, rgbd, opac, gids, bins, H, W, bg4)
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    alpha = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    check(lib.gcb_raster_finalize(_p(img4), _p(fT), _p(rgb), _p(depth), _p(alpha), H * W, _stream()))
    ops.LAUNCHES[0] += 1
    return rgb, depth, alpha, xys, radii


def rasterize_rgbd(xys, depths, radii, conics, num_tiles_hit, rgbs, opacity, img_height, img_width, background):
    """Fused replacement of the reference's two rasterize calls + epilogue (gc_model.py:174-204):
    one binning, one 4-channel composite (r,g,b,depth), then rgb=min(rgb,1), depth=depth/alpha (1000 where alpha==0).
    -> (rgb [H,W,3], depth [H,W,1], alpha [H,W,1])"""
    dev = rgbs.device
    H, W = int(img_height), int(img_width)
    tile_bounds = ((W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK, 1)
    gids, bins, _, M = bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds)
    col4 = torch.cat([_f32(rgbs), _f32(depths)[:, None]], dim=1).contiguous()
    bg4 = torch.cat([background.detach().to(dev, torch.float32).reshape(3), torch.zeros(1, device=dev)])
    img4, fT, _ = rasterize_sorted(xys, conics, col4, opacity, gids, bins, H, W, bg4)
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    alpha = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    check(lib.gcb_raster_finalize(_p(img4), _p(fT), _p(rgb), _p(depth), _p(alpha), H * W, _stream()))
    ops.LAUNCHES[0] += 1
    return rgb, depth, alpha
