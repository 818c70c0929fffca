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


class _BinBuffers:
    """Per-(device, N, tile grid) buffers of the binning, sized for an intersection CAPACITY instead of the exact count so
    that nothing has to be read back before the kernels are enqueued.  The capacity is sticky and grows on overflow."""
    cache: dict = {}

    def __init__(self, N: int, tbx: int, tby: int, dev, cap: int):
        self.N, self.tbx, self.tby, self.cap = N, tbx, tby, cap
        self.gids = torch.empty((cap,), dtype=torch.int32, device=dev)
        self.bins = torch.empty((tbx * tby, 2), dtype=torch.int32, device=dev)
        self.count = torch.zeros((2,), dtype=torch.int32, device=dev)
        nb = lib.gcb_bin_gaussians_workspace_bytes(N, cap, tbx, tby)
        self.ws = torch.empty((nb,), dtype=torch.uint8, device=dev)

    @classmethod
    def get(cls, N: int, tbx: int, tby: int, dev, min_cap: int = 0) -> "_BinBuffers":
        key = (str(dev), N, tbx, tby, torch.cuda.current_stream().cuda_stream)
        buf = cls.cache.get(key)
        if buf is None or buf.cap < min_cap:
            cap = max(min_cap, 1 << 16, min(8 * N, (1 << 30) - 1)) if buf is None else max(min_cap, 2 * buf.cap)
            cap = min(cap, (1 << 30) - 1)
            buf = cls(N, tbx, tby, dev, cap)
            if len(cls.cache) > 8:
                cls.cache.clear()
            cls.cache[key] = buf
        return buf


PENDING_OVERFLOW: list = []   # device flags of deferred (sync-free) binnings, see check_deferred_overflow()


def _bin_launch(xys, depths, radii, num_tiles_hit, tbx: int, tby: int, buf: _BinBuffers, keys=None) -> None:
    check(lib.gcb_bin_gaussians(_p(xys), _p(depths), _p(radii), _p(num_tiles_hit), depths.shape[0], tbx, tby, buf.cap,
                                _p(buf.gids), _p(buf.bins), _p(buf.count), _p(keys), _p(buf.ws), buf.ws.numel(),
                                _stream()))
    ops.LAUNCHES[0] += 11


def bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds, want_keys: bool = False):
    """Depth-order the Gaussians, emit tile intersections, group by tile (gcb_bin_gaussians).
    -> (gaussian_ids [M] i32, tile_bins [T,2] i32, isect_keys [M] i64 or None, M).
    This seam-level helper returns exactly-sized tensors, so it reads M back ONCE, after everything is enqueued (gsplat's
    rasterize_gaussians has the same read in the middle of its pipeline); the eval render path does not read it at all."""
    xys, depths = _f32(xys), _f32(depths)
    N, dev = depths.shape[0], depths.device
    tbx, tby = int(tile_bounds[0]), int(tile_bounds[1])
    buf = _BinBuffers.get(N, tbx, tby, dev)
    while True:
        keys = torch.empty((buf.cap,), dtype=torch.int64, device=dev) if want_keys else None
        _bin_launch(xys, depths, radii, num_tiles_hit, tbx, tby, buf, keys)
        M, overflow = (int(v) for v in buf.count.tolist())
        if not overflow:
            break
        buf = _BinBuffers.get(N, tbx, tby, dev, min_cap=M)   # truncated: redo with room for all M intersections
    LAST_M[0] = M
    return buf.gids[:M].clone(), buf.bins.clone(), (keys[:M] if keys is not None else None), M


def check_deferred_overflow() -> None:
    """One synchronisation for ALL sync-free renders since the last call: raises if any of them overflowed its
    intersection capacity (the images it produced are truncated and must be re-rendered; the capacity has been grown)."""
    global PENDING_OVERFLOW
    pend, PENDING_OVERFLOW = PENDING_OVERFLOW, []
    if not pend:
        return
    torch.cuda.synchronize()
    worst = 0
    n_bad = 0
    for counts, key in pend:
        c = counts.reshape(-1, 2).cpu()
        LAST_M[0] = int(c[-1, 0])
        bad = c[c[:, 1] != 0]
        if bad.numel():
            m = int(bad[:, 0].max())
            n_bad += bad.shape[0]
            worst = max(worst, m)
            _grow_capacity(key, m)
    if n_bad:
        raise IsectOverflow(f"{n_bad} render(s) exceeded the intersection capacity (largest M = {worst}); capacity grown - "
                            f"render again")


def _grow_capacity(key, m: int) -> None:
    """key = (N, tbx, tby, dev) of a per-view binning or ("batch", N, H, W, dev) of a batched render."""
    if key and key[0] == "batch":
        _, N, H, W, dev = key
        for k in [k for k in _BatchBuffers.cache if k[0] == str(dev) and k[2:] == (N, H, W)]:
            del _BatchBuffers.cache[k]
        _BatchBuffers.min_cap = max(getattr(_BatchBuffers, "min_cap", 0), m)
    else:
        _BinBuffers.get(*key, min_cap=m)


class IsectOverflow(RuntimeError):
    pass


def render_views_multistream(render_one, items, streams):
    """Run `render_one(item)` for every item, item i on stream i % len(streams), and join them on the current stream.
    The binning kernels of one view are latency-bound (look-back chains, ~250-2000 blocks of dependent work) and leave
    most of the GPU idle; a second and third view on other streams fill it.  Every stream has its own binning workspace
    (_BinBuffers is keyed by stream); outputs are handed to the current stream with `record_stream`."""
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    outs = []
    for i, item in enumerate(items):
        s = streams[i % len(streams)]
        with torch.cuda.stream(s):
            out = render_one(item)
        for v in (out.values() if isinstance(out, dict) else out):
            if isinstance(v, torch.Tensor) and v.is_cuda:
                v.record_stream(cur)
        outs.append(out)
    for s in streams:
        cur.wait_stream(s)
    return outs


def rasterize_sorted(xys, conics, colors, opacity, gids, bins, img_height, img_width, background, radii=None):
    """-> (img [H,W,C], final_T [H,W], final_idx [H,W]).  `radii` (optional) only enables per-warp culling."""
    colors = _f32(colors)
    C = colors.shape[1]
    dev = colors.device
    H, W = int(img_height), int(img_width)
    out = torch.empty((H, W, C), dtype=torch.float32, device=dev)
    fT = torch.empty((H, W), dtype=torch.float32, device=dev)
    fidx = torch.empty((H, W), dtype=torch.int32, device=dev)
    bg = (ctypes.c_float * C)(*[float(v) for v in background.detach().cpu().reshape(-1).tolist()])
    check(lib.gcb_rasterize_fwd(_p(_f32(xys)), _p(_f32(conics)), _p(colors), _p(_f32(opacity).reshape(-1)), _p(gids),
                                _p(bins), _p(radii), H, W, C, bg, _p(out), _p(fT), _p(fidx), _stream()))
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
        img, fT, fidx = rasterize_sorted(xys, conics, colors, opacity, gids, bins, H, W, background, radii=radii)
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
                      background, defer_check: bool = False):
    """Eval-mode get_outputs in 3 stages on raw splatfacto parameters: fused project+SH+activations (one pass over the
    236 B/Gaussian record), binning (11 launches, intersection count stays on the device), fused rgb+depth composite with
    the clamp / depth-normalise epilogue.  No host synchronisation in between; `defer_check=True` also defers the
    capacity-overflow check to `check_deferred_overflow()` so that consecutive views pipeline on the GPU.
    -> (rgb [H,W,3], depth [H,W,1], alpha [H,W,1], xys, radii)."""
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
    buf = _BinBuffers.get(N, tbx, tby, dev)
    _bin_launch(xys, depths, radii, nth, tbx, tby, buf)
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    alpha = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    bg3 = background.detach().to(dev, torch.float32).reshape(3).contiguous()   # stays on the device: no sync
    check(lib.gcb_rasterize_rgbd_fwd(_p(xys), _p(conics), _p(rgbd), _p(opac), _p(buf.gids), _p(buf.bins), _p(radii), H, W,
                                     _p(bg3), _p(rgb), _p(depth), _p(alpha), _stream()))
    ops.LAUNCHES[0] += 1
    # M == 0 needs no special case: every tile run is empty, so rgb = background, depth = 1000, alpha = 0 - what the
    # reference returns for `radii.sum() == 0` plus the two keys it omits there (gc_model.py:155-156)
    PENDING_OVERFLOW.append((buf.count.clone(), (N, tbx, tby, dev)))
    if not defer_check:
        check_deferred_overflow()
    return rgb, depth, alpha, xys, radii


class _BatchBuffers:
    """Per-(device, stream, N, H, W) workspace of gcb_render_eval_batch, sized for a sticky intersection capacity."""
    cache: dict = {}

    @classmethod
    def get(cls, N: int, H: int, W: int, dev, min_cap: int = 0):
        key = (str(dev), torch.cuda.current_stream().cuda_stream, N, H, W)
        ent = cls.cache.get(key)
        if ent is None or ent[0] < min_cap:
            cap = max(min_cap, 1 << 16, min(8 * N, (1 << 30) - 1)) if ent is None else max(min_cap, 2 * ent[0])
            cap = min(cap, (1 << 30) - 1)
            nb = lib.gcb_render_eval_batch_workspace_bytes(N, cap, H, W)
            ent = (cap, torch.empty((nb,), dtype=torch.uint8, device=dev))
            if len(cls.cache) > 16:
                cls.cache.clear()
            cls.cache[key] = ent
        return ent


def render_eval_batch(params, cameras, img_height, img_width, sh_degree, background, streams=None):
    """Eval renders of V views in one C-ABI call per stream (gcb_render_eval_batch).  `cameras(i)` returns the host-side
    camera of view i as (viewmat [4,4], projmat = proj @ view [4,4], cam_origin [3], (fx, fy, cx, cy)); `cameras` may also
    be a tuple of four stacked tensors.  The views are split into len(streams) contiguous chunks that render concurrently;
    a chunk's cameras are evaluated right before its launch, so the host-side matrix math of chunk k+1 overlaps the GPU
    work of chunk k.  No host synchronisation; the capacity check is deferred (check_deferred_overflow()).
    -> rgb [V,H,W,3], depth [V,H,W,1], alpha [V,H,W,1]."""
    import numpy as np
    means = _f32(params["means"])
    N, dev = means.shape[0], means.device
    H, W = int(img_height), int(img_width)
    if callable(cameras):
        cam_fn, V = cameras, int(getattr(cameras, "n_views"))
    else:
        vm_t, pm_t, org_t, intr_t = cameras
        V = int(vm_t.shape[0])
        cam_fn = lambda i: (vm_t[i], pm_t[i], org_t[i], intr_t[i])  # noqa: E731
    tens = [_f32(params[k]) for k in ("scales", "quats", "features_dc")]
    rest = params.get("features_rest")
    rest = None if rest is None else _f32(rest)
    opl = _f32(params["opacities"]).reshape(-1)
    rgb = torch.empty((V, H, W, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((V, H, W, 1), dtype=torch.float32, device=dev)
    alpha = torch.empty((V, H, W, 1), dtype=torch.float32, device=dev)
    counts = torch.zeros((V, 2), dtype=torch.int32, device=dev)
    bg3 = background.detach().to(dev, torch.float32).reshape(3).contiguous()
    fp = ctypes.POINTER(ctypes.c_float)
    cur = torch.cuda.current_stream()
    lanes = list(streams) if streams else [cur]
    per = -(-V // len(lanes)) if V else 0
    px = H * W
    f32 = lambda t: np.ascontiguousarray(torch.as_tensor(t, dtype=torch.float32).detach().cpu().numpy(), dtype=np.float32)  # noqa: E731
    for li, st in enumerate(lanes):
        v0, v1 = li * per, min(V, (li + 1) * per)
        if v0 >= v1:
            break
        cams = [cam_fn(i) for i in range(v0, v1)]
        vm = np.stack([f32(c[0]).reshape(16) for c in cams])
        pm = np.stack([f32(c[1]).reshape(16) for c in cams])
        org = np.stack([f32(c[2]).reshape(3) for c in cams])
        intr = np.stack([f32(c[3]).reshape(4) for c in cams])
        if st is not cur:
            st.wait_stream(cur)
        with torch.cuda.stream(st):
            cap, ws = _BatchBuffers.get(N, H, W, dev, getattr(_BatchBuffers, "min_cap", 0))
            check(lib.gcb_render_eval_batch(
                _p(means), _p(tens[0]), _p(tens[1]), _p(tens[2]), _p(rest), _p(opl), N, int(sh_degree), v1 - v0,
                vm.ctypes.data_as(fp), pm.ctypes.data_as(fp), org.ctypes.data_as(fp), intr.ctypes.data_as(fp), H, W, _p(bg3),
                cap, _p(rgb, v0 * px * 3), _p(depth, v0 * px), _p(alpha, v0 * px), _p(counts, 2 * v0), _p(ws), ws.numel(),
                _stream()))
        ops.LAUNCHES[0] += 13 * (v1 - v0)
    for st in lanes:
        if st is not cur:
            cur.wait_stream(st)
    PENDING_OVERFLOW.append((counts, ("batch", N, H, W, dev)))
    return rgb, depth, alpha


def rasterize_rgbd(xys, depths, radii, conics, num_tiles_hit, rgbs, opacity, img_height, img_width, background):
    """Fused replacement of the reference's two rasterize calls + epilogue (gc_model.py:174-204):
    one binning, one 4-channel composite (r,g,b,depth), then rgb=min(rgb,1), depth=depth/alpha (1000 where alpha==0).
    -> (rgb [H,W,3], depth [H,W,1], alpha [H,W,1])"""
    dev = rgbs.device
    H, W = int(img_height), int(img_width)
    tile_bounds = ((W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK, 1)
    gids, bins, _, M = bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds)
    col4 = torch.cat([_f32(rgbs), _f32(depths)[:, None]], dim=1).contiguous()
    rgb = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
    depth = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    alpha = torch.empty((H, W, 1), dtype=torch.float32, device=dev)
    bg3 = background.detach().to(dev, torch.float32).reshape(3).contiguous()
    check(lib.gcb_rasterize_rgbd_fwd(_p(_f32(xys)), _p(_f32(conics)), _p(col4), _p(_f32(opacity).reshape(-1)), _p(gids),
                                     _p(bins), _p(radii), H, W, _p(bg3), _p(rgb), _p(depth), _p(alpha), _stream()))
    ops.LAUNCHES[0] += 1
    return rgb, depth, alpha
