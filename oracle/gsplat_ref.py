"""Oracle: 3D-Gaussian projection / SH / tile binning / alpha compositing (TEST INFRASTRUCTURE).

PARITY UNPINNED.  The algorithm lives in gsplat==0.1.2|0.1.3 (README.md:59-60), an un-vendored CUDA dependency
that is absent from /root/reference and not installable here.  This file restates gsplat 0.1.3's published
algorithm (project_gaussians_forward_kernel, compute_sh_forward_kernel, map_gaussian_to_intersects,
get_tile_bin_edges, rasterize_forward) in fp32 PyTorch, anchored on the reference's call sites:
    gc_model.py:140-154  project_gaussians(means, scales, 1, quats, viewmat[:3], projmat@viewmat, fx,fy,cx,cy,H,W,tile_bounds)
    gc_model.py:162-167  spherical_harmonics(n, viewdirs, colors)
    gc_model.py:174-186  rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, rgbs, opac, H, W, background, return_alpha)
    gc_model.py:191-204  second rasterize on depths, depth/alpha, 1000 where alpha == 0

Conventions fixed by decree (SURVEY §8c open items; recorded in DESIGN.md):
  * ndc2pix(x, W, c) = 0.5*W*x + c - 0.5 and the centre of pixel (i, j) is (px, py) = (j, i)   [gsplat 0.1.x]
  * clip_thresh = 0.01, cov2d blur +0.3, radius = ceil(3*sqrt(max eigenvalue)), BLOCK = 16
  * compositing thresholds: alpha = min(0.999, o*exp(-sigma)); skip sigma<0 or alpha<1/255; stop when T*(1-alpha) <= 1e-4
  * intersection order = STABLE sort by key (tile_id << 32 | float_bits(depth)), ties by Gaussian id
    (the reference's torch.sort is unstable => tie order there is unspecified, SURVEY §8a gotcha 7)

Every fp32 expression in `project_gaussians` is written as an explicit tree of single IEEE operations
(no fused multiply-add, fixed association) so the CUDA kernel, compiled with -fmad=false and the same
trees, reproduces it BIT-EXACTLY.  All functions are differentiable where gsplat's are (autograd of this
restatement is the backward oracle)."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

BLOCK = 16
CLIP_THRESH = 0.01

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396)
SH_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


def num_sh_bases(degree: int) -> int:
    return (degree + 1) ** 2


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float) -> torch.Tensor:
    """nerfstudio 1.0.0 `splatfacto.projection_matrix` (called at gc_model.py:115)."""
    t = znear * math.tan(0.5 * fovy)
    b = -t
    r = znear * math.tan(0.5 * fovx)
    l = -r
    n, f = znear, zfar
    return torch.tensor([
        [2 * n / (r - l), 0.0, (r + l) / (r - l), 0.0],
        [0.0, 2 * n / (t - b), (t + b) / (t - b), 0.0],
        [0.0, 0.0, (f + n) / (f - n), -1.0 * f * n / (f - n)],
        [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32)


def viewmat_from_c2w(c2w: torch.Tensor) -> torch.Tensor:
    """gc_model.py:97-107: flip y/z, analytic inverse."""
    R = c2w[:3, :3]
    T = c2w[:3, 3:4]
    R = R @ torch.diag(torch.tensor([1.0, -1.0, -1.0], dtype=R.dtype))
    R_inv = R.T
    T_inv = -R_inv @ T
    vm = torch.eye(4, dtype=R.dtype)
    vm[:3, :3] = R_inv
    vm[:3, 3:4] = T_inv
    return vm


def _sqrt(x: torch.Tensor) -> torch.Tensor:
    """Correctly rounded fp32 square root.  torch.sqrt on CPU goes through MKL VML and is NOT correctly rounded
    (0.6 % of inputs are 1 ulp off); sqrt in fp64 followed by rounding to fp32 is (53 >= 2*24+2 bits), which is what
    IEEE hardware sqrt (CUDA __fsqrt_rn, numpy) returns."""
    return torch.sqrt(x.double()).to(x.dtype)


def _quat_to_rotmat(q: torch.Tensor):
    """gsplat quat_to_rotmat: (w,x,y,z), normalised inside."""
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    inv = 1.0 / _sqrt(((w * w + x * x) + y * y) + z * z)
    w, x, y, z = w * inv, x * inv, y * inv, z * inv
    r00 = 1.0 - 2.0 * (y * y + z * z)
    r01 = 2.0 * (x * y - w * z)
    r02 = 2.0 * (x * z + w * y)
    r10 = 2.0 * (x * y + w * z)
    r11 = 1.0 - 2.0 * (x * x + z * z)
    r12 = 2.0 * (y * z - w * x)
    r20 = 2.0 * (x * z - w * y)
    r21 = 2.0 * (y * z + w * x)
    r22 = 1.0 - 2.0 * (x * x + y * y)
    return r00, r01, r02, r10, r11, r12, r20, r21, r22


def project_gaussians(means3d, scales, glob_scale, quats, viewmat, projmat, fx, fy, cx, cy, img_height, img_width,
                      tile_bounds, clip_thresh: float = CLIP_THRESH):
    """gsplat 0.1.3 project_gaussians forward -> (xys, depths, radii, conics, num_tiles_hit, cov3d).

    viewmat: [3,4] or [4,4] world->camera; projmat: [4,4] full projection (proj @ view)."""
    f32 = torch.float32
    means3d, scales, quats = means3d.to(f32), scales.to(f32), quats.to(f32)
    vm = viewmat.to(f32)
    pm = projmat.to(f32)
    px, py, pz = means3d[:, 0], means3d[:, 1], means3d[:, 2]

    def aff(m, r):  # ((m0*x + m1*y) + m2*z) + m3
        return ((m[r, 0] * px + m[r, 1] * py) + m[r, 2] * pz) + m[r, 3]

    tx, ty, tz = aff(vm, 0), aff(vm, 1), aff(vm, 2)
    valid = tz > clip_thresh

    # cov3d = (R S)(R S)^T, upper triangle
    r00, r01, r02, r10, r11, r12, r20, r21, r22 = _quat_to_rotmat(quats)
    gs = float(glob_scale)
    sx, sy, sz = gs * scales[:, 0], gs * scales[:, 1], gs * scales[:, 2]
    m00, m01, m02 = r00 * sx, r01 * sy, r02 * sz
    m10, m11, m12 = r10 * sx, r11 * sy, r12 * sz
    m20, m21, m22 = r20 * sx, r21 * sy, r22 * sz
    c00 = (m00 * m00 + m01 * m01) + m02 * m02
    c01 = (m00 * m10 + m01 * m11) + m02 * m12
    c02 = (m00 * m20 + m01 * m21) + m02 * m22
    c11 = (m10 * m10 + m11 * m11) + m12 * m12
    c12 = (m10 * m20 + m11 * m21) + m12 * m22
    c22 = (m20 * m20 + m21 * m21) + m22 * m22
    cov3d = torch.stack([c00, c01, c02, c11, c12, c22], dim=-1)

    # EWA projection
    fx32, fy32 = np.float32(fx), np.float32(fy)
    # gsplat's kernel receives fx, fy as fp32 and evaluates `0.5 * img_size / fx` in double (0.5 is a double literal)
    tan_fovx = np.float32(0.5 * img_width / float(fx32))
    tan_fovy = np.float32(0.5 * img_height / float(fy32))
    lim_x = float(np.float32(1.3) * tan_fovx)
    lim_y = float(np.float32(1.3) * tan_fovy)
    tz_safe = torch.where(valid, tz, torch.ones_like(tz))
    rz = 1.0 / tz_safe
    txc = tz_safe * torch.clamp(tx * rz, min=-lim_x, max=lim_x)
    tyc = tz_safe * torch.clamp(ty * rz, min=-lim_y, max=lim_y)
    rz2 = rz * rz
    j00 = float(fx32) * rz
    j02 = (-float(fx32) * txc) * rz2
    j11 = float(fy32) * rz
    j12 = (-float(fy32) * tyc) * rz2
    # T = J W  (W = rotation part of viewmat), rows 0 and 1 only
    t00 = j00 * vm[0, 0] + j02 * vm[2, 0]
    t01 = j00 * vm[0, 1] + j02 * vm[2, 1]
    t02 = j00 * vm[0, 2] + j02 * vm[2, 2]
    t10 = j11 * vm[1, 0] + j12 * vm[2, 0]
    t11 = j11 * vm[1, 1] + j12 * vm[2, 1]
    t12 = j11 * vm[1, 2] + j12 * vm[2, 2]
    # cov2d = T V T^T
    v0x = (t00 * c00 + t01 * c01) + t02 * c02
    v0y = (t00 * c01 + t01 * c11) + t02 * c12
    v0z = (t00 * c02 + t01 * c12) + t02 * c22
    v1x = (t10 * c00 + t11 * c01) + t12 * c02
    v1y = (t10 * c01 + t11 * c11) + t12 * c12
    v1z = (t10 * c02 + t11 * c12) + t12 * c22
    a = ((v0x * t00 + v0y * t01) + v0z * t02) + 0.3
    b = (v0x * t10 + v0y * t11) + v0z * t12
    c = ((v1x * t10 + v1y * t11) + v1z * t12) + 0.3

    det = a * c - b * b
    valid = valid & (det != 0.0)
    det_safe = torch.where(det != 0.0, det, torch.ones_like(det))
    inv_det = 1.0 / det_safe
    conic = torch.stack([c * inv_det, (-b) * inv_det, a * inv_det], dim=-1)
    b_mid = 0.5 * (a + c)
    disc = _sqrt(torch.clamp(b_mid * b_mid - det, min=0.1))
    v1 = b_mid + disc
    v2 = b_mid - disc
    radius = torch.ceil(3.0 * _sqrt(torch.maximum(v1, v2)))

    def hom(r):
        return ((pm[r, 0] * px + pm[r, 1] * py) + pm[r, 2] * pz) + pm[r, 3]

    hx, hy, hw = hom(0), hom(1), hom(3)
    rw = 1.0 / (hw + 1e-6)
    ndc_x, ndc_y = hx * rw, hy * rw
    cxf, cyf = float(np.float32(cx)), float(np.float32(cy))
    xs = ((0.5 * float(img_width)) * ndc_x + cxf) - 0.5
    ys = ((0.5 * float(img_height)) * ndc_y + cyf) - 0.5

    tbx, tby = int(tile_bounds[0]), int(tile_bounds[1])
    with torch.no_grad():
        tcx, tcy, tr = xs / BLOCK, ys / BLOCK, radius / BLOCK
        # float->int conversion saturates like CUDA's cvt.rzi.s32.f32 (torch's .to(int32) is UB outside range)
        def f2i(v):
            return torch.clamp(torch.nan_to_num(v, nan=0.0), -2.0e9, 2.0e9).to(torch.int64)
        tmin_x = torch.clamp(f2i(tcx - tr), 0, tbx)
        tmax_x = torch.clamp(f2i((tcx + tr) + 1.0), 0, tbx)
        tmin_y = torch.clamp(f2i(tcy - tr), 0, tby)
        tmax_y = torch.clamp(f2i((tcy + tr) + 1.0), 0, tby)
        area = (tmax_x - tmin_x) * (tmax_y - tmin_y)
        valid = valid & (area > 0)
        num_tiles_hit = torch.where(valid, area, torch.zeros_like(area)).to(torch.int32)
        radii = torch.where(valid, f2i(radius), torch.zeros_like(area)).to(torch.int32)

    zero = torch.zeros_like(tz)
    xys = torch.stack([torch.where(valid, xs, zero), torch.where(valid, ys, zero)], dim=-1)
    depths = torch.where(valid, tz, zero)
    conics = torch.where(valid[:, None], conic, torch.zeros_like(conic))
    return xys, depths, radii, conics, num_tiles_hit, cov3d


def spherical_harmonics(degree: int, viewdirs: torch.Tensor, coeffs: torch.Tensor) -> torch.Tensor:
    """gsplat 0.1.3 sh_coeffs_to_color. viewdirs [N,3] unit; coeffs [N,K,3]; returns [N,3] (no +0.5, no clamp)."""
    x, y, z = viewdirs[:, 0:1], viewdirs[:, 1:2], viewdirs[:, 2:3]
    c = coeffs
    col = SH_C0 * c[:, 0]
    if degree < 1:
        return col
    col = col + SH_C1 * (((-y) * c[:, 1] + z * c[:, 2]) - x * c[:, 3])
    if degree < 2:
        return col
    xx, xy, xz, yy, yz, zz = x * x, x * y, x * z, y * y, y * z, z * z
    col = col + ((((SH_C2[0] * xy) * c[:, 4] + (SH_C2[1] * yz) * c[:, 5])
                  + (SH_C2[2] * ((2.0 * zz - xx) - yy)) * c[:, 6])
                 + (SH_C2[3] * xz) * c[:, 7]) + (SH_C2[4] * (xx - yy)) * c[:, 8]
    if degree < 3:
        return col
    col = col + ((((((SH_C3[0] * y * (3.0 * xx - yy)) * c[:, 9] + (SH_C3[1] * xy * z) * c[:, 10])
                    + (SH_C3[2] * y * ((4.0 * zz - xx) - yy)) * c[:, 11])
                   + (SH_C3[3] * z * ((2.0 * zz - 3.0 * xx) - 3.0 * yy)) * c[:, 12])
                  + (SH_C3[4] * x * ((4.0 * zz - xx) - yy)) * c[:, 13])
                 + (SH_C3[5] * z * (xx - yy)) * c[:, 14]) + (SH_C3[6] * x * (xx - 3.0 * yy)) * c[:, 15]
    return col


def bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds):
    """cumsum -> map_gaussian_to_intersects -> STABLE sort -> tile bin edges.  All integer work (numpy).
    Returns (isect_keys_sorted int64 [M], gaussian_ids_sorted int32 [M], tile_bins int32 [T,2])."""
    xys_n = xys.detach().cpu().numpy().astype(np.float32)
    dep_n = depths.detach().cpu().numpy().astype(np.float32)
    rad_n = radii.detach().cpu().numpy().astype(np.int32)
    nth = num_tiles_hit.detach().cpu().numpy().astype(np.int64)
    tbx, tby = int(tile_bounds[0]), int(tile_bounds[1])
    cum = np.cumsum(nth)
    M = int(cum[-1]) if len(cum) else 0
    keys = np.zeros(M, dtype=np.int64)
    gids = np.zeros(M, dtype=np.int32)
    depth_bits = dep_n.view(np.int32).astype(np.int64) & 0xFFFFFFFF
    blk = np.float32(BLOCK)
    for g in np.nonzero(rad_n > 0)[0]:
        tcx, tcy, tr = xys_n[g, 0] / blk, xys_n[g, 1] / blk, np.float32(rad_n[g]) / blk
        x0 = min(tbx, max(0, int(tcx - tr)))
        x1 = min(tbx, max(0, int((tcx + tr) + np.float32(1.0))))
        y0 = min(tby, max(0, int(tcy - tr)))
        y1 = min(tby, max(0, int((tcy + tr) + np.float32(1.0))))
        cur = int(cum[g - 1]) if g > 0 else 0
        for i in range(y0, y1):
            for j in range(x0, x1):
                keys[cur] = ((i * tbx + j) << 32) | int(depth_bits[g])
                gids[cur] = g
                cur += 1
    order = np.argsort(keys, kind="stable")
    keys_s, gids_s = keys[order], gids[order]
    ntiles = tbx * tby
    bins = np.zeros((ntiles, 2), dtype=np.int32)
    if M:
        tile_of = (keys_s >> 32).astype(np.int64)
        starts = np.searchsorted(tile_of, np.arange(ntiles), side="left")
        ends = np.searchsorted(tile_of, np.arange(ntiles), side="right")
        bins[:, 0], bins[:, 1] = starts, ends
    return keys_s, gids_s, bins


def bin_and_sort_vectorized(xys, depths, radii, num_tiles_hit, tile_bounds):
    """Same result as `bin_and_sort` without the per-Gaussian Python loop (numpy repeat/arange instead), for the
    full-size parity tests (1 M Gaussians, ~4 M intersections).  tests/test_raster_cpu.py checks it equals the loop."""
    xys_n = xys.detach().cpu().numpy().astype(np.float32)
    dep_n = depths.detach().cpu().numpy().astype(np.float32)
    rad_n = radii.detach().cpu().numpy().astype(np.int32)
    nth = num_tiles_hit.detach().cpu().numpy().astype(np.int64)
    tbx, tby = int(tile_bounds[0]), int(tile_bounds[1])
    blk = np.float32(BLOCK)
    tcx, tcy, tr = xys_n[:, 0] / blk, xys_n[:, 1] / blk, rad_n.astype(np.float32) / blk
    trunc = lambda a: np.trunc(a).astype(np.int64)  # noqa: E731  (Python int() of a float truncates toward zero)
    x0 = np.clip(trunc(tcx - tr), 0, tbx)
    x1 = np.clip(trunc((tcx + tr) + np.float32(1.0)), 0, tbx)
    y0 = np.clip(trunc(tcy - tr), 0, tby)
    y1 = np.clip(trunc((tcy + tr) + np.float32(1.0)), 0, tby)
    wdt = x1 - x0
    cnt = np.where(rad_n > 0, wdt * (y1 - y0), 0)
    assert np.array_equal(cnt, nth), "num_tiles_hit does not match the tile bounding boxes"
    M = int(cnt.sum())
    g_of = np.repeat(np.arange(len(cnt), dtype=np.int64), cnt)
    local = np.arange(M, dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    wg = np.maximum(wdt[g_of], 1)
    tile = (y0[g_of] + local // wg) * tbx + (x0[g_of] + local % wg)
    depth_bits = dep_n.view(np.int32).astype(np.int64) & 0xFFFFFFFF
    keys = (tile << 32) | depth_bits[g_of]
    order = np.argsort(keys, kind="stable")
    keys_s, gids_s = keys[order], g_of[order].astype(np.int32)
    ntiles = tbx * tby
    bins = np.zeros((ntiles, 2), dtype=np.int32)
    if M:
        tile_of = keys_s >> 32
        bins[:, 0] = np.searchsorted(tile_of, np.arange(ntiles), side="left")
        bins[:, 1] = np.searchsorted(tile_of, np.arange(ntiles), side="right")
    return keys_s, gids_s, bins


def rasterize_sorted(xys, conics, colors, opacities, gids_sorted, tile_bins, img_height, img_width,
                     background: Optional[torch.Tensor]):
    """rasterize_forward for C channels, differentiable (autograd = backward oracle).

    colors [N,C]; opacities [N] or [N,1]; returns (img [H,W,C], alpha [H,W], final_idx int32 [H,W])."""
    H, W = int(img_height), int(img_width)
    C = colors.shape[1]
    tbx = (W + BLOCK - 1) // BLOCK
    tby = (H + BLOCK - 1) // BLOCK
    opac = opacities.reshape(-1)
    dev = colors.device  # CPU in the small-size tests; the full-size GPU parity tests run this same code in fp32 on CUDA
    if background is None:
        background = torch.ones(C, dtype=colors.dtype, device=dev)
    background = background.to(dev)
    img_rows = [[None] * tbx for _ in range(tby)]
    alpha_rows = [[None] * tbx for _ in range(tby)]
    fidx = np.zeros((H, W), dtype=np.int32)
    gids_t = torch.as_tensor(np.asarray(gids_sorted), dtype=torch.long).to(dev)
    tile_bins = np.asarray(tile_bins.cpu() if isinstance(tile_bins, torch.Tensor) else tile_bins)
    for ti in range(tby):
        for tj in range(tbx):
            h0, w0 = ti * BLOCK, tj * BLOCK
            hh, ww = min(BLOCK, H - h0), min(BLOCK, W - w0)
            start, end = int(tile_bins[ti * tbx + tj][0]), int(tile_bins[ti * tbx + tj][1])
            py = torch.arange(h0, h0 + hh, dtype=torch.float32, device=dev)[:, None].expand(hh, ww).reshape(-1)
            px = torch.arange(w0, w0 + ww, dtype=torch.float32, device=dev)[None, :].expand(hh, ww).reshape(-1)
            P = hh * ww
            if end <= start:
                img_rows[ti][tj] = background[None, :].expand(P, C).reshape(hh, ww, C)
                alpha_rows[ti][tj] = torch.zeros(hh, ww, device=dev)
                continue
            ids = gids_t[start:end]
            xy, con, col, op = xys[ids], conics[ids], colors[ids], opac[ids]
            dx = xy[None, :, 0] - px[:, None]
            dy = xy[None, :, 1] - py[:, None]
            sigma = 0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) + con[None, :, 1] * dx * dy
            alpha = torch.clamp(op[None, :] * torch.exp(-sigma), max=0.999)
            keep = (sigma >= 0.0) & (alpha >= 1.0 / 255.0)
            a_eff = torch.where(keep, alpha, torch.zeros_like(alpha))
            one_m = 1.0 - a_eff
            T_incl = torch.cumprod(one_m, dim=1)                       # T after each Gaussian
            T_excl = torch.cat([torch.ones(P, 1, device=dev), T_incl[:, :-1]], dim=1)
            with torch.no_grad():
                stop = keep & (T_incl <= 1e-4)                          # first Gaussian that would end the pixel
                stopped = torch.cumsum(stop.to(torch.int32), dim=1) > 0  # at and after that Gaussian: no contribution
                active = keep & ~stopped
                any_stop = stopped[:, -1]
                first_stop = torch.argmax(stopped.to(torch.int32), dim=1)
                pos = torch.arange(end - start, device=dev)[None, :].expand(P, -1)
                last = torch.where(active, pos, torch.full_like(pos, -1)).max(dim=1).values
            vis = torch.where(active, a_eff * T_excl, torch.zeros_like(a_eff))
            out = vis @ col
            # final T: the T before the stopping Gaussian, or after the last one
            T_final = torch.where(any_stop, T_excl.gather(1, first_stop[:, None])[:, 0], T_incl[:, -1])
            out = out + T_final[:, None] * background[None, :]
            img_rows[ti][tj] = out.reshape(hh, ww, C)
            alpha_rows[ti][tj] = (1.0 - T_final).reshape(hh, ww)
            fidx[h0:h0 + hh, w0:w0 + ww] = (torch.where(last >= 0, last + start, torch.zeros_like(last))
                                            .reshape(hh, ww).cpu().numpy())
    img = torch.cat([torch.cat(r, dim=1) for r in img_rows], dim=0)
    alpha = torch.cat([torch.cat(r, dim=1) for r in alpha_rows], dim=0)
    return img, alpha, fidx


def rasterize_gaussians(xys, depths, radii, conics, num_tiles_hit, colors, opacity, img_height, img_width,
                        background=None, return_alpha=False):
    """gsplat 0.1.3 `rasterize_gaussians` signature (call sites gc_model.py:174-186, 191-202)."""
    tile_bounds = ((img_width + BLOCK - 1) // BLOCK, (img_height + BLOCK - 1) // BLOCK, 1)
    if background is None:
        background = torch.ones(colors.shape[-1], dtype=torch.float32)
    if int(num_tiles_hit.sum()) < 1:
        img = torch.ones(img_height, img_width, colors.shape[-1]) * background
        return (img, torch.zeros(img_height, img_width)) if return_alpha else img
    _, gids, bins = bin_and_sort(xys, depths, radii, num_tiles_hit, tile_bounds)
    img, alpha, _ = rasterize_sorted(xys, conics, colors, opacity, gids, bins, img_height, img_width, background)
    return (img, alpha) if return_alpha else img


def get_outputs(params: dict, c2w: torch.Tensor, fx, fy, cx, cy, H, W, sh_degree_active: int,
                background: torch.Tensor, training: bool = False):
    """Restates GaussCtrlModel.get_outputs (gc_model.py:57-206), eval branch by default.

    params: means [N,3], scales(log) [N,3], quats [N,4], features_dc [N,3], features_rest [N,15,3],
            opacities(logit) [N,1]."""
    vm = viewmat_from_c2w(c2w)
    fovx = 2 * math.atan(W / (2 * fx))
    fovy = 2 * math.atan(H / (2 * fy))
    pm = projection_matrix(0.001, 1000, fovx, fovy)
    tile_bounds = ((W + BLOCK - 1) // BLOCK, (H + BLOCK - 1) // BLOCK, 1)
    means, quats = params["means"], params["quats"]
    colors = torch.cat((params["features_dc"][:, None, :], params["features_rest"]), dim=1)
    xys, depths, radii, conics, nth, _ = project_gaussians(
        means, torch.exp(params["scales"]), 1, quats / quats.norm(dim=-1, keepdim=True), vm[:3, :], pm @ vm,
        fx, fy, cx, cy, H, W, tile_bounds)
    if int(radii.sum()) == 0:
        return {"rgb": background.repeat(H, W, 1)}
    viewdirs = means.detach() - c2w[:3, 3]
    viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
    rgbs = torch.clamp(spherical_harmonics(sh_degree_active, viewdirs, colors) + 0.5, min=0.0)
    opac = torch.sigmoid(params["opacities"])
    rgb, alpha = rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, opac, H, W, background=background,
                                     return_alpha=True)
    alpha = alpha[..., None]
    rgb = torch.clamp(rgb, max=1.0)
    depth_im = None
    if not training:
        depth_im = rasterize_gaussians(xys, depths, radii, conics, nth, depths[:, None].repeat(1, 3), opac, H, W,
                                       background=torch.zeros(3))[..., 0:1]
        depth_im = torch.where(alpha > 0, depth_im / torch.where(alpha > 0, alpha, torch.ones_like(alpha)),
                               torch.full_like(depth_im, 1000.0))
    return {"rgb": rgb, "depth": depth_im, "accumulation": alpha, "xys": xys, "radii": radii, "depths": depths,
            "conics": conics, "num_tiles_hit": nth}
