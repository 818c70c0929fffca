"""GaussCtrlModel on the sm_100a rasteriser: mirror of gaussctrl/gc_model.py:39-221.

`get_outputs(camera)` keeps the reference's contract - returns {"rgb" [H,W,3], "depth" [H,W,1], "accumulation"
[H,W,1]} fp32 and sets `self.xys`, `self.radii`, `self.last_size` - but in eval mode runs ONE binning and ONE fused
4-channel (r,g,b,depth) composite instead of the reference's two bin+sort+rasterize passes (gc_model.py:174-202)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Type, Union

import torch

from . import gsplat_ops
from ._compat import Cameras, SplatfactoModel, SplatfactoModelConfig, renderers


def projection_matrix(znear: float, zfar: float, fovx: float, fovy: float, device="cpu") -> torch.Tensor:
    """nerfstudio 1.0.0 splatfacto.projection_matrix (called at gc_model.py:115)."""
    t = znear * math.tan(0.5 * fovy)
    b = -t
    r = znear * math.tan(0.5 * fovx)
    l = -r
    n, f = znear, zfar
    return torch.tensor([[2 * n / (r - l), 0.0, (r + l) / (r - l), 0.0],
                         [0.0, 2 * n / (t - b), (t + b) / (t - b), 0.0],
                         [0.0, 0.0, (f + n) / (f - n), -1.0 * f * n / (f - n)],
                         [0.0, 0.0, 1.0, 0.0]], dtype=torch.float32, device=device)


def viewmat_from_c2w(c2w: torch.Tensor) -> torch.Tensor:
    """gc_model.py:97-107: flip y/z, analytic rigid inverse.  Tiny host-side 4x4 math."""
    c2w = c2w.detach().to("cpu", torch.float32)
    R = c2w[:3, :3] @ torch.diag(torch.tensor([1.0, -1.0, -1.0]))
    T = c2w[:3, 3:4]
    vm = torch.eye(4)
    vm[:3, :3] = R.T
    vm[:3, 3:4] = -R.T @ T
    return vm


def render_gaussians(params: Dict[str, torch.Tensor], c2w: torch.Tensor, fx, fy, cx, cy, H: int, W: int,
                     sh_degree_active: int, background: torch.Tensor, training: bool = False,
                     state: Optional[dict] = None) -> Dict[str, torch.Tensor]:
    """Functional core of get_outputs (gc_model.py:95-206) for one camera.  params are CUDA tensors under the
    splatfacto names; c2w is the [3,4]/[4,4] camera-to-world matrix."""
    means = params["means"]
    dev = means.device
    vm = viewmat_from_c2w(c2w)
    fovx = 2 * math.atan(W / (2 * fx))
    fovy = 2 * math.atan(H / (2 * fy))
    pm = projection_matrix(0.001, 1000, fovx, fovy)
    tile_bounds = ((W + 15) // 16, (H + 15) // 16, 1)
    if (not training and not torch.is_grad_enabled() and sh_degree_active >= 0 and params["features_rest"].shape[1] == 15
            and gsplat_ops.FUSED_EVAL):
        # eval path (render_reverse, ns-gaussctrl-render): fused front end on the raw parameters
        rgb, depth, alpha, xys, radii = gsplat_ops.render_eval_fused(
            params, vm, pm @ vm, c2w.detach().to("cpu", torch.float32)[:3, 3].tolist(), fx, fy, cx, cy, H, W,
            sh_degree_active, background, defer_check=bool(state is not None and state.get("defer_check")))
        if state is not None:
            state["xys"], state["radii"] = xys, radii
        return {"rgb": rgb, "depth": depth, "accumulation": alpha}
    quats = params["quats"]
    colors = torch.cat((params["features_dc"][:, None, :], params["features_rest"]), dim=1)
    xys, depths, radii, conics, nth, _ = gsplat_ops.project_gaussians(
        means, torch.exp(params["scales"]), 1, quats / quats.norm(dim=-1, keepdim=True), vm[:3, :], pm @ vm, fx, fy, cx,
        cy, H, W, tile_bounds)
    if state is not None:
        state["xys"], state["radii"] = xys, radii
    if int(radii.sum().item()) == 0:
        return {"rgb": background.repeat(H, W, 1)}
    if sh_degree_active >= 0 and colors.shape[1] > 1:
        viewdirs = means.detach() - c2w.detach().to(dev, torch.float32)[:3, 3]
        viewdirs = viewdirs / viewdirs.norm(dim=-1, keepdim=True)
        rgbs = torch.clamp(gsplat_ops.spherical_harmonics(sh_degree_active, viewdirs, colors) + 0.5, min=0.0)
    else:
        rgbs = torch.sigmoid(colors[:, 0, :])
    opac = torch.sigmoid(params["opacities"])
    if training:
        rgb, alpha = gsplat_ops.rasterize_gaussians(xys, depths, radii, conics, nth, rgbs, opac, H, W,
                                                    background=background, return_alpha=True)
        return {"rgb": torch.clamp(rgb, max=1.0), "depth": None, "accumulation": alpha[..., None]}
    rgb, depth, alpha = gsplat_ops.rasterize_rgbd(xys, depths, radii, conics, nth, rgbs, opac, H, W, background)
    return {"rgb": rgb, "depth": depth, "accumulation": alpha}


@dataclass
class GaussCtrlModelConfig(SplatfactoModelConfig):
    """Same fields/defaults as gaussctrl/gc_model.py:39-50."""
    _target: Type = field(default_factory=lambda: GaussCtrlModel)
    use_lpips: bool = True
    use_l1: bool = True
    patch_size: int = 32
    lpips_loss_mult: float = 1.0


class GaussCtrlModel(SplatfactoModel):
    config: GaussCtrlModelConfig

    def get_outputs(self, camera: Cameras) -> Dict[str, Union[torch.Tensor, List]]:
        if not isinstance(camera, Cameras):
            print("Called get_outputs with not a camera")
            return {}
        assert camera.shape[0] == 1, "Only one camera at a time"
        if self.training:
            bc = getattr(self.config, "background_color", "random")
            if bc == "random":
                background = torch.rand(3, device=self.device)
            elif bc == "white":
                background = torch.ones(3, device=self.device)
            elif bc == "black":
                background = torch.zeros(3, device=self.device)
            else:
                background = self.background_color.to(self.device)
        elif renderers.BACKGROUND_COLOR_OVERRIDE is not None:      # gc_model.py:84-85 (ns-render's background override)
            background = renderers.BACKGROUND_COLOR_OVERRIDE.to(self.device)
        else:
            background = self.background_color.to(self.device)
        params = {k: getattr(self, k) for k in ("means", "scales", "quats", "features_dc", "features_rest", "opacities")}
        if self.crop_box is not None and not self.training:
            crop_ids = self.crop_box.within(self.means).squeeze()
            if crop_ids.sum() == 0:
                return {"rgb": background.repeat(int(camera.height.item()), int(camera.width.item()), 1)}
            params = {k: v[crop_ids] for k, v in params.items()}
        # splatfacto's resolution schedule (gc_model.py:96-97, :170): intrinsics and size are read at the downscaled
        # resolution and the camera is restored afterwards
        camera_downscale = self._get_downscale_factor()
        camera.rescale_output_resolution(1 / camera_downscale)
        try:
            W, H = int(camera.width.item()), int(camera.height.item())
            self.last_size = (H, W)
            intr = (camera.fx.item(), camera.fy.item(), camera.cx.item(), camera.cy.item())
        finally:
            camera.rescale_output_resolution(camera_downscale)
        sh_degree = getattr(self.config, "sh_degree", 3)
        n = min(self.step // getattr(self.config, "sh_degree_interval", 1000), sh_degree) if sh_degree > 0 else -1
        # sync-free eval renders: the caller promises to call gsplat_ops.check_deferred_overflow() (render_reverse does)
        state: dict = {"defer_check": bool(getattr(self, "defer_isect_check", False))}
        out = render_gaussians(params, camera.camera_to_worlds[0], *intr, H, W, n, background, training=self.training,
                               state=state)
        self.xys, self.radii = state.get("xys"), state.get("radii")
        # gc_model.py:159-160: splatfacto's after_train densification callback reads `self.xys.grad`
        if self.training and self.xys is not None and self.xys.requires_grad:
            self.xys.retain_grad()
        return out

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        """nerfstudio 1.0.0 SplatfactoModel.get_loss_dict (what gc_pipeline.py:283-285 calls; the reference's
        GaussCtrlModel inherits it): main_loss = (1-l)*L1 + l*(1-SSIM) with the loss AND its gradient w.r.t. the
        render from one fused C-ABI call (finetune.l1_ssim_loss); `batch["mask"]` blacks out both images first;
        scale regularisation (off by default: use_scale_regularization False) every 10th step as splatfacto does."""
        from .finetune import l1_ssim_loss
        gt = self.get_gt_img(batch["image"]) if hasattr(self, "get_gt_img") else batch["image"].to(self.device)
        pred = outputs["rgb"]
        if "mask" in batch:
            # splatfacto: masked-out pixels are black in both images
            mask = batch["mask"]
            if hasattr(self, "_downscale_if_required"):
                mask = self._downscale_if_required(mask)
            mask = mask.to(self.device).to(pred.dtype)
            assert mask.shape[:2] == gt.shape[:2] == pred.shape[:2]
            gt, pred = gt * mask, pred * mask
        main_loss, parts = l1_ssim_loss(pred, gt, float(getattr(self.config, "ssim_lambda", 0.2)))
        self.last_loss_parts = parts  # device float[3]: (main_loss, L1, ssim) - no host sync here
        if getattr(self.config, "use_scale_regularization", False) and self.step % 10 == 0:
            ratio = float(getattr(self.config, "max_gauss_ratio", 10.0))
            scale_exp = torch.exp(self.scales)
            scale_reg = torch.clamp(scale_exp.amax(dim=-1) / scale_exp.amin(dim=-1), min=ratio) - ratio
            scale_reg = 0.1 * scale_reg.mean()
        else:
            scale_reg = torch.zeros((), device=main_loss.device)
        return {"main_loss": main_loss, "scale_reg": scale_reg}

    @torch.no_grad()
    def get_outputs_for_cameras(self, cameras: List[Cameras], obb_box=None, streams=None) -> List[Dict[str, torch.Tensor]]:
        """B200 extension: `get_outputs_for_camera` for a LIST of cameras (render_reverse's loop over the training views,
        gc_pipeline.py:126-133) as one batched call into the rasteriser (gsplat_ops.render_eval_batch): same kernels and
        results as the per-camera path, without ~0.3 ms of interpreter work per view.  Cameras stay on the host.  Falls
        back to the per-camera path when a crop box is set, the views differ in size, or the SH layout is not degree 3.
        The intersection-capacity check is deferred: call gsplat_ops.check_deferred_overflow() afterwards."""
        cams = [c if len(c.shape) else c.reshape((1,)) for c in cameras]
        sizes = {(int(c.width.item()), int(c.height.item())) for c in cams}
        sh_degree = getattr(self.config, "sh_degree", 3)
        n = min(self.step // getattr(self.config, "sh_degree_interval", 1000), sh_degree) if sh_degree > 0 else -1
        if (obb_box is not None or self.crop_box is not None or len(sizes) != 1 or n < 0 or not cams
                or self.features_rest.shape[1] != 15 or not self.means.is_cuda or not gsplat_ops.FUSED_EVAL):
            self.defer_isect_check = True
            try:
                return [self.get_outputs_for_camera(c, obb_box) for c in cams]
            finally:
                self.defer_isect_check = False
        (W, H), = sizes
        if renderers.BACKGROUND_COLOR_OVERRIDE is not None:
            background = renderers.BACKGROUND_COLOR_OVERRIDE.to(self.device)
        else:
            background = self.background_color.to(self.device)

        def camera(i):
            c = cams[i]
            c2w = c.camera_to_worlds[0].detach().to("cpu", torch.float32)
            fx, fy, cx, cy = c.fx.item(), c.fy.item(), c.cx.item(), c.cy.item()
            vm = viewmat_from_c2w(c2w)
            pm = projection_matrix(0.001, 1000, 2 * math.atan(W / (2 * fx)), 2 * math.atan(H / (2 * fy)))
            return vm, pm @ vm, c2w[:3, 3], (fx, fy, cx, cy)

        camera.n_views = len(cams)
        params = {k: getattr(self, k) for k in ("means", "scales", "quats", "features_dc", "features_rest", "opacities")}
        rgb, depth, alpha = gsplat_ops.render_eval_batch(params, camera, H, W, n, background, streams=streams)
        self.last_size = (H, W)
        return [{"rgb": rgb[i], "depth": depth[i], "accumulation": alpha[i]} for i in range(len(cams))]

    @torch.no_grad()
    def get_outputs_for_camera(self, camera: Cameras, obb_box=None) -> Dict[str, torch.Tensor]:
        assert camera is not None, "must provide camera to gaussian model"
        self.set_crop(obb_box)
        self.training = False
        # the reference moves the camera to the device first (gc_model.py:219); the kernels here take the pose and the
        # intrinsics as launch arguments, so a host-resident camera is read without a device synchronisation
        outs = self.get_outputs(camera)
        self.training = True
        return outs  # type: ignore
