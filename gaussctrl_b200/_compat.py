"""nerfstudio compatibility shims.  With nerfstudio 1.0.0 installed the real base classes are used and the plugin
registers exactly like the reference (pyproject entry point `nerfstudio.method_configs`).  Without it (this image
has no nerfstudio) minimal stand-ins with the same attribute names keep the hot-path classes importable and
testable; they carry no behaviour beyond what the hot path reads."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional, Type

import torch

try:  # pragma: no cover - exercised only where nerfstudio exists
    from nerfstudio.cameras.cameras import Cameras  # type: ignore
    from nerfstudio.models.splatfacto import SplatfactoModel, SplatfactoModelConfig  # type: ignore
    from nerfstudio.pipelines.base_pipeline import VanillaPipeline, VanillaPipelineConfig  # type: ignore
    HAVE_NERFSTUDIO = True
except Exception:  # ModuleNotFoundError here
    HAVE_NERFSTUDIO = False

    class Cameras:  # the attributes GaussCtrlModel.get_outputs reads (gc_model.py:65-113)
        def __init__(self, camera_to_worlds, fx, fy, cx, cy, width, height):
            self.camera_to_worlds = torch.as_tensor(camera_to_worlds, dtype=torch.float32).reshape(-1, 3, 4)
            n = self.camera_to_worlds.shape[0]

            def t(v, dtype=torch.float32):  # scalars broadcast over the n cameras, like nerfstudio's Cameras
                v = torch.as_tensor(v, dtype=dtype).reshape(-1, 1)
                return v.expand(n, 1).contiguous() if v.shape[0] == 1 and n > 1 else v

            self.fx, self.fy, self.cx, self.cy = t(fx), t(fy), t(cx), t(cy)
            self.width, self.height = t(width, torch.int64), t(height, torch.int64)

        @property
        def shape(self):
            return (self.camera_to_worlds.shape[0],)

        def __len__(self):
            return self.camera_to_worlds.shape[0]

        def __getitem__(self, i):
            if isinstance(i, int):
                i = slice(i, i + 1)
            return Cameras(self.camera_to_worlds[i], self.fx[i], self.fy[i], self.cx[i], self.cy[i], self.width[i],
                           self.height[i])

        def to(self, device):
            c = Cameras(self.camera_to_worlds.to(device), self.fx.to(device), self.fy.to(device), self.cx.to(device),
                        self.cy.to(device), self.width.to(device), self.height.to(device))
            return c

        def rescale_output_resolution(self, s):
            return None

    @dataclass
    class SplatfactoModelConfig:
        _target: Type = field(default_factory=lambda: SplatfactoModel)
        sh_degree: int = 3
        sh_degree_interval: int = 1000
        background_color: str = "random"
        ssim_lambda: float = 0.2

    class SplatfactoModel(torch.nn.Module):
        """Owns the Gaussian parameters under the nerfstudio names (gc_model.py:124-136 reads them)."""

        def __init__(self, config: Optional[Any] = None, num_points: int = 0, seed_points=None, device="cpu", **_):
            super().__init__()
            self.config = config if config is not None else SplatfactoModelConfig()
            P = torch.nn.Parameter
            n = num_points
            self.means = P(torch.zeros(n, 3))
            self.scales = P(torch.zeros(n, 3))
            self.quats = P(torch.zeros(n, 4))
            self.features_dc = P(torch.zeros(n, 3))
            self.features_rest = P(torch.zeros(n, 15, 3))
            self.opacities = P(torch.zeros(n, 1))
            self.background_color = torch.zeros(3)
            self.crop_box = None
            self.step = 30000
            self.xys = None
            self.radii = None
            self.last_size = None

        @property
        def device(self):
            return self.means.device

        def _get_downscale_factor(self):
            return 1

        def set_crop(self, box):
            self.crop_box = box

        def forward(self, camera):
            return self.get_outputs(camera)

        def get_gt_img(self, image: torch.Tensor) -> torch.Tensor:
            return image.to(self.device)  # downscale factor 1 after step 30000 (resolution schedule finished)

        @torch.no_grad()
        def get_metrics_dict(self, outputs, batch):
            gt = self.get_gt_img(batch["image"])
            mse = torch.mean((outputs["rgb"].detach() - gt) ** 2)
            return {"psnr": 10.0 * torch.log10(1.0 / mse), "gaussian_count": self.means.shape[0]}

    @dataclass
    class VanillaPipelineConfig:
        _target: Type = field(default_factory=lambda: VanillaPipeline)
        datamanager: Any = None
        model: Any = None

    class VanillaPipeline(torch.nn.Module):
        def __init__(self, config=None, device="cpu", test_mode="val", world_size=1, local_rank=0, grad_scaler=None):
            super().__init__()
            self.config = config
            self.test_mode = test_mode
            self.world_size = world_size
            self.datamanager = None
            self._model = None

        @property
        def model(self):
            return self._model
