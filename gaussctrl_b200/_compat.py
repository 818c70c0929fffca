"""nerfstudio compatibility shims.  With nerfstudio 1.0.0 installed the real base classes are used and the plugin
registers exactly like the reference (pyproject entry point `nerfstudio.method_configs`).  Without it (this image
has no nerfstudio) minimal stand-ins with the same attribute names keep the hot-path classes importable and
testable; they carry no behaviour beyond what the hot path reads."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional, Type

import torch

try:  # exercised by tests/test_plugin_seam_cpu.py against a fake `nerfstudio` tree that enforces the same contracts
    from nerfstudio.cameras.cameras import Cameras  # type: ignore
    from nerfstudio.data.datamanagers.full_images_datamanager import (  # type: ignore
        FullImageDatamanager, FullImageDatamanagerConfig)
    from nerfstudio.model_components import renderers  # type: ignore
    from nerfstudio.models.splatfacto import SplatfactoModel, SplatfactoModelConfig  # type: ignore
    from nerfstudio.pipelines.base_pipeline import VanillaPipeline, VanillaPipelineConfig  # type: ignore
    HAVE_NERFSTUDIO = True
except ModuleNotFoundError:
    HAVE_NERFSTUDIO = False
    import types as _types

    renderers = _types.SimpleNamespace(BACKGROUND_COLOR_OVERRIDE=None)  # nerfstudio.model_components.renderers

    class _InstantiateConfig:
        """nerfstudio.configs.base_config.InstantiateConfig: `setup(**kwargs)` builds `_target(self, **kwargs)`."""

        def setup(self, **kwargs):
            return self._target(self, **kwargs)

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
    class SplatfactoModelConfig(_InstantiateConfig):
        _target: Type = field(default_factory=lambda: SplatfactoModel)
        sh_degree: int = 3
        sh_degree_interval: int = 1000
        background_color: str = "random"
        ssim_lambda: float = 0.2
        use_scale_regularization: bool = False
        max_gauss_ratio: float = 10.0
        stop_split_at: int = 15000

    class SplatfactoModel(torch.nn.Module):
        """Owns the Gaussian parameters under the nerfstudio names (gc_model.py:124-136 reads them)."""

        def __init__(self, config: Optional[Any] = None, num_points: int = 0, seed_points=None, device="cpu", **_):
            super().__init__()
            self.config = config if config is not None else SplatfactoModelConfig()
            P = torch.nn.Parameter
            n = num_points if seed_points is None else int(seed_points[0].shape[0])
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
    class FullImageDatamanagerConfig(_InstantiateConfig):
        """The fields of nerfstudio 1.0.0's FullImageDatamanagerConfig that the hot path or its callers read."""
        _target: Type = field(default_factory=lambda: FullImageDatamanager)
        dataparser: Any = None
        camera_res_scale_factor: float = 1.0
        eval_num_images_to_sample_from: int = -1
        eval_num_times_to_repeat_images: int = -1
        eval_image_indices: Optional[tuple] = (0,)
        cache_images: str = "cpu"
        cache_images_type: str = "float32"

    class FullImageDatamanager(torch.nn.Module):
        """Stand-in: holds `train_dataset`/`cached_train` when a dataparser is configured; nothing else."""

        def __init__(self, config=None, device="cpu", test_mode="val", world_size=1, local_rank=0, **_):
            super().__init__()
            self.config, self.device, self.test_mode = config, device, test_mode
            self.world_size, self.local_rank = world_size, local_rank
            self.train_dataset = None
            self.eval_dataset = None
            self.cached_train: list = []

    @dataclass
    class VanillaPipelineConfig(_InstantiateConfig):
        _target: Type = field(default_factory=lambda: VanillaPipeline)
        datamanager: Any = None
        model: Any = None

    class VanillaPipeline(torch.nn.Module):
        """Same constructor contract as nerfstudio 1.0.0's VanillaPipeline: the datamanager and the model are BUILT
        from their configs (`config.datamanager.setup(...)`, `config.model.setup(...)`)."""

        def __init__(self, config=None, device="cpu", test_mode="val", world_size=1, local_rank=0, grad_scaler=None):
            super().__init__()
            self.config = config
            self.test_mode = test_mode
            self.datamanager = config.datamanager.setup(device=device, test_mode=test_mode, world_size=world_size,
                                                        local_rank=local_rank)
            seed_pts = None
            dpo = getattr(self.datamanager, "train_dataparser_outputs", None)
            if dpo is not None and "points3D_xyz" in dpo.metadata:
                seed_pts = (dpo.metadata["points3D_xyz"], dpo.metadata["points3D_rgb"])
            assert self.datamanager.train_dataset is not None, "Missing input dataset"
            self._model = config.model.setup(scene_box=self.datamanager.train_dataset.scene_box,
                                             num_train_data=len(self.datamanager.train_dataset),
                                             metadata=self.datamanager.train_dataset.metadata, device=device,
                                             grad_scaler=grad_scaler, seed_points=seed_pts)
            self._model.to(device)
            self.world_size = world_size

        @property
        def model(self):
            return self._model
