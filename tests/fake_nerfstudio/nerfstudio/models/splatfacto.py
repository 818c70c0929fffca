from dataclasses import dataclass, field
from typing import Type

import torch

from nerfstudio.configs.base_config import InstantiateConfig


@dataclass
class SplatfactoModelConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: SplatfactoModel)
    sh_degree: int = 3
    sh_degree_interval: int = 1000
    background_color: str = "random"
    ssim_lambda: float = 0.2
    use_scale_regularization: bool = False
    max_gauss_ratio: float = 10.0
    stop_split_at: int = 15000
    num_downscales: int = 2
    resolution_schedule: int = 3000


class SplatfactoModel(torch.nn.Module):
    """nerfstudio Model contract: __init__(config, scene_box, num_train_data, **kwargs) -> populate_modules()."""

    def __init__(self, config, scene_box, num_train_data, seed_points=None, metadata=None, device=None, grad_scaler=None,
                 **kwargs):
        super().__init__()
        assert scene_box is not None and num_train_data > 0, "VanillaPipeline passes scene_box / num_train_data"
        self.config, self.scene_box, self.num_train_data, self.seed_points = config, scene_box, num_train_data, seed_points
        n = seed_points[0].shape[0] if seed_points is not None else 50
        P = torch.nn.Parameter
        self.means = P(seed_points[0].clone().float() if seed_points is not None else torch.zeros(n, 3))
        self.scales, self.quats = P(torch.zeros(n, 3)), P(torch.zeros(n, 4))
        self.features_dc, self.features_rest = P(torch.zeros(n, 3)), P(torch.zeros(n, 15, 3))
        self.opacities = P(torch.zeros(n, 1))
        self.background_color = torch.tensor([0.1, 0.2, 0.3])
        self.crop_box = None
        self.step = 0
        self.xys = self.radii = self.last_size = None

    @property
    def device(self):
        return self.means.device

    def _get_downscale_factor(self):
        if self.training:
            return 2 ** max(self.config.num_downscales - self.step // self.config.resolution_schedule, 0)
        return 1

    def set_crop(self, crop_box):
        self.crop_box = crop_box

    def get_gt_img(self, image):
        return image.to(self.device)

    def forward(self, camera):
        return self.get_outputs(camera)

    def after_train(self, step):
        """The densification callback's precondition (splatfacto.after_train)."""
        assert step == self.step
        if self.step >= self.config.stop_split_at:
            return
        assert self.xys.grad is not None, "xys.grad is None: get_outputs must call self.xys.retain_grad() in training"
