from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Type

import torch

from nerfstudio.configs.base_config import InstantiateConfig


@dataclass
class NerfstudioDataParserConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Nerfstudio)
    load_3D_points: bool = False
    fake_num_images: int = 96      # FAKE-only knob: how many posed images the scene has (bear: 96)
    fake_hw: int = 32


class Nerfstudio:
    def __init__(self, config):
        self.config = config

    def get_dataparser_outputs(self, split="train"):
        from nerfstudio.cameras.cameras import Cameras
        n, hw = self.config.fake_num_images, self.config.fake_hw
        c2w = torch.eye(4)[:3].repeat(n, 1, 1)
        c2w[:, 0, 3] = torch.arange(n, dtype=torch.float32)
        meta = {}
        if self.config.load_3D_points:
            meta = {"points3D_xyz": torch.rand(123, 3), "points3D_rgb": torch.randint(0, 255, (123, 3))}
        return SimpleNamespace(image_filenames=[f"frame_{i:05d}.jpg" for i in range(n)],
                               cameras=Cameras(c2w, 30.0, 30.0, hw / 2, hw / 2, hw, hw), metadata=meta,
                               scene_box=SimpleNamespace(aabb=torch.tensor([[-1.0] * 3, [1.0] * 3])))
