from dataclasses import dataclass, field
from typing import Any, Type

import torch

from nerfstudio.configs.base_config import InstantiateConfig


def _undistort_image(*args, **kwargs):
    raise NotImplementedError


@dataclass
class FullImageDatamanagerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: FullImageDatamanager)
    dataparser: Any = None
    camera_res_scale_factor: float = 1.0
    cache_images: str = "cpu"
    cache_images_type: str = "float32"


class FullImageDatamanager(torch.nn.Module):
    """Builds train_dataset from config.dataparser and caches its images (all zeros here)."""

    def __init__(self, config, device="cpu", test_mode="val", world_size=1, local_rank=0, **kwargs):
        super().__init__()
        assert config.dataparser is not None, "FullImageDatamanager needs a dataparser config"
        self.config, self.device, self.test_mode = config, device, test_mode
        self.world_size, self.local_rank = world_size, local_rank
        self.dataparser = config.dataparser.setup()
        self.train_dataparser_outputs = self.dataparser.get_dataparser_outputs(split="train")
        dpo = self.train_dataparser_outputs
        n = len(dpo.image_filenames)
        hw = int(dpo.cameras.height[0].item())
        self.train_dataset = _Dataset(dpo, n)
        self.eval_dataset = None
        self.cached_train = [{"image": torch.zeros(hw, hw, 3), "image_idx": i} for i in range(n)]
        self.cached_eval = []


class _Dataset:
    def __init__(self, dpo, n):
        self.cameras, self.scene_box, self.metadata, self._dataparser_outputs, self._n = \
            dpo.cameras, dpo.scene_box, dpo.metadata, dpo, n

    def __len__(self):
        return self._n
