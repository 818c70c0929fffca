from dataclasses import dataclass, field
from typing import Any, Type

import torch

from nerfstudio.configs.base_config import InstantiateConfig


@dataclass
class VanillaPipelineConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: VanillaPipeline)
    datamanager: Any = None
    model: Any = None


class VanillaPipeline(torch.nn.Module):
    """nerfstudio 1.0.0 VanillaPipeline.__init__: datamanager and model are built from their CONFIGS."""

    def __init__(self, config, device, test_mode="val", world_size=1, local_rank=0, grad_scaler=None):
        super().__init__()
        self.config = config
        self.test_mode = test_mode
        self.datamanager = config.datamanager.setup(device=device, test_mode=test_mode, world_size=world_size,
                                                    local_rank=local_rank)
        seed_pts = None
        if (hasattr(self.datamanager, "train_dataparser_outputs")
                and "points3D_xyz" in self.datamanager.train_dataparser_outputs.metadata):
            md = self.datamanager.train_dataparser_outputs.metadata
            seed_pts = (md["points3D_xyz"], md["points3D_rgb"])
        self.datamanager.to(device)
        assert self.datamanager.train_dataset is not None, "Missing input dataset"
        self._model = config.model.setup(scene_box=self.datamanager.train_dataset.scene_box,
                                         num_train_data=len(self.datamanager.train_dataset),
                                         metadata=self.datamanager.train_dataset.metadata, device=device,
                                         grad_scaler=grad_scaler, seed_points=seed_pts)
        self.model.to(device)
        self.world_size = world_size

    @property
    def model(self):
        return self._model

    @property
    def device(self):
        return self.model.device
