from dataclasses import dataclass, field
from typing import Any, Dict, Optional, Type

from nerfstudio.configs.base_config import InstantiateConfig


@dataclass
class TrainerConfig(InstantiateConfig):
    _target: Type = field(default_factory=lambda: Trainer)
    method_name: Optional[str] = None
    steps_per_save: int = 1000
    steps_per_eval_batch: int = 500
    steps_per_eval_image: int = 500
    steps_per_eval_all_images: int = 25000
    max_num_iterations: int = 1000000
    mixed_precision: bool = False
    save_only_latest_checkpoint: bool = True
    gradient_accumulation_steps: Dict[str, int] = field(default_factory=dict)
    pipeline: Any = None
    optimizers: Dict[str, Any] = field(default_factory=dict)
    viewer: Any = None
    vis: str = "viewer"


class Trainer:
    def __init__(self, config, local_rank=0, world_size=1):
        self.config, self.local_rank, self.world_size = config, local_rank, world_size
        self.device = "cpu"
        self.grad_scaler = None
        self._start_step = 0
        self.steps_run = []

    def setup(self, test_mode="val"):
        self.pipeline = self.config.pipeline.setup(device=self.device, test_mode=test_mode, world_size=self.world_size,
                                                   local_rank=self.local_rank, grad_scaler=self.grad_scaler)

    def train(self):
        for step in range(self._start_step, self._start_step + self.config.max_num_iterations):
            self.steps_run.append(step)
