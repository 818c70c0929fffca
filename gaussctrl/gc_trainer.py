"""GaussCtrlTrainerConfig / GaussCtrlTrainer (gaussctrl/gc_trainer.py:41-49, 75-78, 176-187): nerfstudio's Trainer with
the two hot-path calls after setup and `render_rate` fine-tune iterations.  Needs nerfstudio; everything else
(optimizers, checkpoints, viewer, writers) is inherited, not re-implemented."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Type

from nerfstudio.engine.trainer import Trainer, TrainerConfig  # type: ignore


@dataclass
class GaussCtrlTrainerConfig(TrainerConfig):
    _target: Type = field(default_factory=lambda: GaussCtrlTrainer)
    steps_per_save: int = 500


class GaussCtrlTrainer(Trainer):
    def setup(self, test_mode="val") -> None:
        super().setup(test_mode)
        # gc_trainer.py:75-78: invert every view, then (when training, not when only viewing) edit them
        self.pipeline.render_reverse()
        if self.pipeline.test_mode == "val":
            self.pipeline.edit_images()

    def train(self) -> None:
        # gc_trainer.py:186-187: the fine-tune runs `pipeline.config.render_rate` iterations from the loaded step
        self.config.max_num_iterations = self.pipeline.config.render_rate
        super().train()
