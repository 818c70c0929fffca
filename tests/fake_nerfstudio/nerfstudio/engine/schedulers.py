from dataclasses import dataclass


@dataclass
class ExponentialDecaySchedulerConfig:
    lr_final: float = 0.0
    max_steps: int = 100000
