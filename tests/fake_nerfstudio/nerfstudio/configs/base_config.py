from dataclasses import dataclass, field
from typing import Any, Type


@dataclass
class InstantiateConfig:
    _target: Type

    def setup(self, **kwargs) -> Any:
        return self._target(self, **kwargs)


@dataclass
class ViewerConfig:
    num_rays_per_chunk: int = 32768
    quit_on_train_completion: bool = True
