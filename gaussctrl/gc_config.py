"""`gaussctrl_method` - the nerfstudio method registration of gaussctrl/gc_config.py:40-92 (same method name, trainer
settings, optimizer groups and learning rates), pointing at the B200-native pipeline / model.  The dataparser is
nerfstudio's own NerfstudioDataParser (the reference ships a copy of it with stage-A preload paths added;
`GaussCtrlPipeline.load_stage_a(root)` covers that use)."""
from __future__ import annotations

from nerfstudio.configs.base_config import ViewerConfig  # type: ignore
from nerfstudio.data.dataparsers.nerfstudio_dataparser import NerfstudioDataParserConfig  # type: ignore
from nerfstudio.engine.optimizers import AdamOptimizerConfig  # type: ignore
from nerfstudio.engine.schedulers import ExponentialDecaySchedulerConfig  # type: ignore
from nerfstudio.plugins.types import MethodSpecification  # type: ignore

from gaussctrl.gc_datamanager import GaussCtrlDataManager, GaussCtrlDataManagerConfig
from gaussctrl.gc_model import GaussCtrlModelConfig
from gaussctrl.gc_pipeline import GaussCtrlPipelineConfig
from gaussctrl.gc_trainer import GaussCtrlTrainerConfig


def _adam(lr: float):
    return AdamOptimizerConfig(lr=lr, eps=1e-15)


gaussctrl_method = MethodSpecification(
    config=GaussCtrlTrainerConfig(
        method_name="gaussctrl",
        steps_per_eval_image=100,
        steps_per_eval_batch=0,
        steps_per_save=250,
        max_num_iterations=1000,
        steps_per_eval_all_images=1000,
        save_only_latest_checkpoint=True,
        mixed_precision=False,
        gradient_accumulation_steps={"camera_opt": 100},
        pipeline=GaussCtrlPipelineConfig(
            datamanager=GaussCtrlDataManagerConfig(
                _target=GaussCtrlDataManager,
                dataparser=NerfstudioDataParserConfig(load_3D_points=True),
            ),
            model=GaussCtrlModelConfig(),
        ),
        optimizers={
            "xyz": {"optimizer": _adam(1.6e-4),
                    "scheduler": ExponentialDecaySchedulerConfig(lr_final=1.6e-6, max_steps=30000)},
            "features_dc": {"optimizer": _adam(0.0025), "scheduler": None},
            "features_rest": {"optimizer": _adam(0.0025 / 20), "scheduler": None},
            "opacity": {"optimizer": _adam(0.05), "scheduler": None},
            "scaling": {"optimizer": _adam(0.005), "scheduler": None},
            "rotation": {"optimizer": _adam(0.001), "scheduler": None},
            "camera_opt": {"optimizer": _adam(1e-3),
                           "scheduler": ExponentialDecaySchedulerConfig(lr_final=5e-5, max_steps=30000)},
        },
        viewer=ViewerConfig(num_rays_per_chunk=1 << 15),
        vis="viewer",
    ),
    description="GaussCtrl (B200-native editing hot path)",
)
