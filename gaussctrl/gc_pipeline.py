"""gaussctrl/gc_pipeline.py of the reference -> B200-native implementation."""
from gaussctrl_b200.gc_pipeline import GaussCtrlPipeline, GaussCtrlPipelineConfig  # noqa: F401
