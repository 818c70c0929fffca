"""gaussctrl/gc_model.py of the reference -> B200-native implementation."""
from gaussctrl_b200.gc_model import GaussCtrlModel, GaussCtrlModelConfig  # noqa: F401
