"""gaussctrl/utils.py of the reference -> drop-in attention processor on the sm_100a kernels."""
from gaussctrl_b200.utils import CrossViewAttnProcessor, compute_attn, read_depth2disparity  # noqa: F401
