"""gaussctrl/gc_datamanager.py of the reference -> host-side mirror."""
from gaussctrl_b200.gc_datamanager import GaussCtrlDataManager, GaussCtrlDataManagerConfig  # noqa: F401
