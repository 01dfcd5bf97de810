"""dl4vc_b200 — B200-native (sm_100a) implementation of the DL4VC deep-averaging-network forward path.

Public surface: `Basic2DNet` (drop-in for the reference's dl4vc/model.py), `DanConfig`, the synthetic pileup
generator and the C-ABI loader. See DESIGN.md / INTEGRATION.md.
"""
from .config import DanConfig, prod_config, min_config, small_config  # noqa: F401

__all__ = ["DanConfig", "prod_config", "min_config", "small_config", "Basic2DNet"]


def __getattr__(name):
    if name == "Basic2DNet":
        from .model import Basic2DNet
        return Basic2DNet
    raise AttributeError(name)
