"""Drop-in operator surface of the reference's ``op`` package (op/__init__.py:1-2), backed by sm_100a kernels."""
from .fused_act import FusedLeakyReLU, FusedLeakyReLU_kml, fused_leaky_relu
from .upfirdn2d import upfirdn2d

__all__ = ["FusedLeakyReLU", "FusedLeakyReLU_kml", "fused_leaky_relu", "upfirdn2d"]
