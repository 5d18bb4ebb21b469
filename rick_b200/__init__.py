"""rick_b200 -- B200-native (sm_100a) implementation of RICK's StyleGAN2 hot path.

Public surface mirrors the reference (yunqing-me/RICK):
  rick_b200.op        upfirdn2d, fused_leaky_relu, FusedLeakyReLU            (reference: op/)
  rick_b200.stylegan2 Generator, Discriminator, ModulatedConv2d, StyledConv, ToRGB, EqualConv2d, ...
                      (reference: gan_training/models/model_probe_tune.py)
  rick_b200.rick      Fisher accumulation -> per-filter FIM -> quantile -> freeze/prune masks -> mask application
                      (reference: train_dynamic_update_prune.py:214-393, 427-437, 521-539)
  rick_b200.adapt     the adaptation iteration (D step, R1, G step, path-length) and sample generation
                      (reference: train_dynamic_update_prune.py:396-589, gan_training/eval.py:31-46)
All compute goes through the C ABI of ``librick_b200.so`` (include/rick_b200.h); there is no CPU fallback.
"""
from . import _lib  # noqa: F401
from . import conv_tc  # noqa: F401  (registers rick_conv_tc with the ctypes table)

__version__ = "0.1.0"
