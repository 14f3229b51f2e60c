"""The import target of the reference's `use_cuda_kernel` switch.

Reference call sites (BigVGAN/bigvgan.py): lazy import at :96-100 (AMPBlock1), :196-200 (AMPBlock2),
:274-278 (BigVGAN), construction at :108-124 / :208-224 / :345, calls at :135-137, :234, :379.
"""
from afa_b200.modules import Activation1d  # noqa: F401

__all__ = ["Activation1d"]
