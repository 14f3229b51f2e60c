"""afa_b200 -- host side of the B200-native fused anti-aliased activation (Activation1d).

Layout of this directory (it is a sys.path root, not itself a package; the hyphenated repo name is
not importable):

    csrc/                     CUDA kernels (sm_100a) + the C ABI of include/afa_b200.h
    afa_b200/                 ctypes binding, autograd function, nn.Module mirror, clip sharding
    alias_free_activation/    the module names the reference imports
                              (BigVGAN/bigvgan.py:19 and :96/196/274 in the reference tree)

There is no CPU fallback anywhere in this package: every compute entry point needs the built
library `afa_b200/libafa_sm100.so` and a CUDA device, and fails loudly otherwise.
"""
from .build import build_library, library_path  # noqa: F401
from ._lib import AfaError, load_library  # noqa: F401
from .functional import activation1d, activation1d_backward_raw, activation1d_forward_raw  # noqa: F401
from .modules import Activation1d, DownSample1d, LowPassFilter1d, UpSample1d, kaiser_sinc_filter1d  # noqa: F401
from .sharding import gather_waveforms, shard_indices  # noqa: F401

__all__ = [
    "Activation1d", "UpSample1d", "DownSample1d", "LowPassFilter1d", "kaiser_sinc_filter1d",
    "activation1d", "activation1d_forward_raw", "activation1d_backward_raw",
    "build_library", "library_path", "load_library", "AfaError",
    "shard_indices", "gather_waveforms",
]
