"""nn.Module mirror of the reference's anti-aliased activation, backed by the fused CUDA kernel.

The module tree and buffer/parameter names are part of the checkpoint format (SURVEY.md section 1,
section 5): `act.alpha`, `act.beta`, `upsample.filter` [1,1,12], `downsample.lowpass.filter` [1,1,12].
Reference counterparts (paths in the reference tree, BigVGAN/):
    Activation1d     alias_free_activation/act.py:8-30
    UpSample1d       alias_free_activation/resample.py:10-38
    DownSample1d     alias_free_activation/resample.py:41-58
    LowPassFilter1d  alias_free_activation/filter.py:65-101
    kaiser_sinc_filter1d  alias_free_activation/filter.py:30-62

The resampler shells only own the filter buffers; their arithmetic happens inside the fused kernel.
Calling one of them on its own raises: this package has no un-fused or CPU path.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import functional as F_afa


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """Kaiser-windowed sinc low-pass, normalised to unit DC gain -> [1, 1, kernel_size] float32.

    Same torch ops in the same order as filter.py:30-62, so the taps are bit-identical to the
    reference's buffers on the same torch build.
    """
    half = kernel_size // 2
    attenuation = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    if attenuation > 50.0:
        kaiser_beta = 0.1102 * (attenuation - 8.7)
    elif attenuation >= 21.0:
        kaiser_beta = 0.5842 * (attenuation - 21) ** 0.4 + 0.07886 * (attenuation - 21.0)
    else:
        kaiser_beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=kaiser_beta, periodic=False)
    if kernel_size % 2 == 0:
        t = torch.arange(-half, half) + 0.5
    else:
        t = torch.arange(kernel_size) - half
    if cutoff == 0:
        taps = torch.zeros_like(t)
    else:
        taps = 2 * cutoff * window * torch.sinc(2 * cutoff * t)
        taps = taps / taps.sum()
    return taps.view(1, 1, kernel_size)


class _FusedOnly(nn.Module):
    def forward(self, *_args, **_kwargs):
        raise RuntimeError(
            f"{type(self).__name__} only holds the filter buffer of the fused Activation1d; "
            "its arithmetic runs inside the CUDA kernel (no stand-alone or CPU path)."
        )


class LowPassFilter1d(_FusedOnly):
    def __init__(self, cutoff=0.5, half_width=0.6, stride: int = 1, padding: bool = True,
                 padding_mode: str = "replicate", kernel_size: int = 12):
        super().__init__()
        if cutoff < -0.0:
            raise ValueError("Minimum cutoff must be larger than zero.")      # filter.py:79-80
        if cutoff > 0.5:
            raise ValueError("A cutoff above 0.5 does not make sense.")       # filter.py:81-82
        self.kernel_size = kernel_size
        self.even = kernel_size % 2 == 0
        self.pad_left = kernel_size // 2 - int(self.even)
        self.pad_right = kernel_size // 2
        self.stride = stride
        self.padding = padding
        self.padding_mode = padding_mode
        self.register_buffer("filter", kaiser_sinc_filter1d(cutoff, half_width, kernel_size))


class UpSample1d(_FusedOnly):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.stride = ratio
        self.pad = self.kernel_size // ratio - 1
        self.pad_left = self.pad * self.stride + (self.kernel_size - self.stride) // 2
        self.pad_right = self.pad * self.stride + (self.kernel_size - self.stride + 1) // 2
        self.register_buffer(
            "filter", kaiser_sinc_filter1d(cutoff=0.5 / ratio, half_width=0.6 / ratio, kernel_size=self.kernel_size)
        )


class DownSample1d(_FusedOnly):
    def __init__(self, ratio=2, kernel_size=None):
        super().__init__()
        self.ratio = ratio
        self.kernel_size = int(6 * ratio // 2) * 2 if kernel_size is None else kernel_size
        self.lowpass = LowPassFilter1d(cutoff=0.5 / ratio, half_width=0.6 / ratio, stride=ratio,
                                       kernel_size=self.kernel_size)


def _is_snake_like(act: nn.Module) -> bool:
    return isinstance(getattr(act, "alpha", None), torch.Tensor) and hasattr(act, "alpha_logscale")


class Activation1d(nn.Module):
    """Drop-in for `alias_free_activation.cuda.activation1d.Activation1d` (bigvgan.py:96-100).

    Same constructor as the reference's torch Activation1d (act.py:9-16), plus upstream's `fused`
    switch.  `activation` must be a Snake / SnakeBeta module (the only ones bigvgan.py:105-130
    builds): anything exposing `alpha` [C], `alpha_logscale`, and optionally `beta` [C].
    """

    def __init__(self, activation, up_ratio: int = 2, down_ratio: int = 2, up_kernel_size: int = 12,
                 down_kernel_size: int = 12, fused: bool = True):
        super().__init__()
        if not fused:
            raise NotImplementedError("fused=False is not available: this package ships only the fused CUDA path")
        if (up_ratio, down_ratio, up_kernel_size, down_kernel_size) != (2, 2, 12, 12):
            raise NotImplementedError(
                "the fused kernel implements up_ratio=down_ratio=2 with 12-tap filters "
                f"(every call site in bigvgan.py); got {(up_ratio, down_ratio, up_kernel_size, down_kernel_size)}"
            )
        if not _is_snake_like(activation):
            raise TypeError(
                f"fused Activation1d needs a Snake/SnakeBeta activation (alpha[, beta], alpha_logscale); got {type(activation).__name__}"
            )
        self.up_ratio = up_ratio
        self.down_ratio = down_ratio
        self.act = activation
        self.upsample = UpSample1d(up_ratio, up_kernel_size)
        self.downsample = DownSample1d(down_ratio, down_kernel_size)
        self._taps_key = None
        self._taps = None
        self._p32_key = None
        self._p32 = None

    # -- host copies of the filter buffers (they travel in checkpoints; read them, never recompute) --
    # The copy is refreshed when the buffers are REPLACED or RELOADED: a new storage (`.to()`, `.cuda()`, `.double()`, assigning a
    # new tensor) changes the key, `load_state_dict` and `_apply` drop the cache.  In-place writes with the values they already
    # hold -- DistributedDataParallel's broadcast_buffers=True does that before every forward (train_binaural_mel.py:541 uses the
    # default) -- do NOT trigger a re-read: each would be a synchronising device-to-host copy per module per step (measured on
    # two B200s: 64 ms instead of 44 ms per training step, profiles/r02_ddp_probe_n2.log).  Code that edits a filter in place
    # with NEW values calls `refresh_filters()`.
    def _host_taps(self):
        up, dn = self.upsample.filter, self.downsample.lowpass.filter
        key = (up.data_ptr(), dn.data_ptr(), up.dtype, up.device)
        if key != self._taps_key or self._taps is None:
            self._taps = (F_afa.host_taps(up), F_afa.host_taps(dn))
            self._taps_key = key
        return self._taps

    def refresh_filters(self):
        """Forget the cached host copy of the filter taps (after editing `upsample.filter` / `downsample.lowpass.filter` in place)."""
        self._taps_key = None
        self._taps = None

    def _load_from_state_dict(self, *args, **kwargs):
        self.refresh_filters()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *args, **kwargs):
        self.refresh_filters()
        return super()._apply(fn, *args, **kwargs)

    def _params(self):
        """alpha / beta as the autograd function wants them. fp32 parameters pass through untouched
        (so autograd reaches them); a model cast to bf16/fp16 for inference gets cached fp32 copies."""
        alpha = self.act.alpha
        beta = getattr(self.act, "beta", None)
        if alpha.dtype == torch.float32 or torch.is_grad_enabled() and alpha.requires_grad:
            return alpha, beta
        key = (alpha.data_ptr(), alpha._version, None if beta is None else (beta.data_ptr(), beta._version))
        if key != self._p32_key:
            self._p32 = (alpha.detach().float(), None if beta is None else beta.detach().float())
            self._p32_key = key
        return self._p32

    def forward(self, x):
        _, C, _ = x.shape                                                    # resample.py:30
        taps_up, taps_down = self._host_taps()
        alpha, beta = self._params()
        return F_afa.activation1d(x, alpha, beta, taps_up, taps_down, bool(self.act.alpha_logscale))
