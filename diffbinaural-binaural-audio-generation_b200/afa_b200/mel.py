"""Fused log-mel spectrogram on sm_100a behind the reference's own function signatures.

    mel_spectrogram(...)                 <->  BigVGAN/meldataset.py:51-123 (same arguments, same [B, n_mels, frames] result)
    get_mel_spectrogram(wav, h)          <->  BigVGAN/meldataset.py:126-146
    MultiScaleMelSpectrogramLoss(...)    <->  BigVGAN/loss.py:23-211 (same constructor, same scalar loss)

One `afa_logmel_fwd` launch (csrc/afa_mel.cu) replaces pad -> torch.stft -> pow/sum/sqrt -> matmul -> clamp -> log;
`afa_logmel_bwd` (two launches) is its adjoint, wired in as a torch.autograd.Function so that both functions are
differentiable with respect to the waveform, as the training step needs (train_binaural_mel.py:711-787).
There is no CPU or eager fallback: CPU tensors raise.

The filterbank comes from `librosa.filters.mel` in the reference (third-party, absent from this image); the
`slaney_mel_filterbank` below follows librosa's published algorithm (htk=False, norm='slaney') and a caller that has
its own basis (a checkpointed one, or librosa's) passes it as `mel_basis=`.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from ._lib import check, load_library  # noqa: F401  (check is re-exported for callers of the raw C ABI)

AFA_MEL_PAD_REFLECT = 0
AFA_MEL_PAD_ZERO = 1
AFA_MEL_FLAG_RAW = 1
AFA_MEL_FLAG_L1_SIGN = 2
AFA_MEL_FLAG_ACCUMULATE = 4
L1_PARTIALS = 256          # per-CTA partial sums of one L1 (finished by one torch reduction over all scales)


def slaney_mel_filterbank(sr: float, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """float32 [n_mels, 1 + n_fft // 2]: triangular filters on Slaney's mel scale (linear to 1 kHz, then 27 steps per
    factor 6.4), each normalised to unit area in Hz -- what `librosa_mel_fn(sr=, n_fft=, n_mels=, fmin=, fmax=)` gives
    at meldataset.py:89-91 and loss.py:106-108."""
    fmax = float(sr) / 2.0 if fmax is None else float(fmax)
    step_hz, knee_hz, log_step = 200.0 / 3.0, 1000.0, math.log(6.4) / 27.0
    knee_mel = knee_hz / step_hz

    def to_mel(f):
        return f / step_hz if f < knee_hz else knee_mel + math.log(f / knee_hz) / log_step

    def to_hz(m):
        return np.where(m < knee_mel, m * step_hz, knee_hz * np.exp(log_step * (m - knee_mel)))

    edges = to_hz(np.linspace(to_mel(float(fmin)), to_mel(fmax), n_mels + 2))       # n_mels + 2 band edges in Hz
    bins = np.linspace(0.0, float(sr) / 2.0, 1 + n_fft // 2)
    rising = (bins[None, :] - edges[:-2, None]) / (edges[1:-1] - edges[:-2])[:, None]
    falling = (edges[2:, None] - bins[None, :]) / (edges[2:] - edges[1:-1])[:, None]
    tri = np.clip(np.minimum(rising, falling), 0.0, None)
    return (tri * (2.0 / (edges[2:] - edges[:-2]))[:, None]).astype(np.float32)


def _fp(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _ip(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int32))


class MelPlan:
    """Device-resident constants of one (n_fft, window, mel basis) configuration: window, FFT twiddles and the banded
    form of the basis.  Built once per configuration and device, like the reference's mel_basis_cache /
    hann_window_cache (meldataset.py:47-48, 86-93)."""

    def __init__(self, n_fft: int, window: torch.Tensor, mel_basis, device):
        if n_fft & (n_fft - 1) or not 32 <= n_fft <= 2048:
            raise ValueError(f"n_fft must be a power of two in [32, 2048], got {n_fft} (no other STFT size is built)")
        window = window.detach().to(torch.float64).cpu()
        if window.numel() > n_fft:
            raise ValueError("win_size must not exceed n_fft")
        if window.numel() < n_fft:                      # torch.stft centres a short window inside n_fft
            left = (n_fft - window.numel()) // 2
            window = torch.nn.functional.pad(window, (left, n_fft - window.numel() - left))
        basis = np.asarray(mel_basis.detach().cpu() if isinstance(mel_basis, torch.Tensor) else mel_basis, dtype=np.float32)
        if basis.ndim != 2 or basis.shape[1] != n_fft // 2 + 1:
            raise ValueError(f"mel basis must be [n_mels, {n_fft // 2 + 1}], got {basis.shape}")
        starts, lens, offs, weights = banded(basis)
        t = np.arange(n_fft // 2, dtype=np.float64) * (2.0 * np.pi / n_fft)
        tw = np.stack([np.cos(t), -np.sin(t)], axis=1).astype(np.float32)
        self.n_fft = n_fft
        self.n_mels = int(basis.shape[0])
        self.device = torch.device(device)
        self.window = window.to(torch.float32).to(self.device)
        self.twiddle = torch.from_numpy(tw).to(self.device)
        self.band_start = torch.from_numpy(starts).to(self.device)
        self.band_len = torch.from_numpy(lens).to(self.device)
        self.band_off = torch.from_numpy(offs).to(self.device)
        self.band_w = torch.from_numpy(weights).to(self.device)
        mlo, mhi = bin_cover(starts, lens, n_fft // 2 + 1)
        self.bin_mlo = torch.from_numpy(mlo).to(self.device)
        self.bin_mhi = torch.from_numpy(mhi).to(self.device)
        # the constant tail of every launch's argument list, cast once (a launch is a few microseconds of GPU time:
        # per-call ctypes casts would cost more than the kernel)
        self._c_basis = (_fp(self.window), _fp(self.twiddle), self.n_mels, _ip(self.band_start), _ip(self.band_len),
                         _ip(self.band_off), _fp(self.band_w))
        self._c_cover = (_ip(self.bin_mlo), _ip(self.bin_mhi))


def banded(basis: np.ndarray):
    """Dense [n_mels, n_freq] -> (first bin, run length, offset, packed weights): every row's non-zero support as one
    contiguous run (interior zeros, if any, stay in the run, so the contraction is exact)."""
    n_mels = basis.shape[0]
    starts = np.zeros(n_mels, dtype=np.int32)
    lens = np.zeros(n_mels, dtype=np.int32)
    offs = np.zeros(n_mels, dtype=np.int32)
    packed = []
    total = 0
    for m in range(n_mels):
        nz = np.flatnonzero(basis[m])
        if nz.size:
            starts[m], lens[m] = nz[0], nz[-1] - nz[0] + 1
            packed.append(basis[m, nz[0]:nz[-1] + 1])
        offs[m] = total
        total += int(lens[m])
    weights = np.concatenate(packed).astype(np.float32) if packed else np.zeros(1, dtype=np.float32)
    return starts, lens, offs, np.ascontiguousarray(weights)


def bin_cover(starts: np.ndarray, lens: np.ndarray, n_freq: int):
    """For every bin k: [mlo, mhi) bounding the filters whose run contains k (the backward kernel walks this range and
    checks membership, so any basis is handled; for a mel basis the range is 1-3 filters)."""
    mlo = np.zeros(n_freq, dtype=np.int32)
    mhi = np.zeros(n_freq, dtype=np.int32)
    first = np.full(n_freq, -1, dtype=np.int64)
    last = np.full(n_freq, -1, dtype=np.int64)
    for m in range(len(starts)):
        a, b = int(starts[m]), int(starts[m] + lens[m])
        if b > a:
            seg = first[a:b]
            seg[seg < 0] = m
            last[a:b] = m
    has = first >= 0
    mlo[has] = first[has]
    mhi[has] = last[has] + 1
    return mlo, mhi


def num_frames(T: int, n_fft: int, hop: int, pad: int) -> int:
    """Frames of a T-sample row: what afa_logmel_num_frames() returns (tests check the two agree)."""
    if T <= 0 or n_fft <= 0 or hop <= 0 or pad < 0 or T + 2 * pad < n_fft:
        return 0
    return 1 + (T + 2 * pad - n_fft) // hop


class _on_device:
    """torch.cuda.device(), entered only when the tensor's device is not already current."""

    def __init__(self, device):
        self.ctx = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *a):
        if self.ctx is not None:
            self.ctx.__exit__(*a)


def _check_wav(wav: torch.Tensor, plan: MelPlan, pad: int, pad_mode: int) -> torch.Tensor:
    if not wav.is_cuda:
        raise RuntimeError("afa_b200.mel: the fused log-mel kernel needs a CUDA tensor (there is no CPU fallback)")
    if wav.dtype != torch.float32 or wav.dim() != 2:
        raise TypeError(f"afa_b200.mel: expected a float32 [rows, T] tensor, got {wav.dtype} {tuple(wav.shape)}")
    if wav.device != plan.device:
        raise RuntimeError(f"afa_b200.mel: plan lives on {plan.device}, waveform on {wav.device}")
    if wav.stride(1) != 1 or (wav.shape[0] > 1 and wav.stride(0) < wav.shape[1]):
        wav = wav.contiguous()
    if pad_mode == AFA_MEL_PAD_REFLECT and pad >= wav.shape[1]:
        raise RuntimeError(f"Padding size should be less than the corresponding input dimension, but got: padding ({pad}, {pad}) "
                           f"at dimension 1 of input {tuple(wav.shape)}")   # F.pad's message for the same input
    return wav


def logmel_forward_raw(wav: torch.Tensor, plan: MelPlan, hop: int, pad: int, pad_mode: int, mag_eps: float, clamp_eps: float,
                       log_scale: float, raw: bool, out: torch.Tensor | None = None) -> torch.Tensor:
    """One launch of afa_logmel_fwd on the current stream (no autograd)."""
    wav = _check_wav(wav, plan, pad, pad_mode)
    rows, T = wav.shape
    nf = num_frames(T, plan.n_fft, hop, pad)
    if out is None:
        out = torch.empty(rows, plan.n_mels, nf, device=wav.device, dtype=torch.float32)
    elif tuple(out.shape) != (rows, plan.n_mels, nf) or out.dtype != torch.float32 or not out.is_contiguous():
        raise ValueError("afa_b200.mel: `out` must be a contiguous float32 [rows, n_mels, n_frames] tensor")
    with _on_device(wav.device):
        rc = load_library().afa_logmel_fwd(
            _fp(wav), _fp(out), rows, T, wav.stride(0) if rows > 1 else T, plan.n_fft, hop, pad, pad_mode,
            *plan._c_basis, mag_eps, clamp_eps, log_scale, AFA_MEL_FLAG_RAW if raw else 0,
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(rc, "afa_logmel_fwd")
    return out


def logmel_backward_raw(wav: torch.Tensor, gout: torch.Tensor, plan: MelPlan, hop: int, pad: int, pad_mode: int, mag_eps: float,
                        clamp_eps: float, log_scale: float, raw: bool, *, l1_other: torch.Tensor | None = None, l1_coef: float = 0.0,
                        l1_scale: torch.Tensor | None = None, accumulate_into: torch.Tensor | None = None) -> torch.Tensor:
    """d loss / d wav ([rows, T]) from gout = d loss / d out: afa_logmel_bwd (two launches) on the current stream.

    With `l1_other`, `gout` / `l1_other` are the two log-mel tensors of an L1 loss and the output gradient is
    sign(gout - l1_other) * l1_coef * l1_scale (a device scalar), formed inside the kernel (AFA_MEL_FLAG_L1_SIGN).
    `accumulate_into`: add to this [rows, T] tensor instead of writing a new one (AFA_MEL_FLAG_ACCUMULATE)."""
    wav = _check_wav(wav, plan, pad, pad_mode)
    rows, T = wav.shape
    nf = num_frames(T, plan.n_fft, hop, pad)
    if tuple(gout.shape) != (rows, plan.n_mels, nf):
        raise ValueError(f"afa_b200.mel: gradient of shape {tuple(gout.shape)}, expected {(rows, plan.n_mels, nf)}")
    gout = gout.to(torch.float32).contiguous()
    flags = AFA_MEL_FLAG_RAW if raw else 0
    other_p, scale_p = None, None
    if l1_other is not None:
        if l1_other.shape != gout.shape or l1_other.dtype != torch.float32 or not l1_other.is_contiguous():
            raise ValueError("afa_b200.mel: l1_other must match gout (contiguous float32)")
        flags |= AFA_MEL_FLAG_L1_SIGN
        other_p = _fp(l1_other)
        if l1_scale is not None:
            l1_scale = l1_scale.to(device=wav.device, dtype=torch.float32).reshape(1).contiguous()
            scale_p = _fp(l1_scale)
    if accumulate_into is None:
        gwav = torch.empty(rows, T, device=wav.device, dtype=torch.float32)
    else:
        gwav = accumulate_into
        if tuple(gwav.shape) != (rows, T) or gwav.dtype != torch.float32 or not gwav.is_contiguous():
            raise ValueError("afa_b200.mel: accumulate_into must be a contiguous float32 [rows, T] tensor")
        flags |= AFA_MEL_FLAG_ACCUMULATE
    lib = load_library()
    nbytes = rows * nf * plan.n_fft * 4                     # == afa_logmel_bwd_workspace_bytes(rows, T, n_fft, hop, pad)
    ws = torch.empty(max(nbytes // 4, 2), device=wav.device, dtype=torch.float32)
    with _on_device(wav.device):
        rc = lib.afa_logmel_bwd(
            _fp(wav), _fp(gout), _fp(gwav), rows, T, wav.stride(0) if rows > 1 else T, T, plan.n_fft, hop, pad, pad_mode,
            *plan._c_basis, *plan._c_cover, mag_eps, clamp_eps, log_scale, flags, other_p, l1_coef, scale_p,
            ctypes.c_void_p(ws.data_ptr()), ws.numel() * 4,
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(rc, "afa_logmel_bwd")
    return gwav


def l1_partial_sums(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[:] = per-CTA partial sums of |a - b| (afa_l1_partial_sums; `out` is a contiguous float32 CUDA vector)."""
    if a.shape != b.shape or a.dtype != torch.float32 or b.dtype != torch.float32 or not (a.is_contiguous() and b.is_contiguous()):
        raise ValueError("afa_b200.mel.l1_partial_sums: two contiguous float32 tensors of one shape expected")
    if not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous() or out.dim() != 1:
        raise ValueError("afa_b200.mel.l1_partial_sums: `out` must be a contiguous float32 CUDA vector")
    with _on_device(a.device):
        rc = load_library().afa_l1_partial_sums(_fp(a), _fp(b), a.numel(), _fp(out), out.numel(),
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(rc, "afa_l1_partial_sums")
    return out


class _LogMelFn(torch.autograd.Function):
    """Saves only the waveform; the backward recomputes the spectrum inside afa_logmel_bwd."""

    @staticmethod
    def forward(ctx, wav, plan, hop, pad, pad_mode, mag_eps, clamp_eps, log_scale, raw):
        ctx.save_for_backward(wav)
        ctx.cfg = (plan, hop, pad, pad_mode, mag_eps, clamp_eps, log_scale, raw)
        return logmel_forward_raw(wav, plan, hop, pad, pad_mode, mag_eps, clamp_eps, log_scale, raw)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        (wav,) = ctx.saved_tensors
        return (logmel_backward_raw(wav, gout, *ctx.cfg),) + (None,) * 8


def logmel(wav: torch.Tensor, plan: MelPlan, hop: int, pad: int, pad_mode: int = AFA_MEL_PAD_REFLECT, mag_eps: float = 1e-9,
           clamp_eps: float = 1e-5, log_scale: float = 1.0, raw: bool = False, out: torch.Tensor | None = None) -> torch.Tensor:
    """[rows, T] float32 CUDA -> [rows, n_mels, n_frames] float32; differentiable with respect to `wav`
    (pass `out=` only outside autograd, e.g. under CUDA-graph capture with preallocated buffers)."""
    if out is None and torch.is_grad_enabled() and wav.requires_grad:
        return _LogMelFn.apply(wav, plan, hop, pad, pad_mode, mag_eps, clamp_eps, log_scale, raw)
    if out is not None and torch.is_grad_enabled() and wav.requires_grad:
        raise RuntimeError("afa_b200.mel.logmel: `out=` cannot be combined with autograd")
    return logmel_forward_raw(wav.detach(), plan, hop, pad, pad_mode, mag_eps, clamp_eps, log_scale, raw, out)


# --------------------------------------------------------------------------------------
# the reference's public functions
# --------------------------------------------------------------------------------------
mel_plan_cache: dict = {}


def mel_spectrogram(y: torch.Tensor, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int, win_size: int, fmin: int,
                    fmax: int = None, center: bool = False, *, mel_basis=None, check_range: bool = True) -> torch.Tensor:
    """Drop-in for BigVGAN/meldataset.py:51-123: log(clamp(mel_basis @ |STFT(y)|, 1e-5)).

    y: [B, T] (reflect-padded by (n_fft - hop_size) // 2, :98-101) or [T] (zero-padded, :96-97; the result then has
    no batch dimension, as torch.stft returns for 1-D input).  `check_range=False` skips the reference's two
    out-of-[-1, 1] warnings (:78-81), which cost a device synchronisation each."""
    if center:
        raise NotImplementedError("center=True is not used by the reference (meldataset.py:110 passes center=False)")
    if check_range:
        lo, hi = torch.aminmax(y)
        if lo < -1.0:
            print(f"[WARNING] Min value of input waveform signal is {lo}")
        if hi > 1.0:
            print(f"[WARNING] Max value of input waveform signal is {hi}")
    key = f"{n_fft}_{num_mels}_{sampling_rate}_{hop_size}_{win_size}_{fmin}_{fmax}_{y.device}"
    if mel_basis is not None:
        # a caller-supplied basis (a checkpointed one, librosa's): cached by identity, so the plan (banded tables, host-to-device
        # copies) is not rebuilt on every call; the basis object is kept alive by the cache entry
        key += f"_basis{id(mel_basis)}"
    entry = mel_plan_cache.get(key)
    plan = entry[0] if entry is not None and entry[1] is mel_basis else None
    if plan is None:
        basis = slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax) if mel_basis is None else mel_basis
        plan = MelPlan(n_fft, torch.hann_window(win_size, dtype=torch.float64), basis, y.device)
        mel_plan_cache[key] = (plan, mel_basis)
    pad = (n_fft - hop_size) // 2
    if y.dim() == 1:
        return logmel(y.unsqueeze(0), plan, hop_size, pad, AFA_MEL_PAD_ZERO)[0]
    if y.dim() != 2:
        raise RuntimeError(f"mel_spectrogram expects a [B, T] or [T] waveform, got {tuple(y.shape)}")
    return logmel(y, plan, hop_size, pad, AFA_MEL_PAD_REFLECT)


def get_mel_spectrogram(wav: torch.Tensor, h) -> torch.Tensor:
    """Drop-in for BigVGAN/meldataset.py:126-146 (called per channel at BigVGAN/inference_binaural.py:131-132):
    `h` carries n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax."""
    return mel_spectrogram(wav, h.n_fft, h.num_mels, h.sampling_rate, h.hop_size, h.win_size, h.fmin, h.fmax)


class _MultiScaleL1Fn(torch.autograd.Function):
    """sum over scales of weight * mean |log10 mel(x) - log10 mel(y)| with the L1 inside the kernels: per scale two
    afa_logmel_fwd launches and one afa_l1_partial_sums forward, one afa_logmel_bwd (sign formed in-kernel, waveform
    gradients of the scales accumulated in place) per differentiated input backward."""

    @staticmethod
    def forward(ctx, x, y, module):
        B, C, T = x.shape
        x2, y2 = x.reshape(B * C, T), y.reshape(B * C, T)
        n_scales = len(module.window_lengths)
        partials = torch.empty(n_scales, L1_PARTIALS, device=x.device, dtype=torch.float32)
        coefs, saved = [], []
        for s in range(n_scales):
            w = module.window_lengths[s]
            cfg = (module._plan(s, x.device), w // 4, w // 2, AFA_MEL_PAD_REFLECT, 0.0, module.clamp_eps, 1.0 / math.log(10.0), False)
            lx, ly = logmel_forward_raw(x2, *cfg), logmel_forward_raw(y2, *cfg)
            l1_partial_sums(lx, ly, partials[s])
            coefs.append((module.log_weight + module.mag_weight) / max(lx.numel(), 1))   # both terms compare the log mels (loss.py:206-207)
            saved += [lx, ly]
        ctx.module, ctx.coefs, ctx.shape = module, coefs, (B, C, T)
        ctx.save_for_backward(x2, y2, *saved)
        # the weights are part of the key: forward must scale with the same coefficients backward reads from ctx.coefs
        key = (str(x.device), B * C, T, float(module.log_weight), float(module.mag_weight), float(module.clamp_eps))
        if key not in module._coefs:                            # built once per shape: no host-to-device copy in the step (graph capture)
            module._coefs[key] = torch.tensor(coefs, device=x.device, dtype=torch.float32)
        return (partials.sum(dim=1) * module._coefs[key]).sum()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gloss):
        x2, y2, *saved = ctx.saved_tensors
        module, (B, C, T) = ctx.module, ctx.shape
        grads = []
        for which, wav in ((0, x2), (1, y2)):
            if not ctx.needs_input_grad[which]:
                grads.append(None)
                continue
            g = None
            for s, w in enumerate(module.window_lengths):
                lx, ly = saved[2 * s], saved[2 * s + 1]
                a, b = (lx, ly) if which == 0 else (ly, lx)
                cfg = (module._plan(s, wav.device), w // 4, w // 2, AFA_MEL_PAD_REFLECT, 0.0, module.clamp_eps, 1.0 / math.log(10.0), False)
                g = logmel_backward_raw(wav, a, *cfg, l1_other=b, l1_coef=ctx.coefs[s], l1_scale=gloss, accumulate_into=g)
            grads.append(g.view(B, C, T))
        return grads[0], grads[1], None


class MultiScaleMelSpectrogramLoss(torch.nn.Module):
    """BigVGAN/loss.py:23-211 with the fused kernels: seven (window, n_mels) scales, log10 of the clamped mels, L1.
    Differentiable with respect to both waveforms (the L1 and the sum stay torch ops on the small log-mel tensors)."""

    def __init__(self, sampling_rate: int, n_mels=(5, 10, 20, 40, 80, 160, 320), window_lengths=(32, 64, 128, 256, 512, 1024, 2048),
                 loss_fn=None, clamp_eps: float = 1e-5, mag_weight: float = 0.0, log_weight: float = 1.0, pow: float = 1.0,
                 weight: float = 1.0, match_stride: bool = False, mel_fmin=(0, 0, 0, 0, 0, 0, 0),
                 mel_fmax=(None, None, None, None, None, None, None), window_type: str = "hann"):
        super().__init__()
        if match_stride or window_type != "hann" or pow != 1.0:
            raise NotImplementedError("only the configuration the reference trains with is built "
                                      "(match_stride=False, hann, pow=1.0; train_binaural_mel.py:458-460)")
        self.sampling_rate = sampling_rate
        self.n_mels = list(n_mels)
        self.window_lengths = list(window_lengths)
        self.loss_fn = loss_fn if loss_fn is not None else torch.nn.L1Loss()
        self.clamp_eps, self.mag_weight, self.log_weight, self.weight = clamp_eps, mag_weight, log_weight, weight
        self.mel_fmin, self.mel_fmax = list(mel_fmin), list(mel_fmax)
        self._plans: dict = {}
        self._coefs: dict = {}

    def _plan(self, scale: int, device) -> MelPlan:
        key = (scale, str(device))
        if key not in self._plans:
            w, nm = self.window_lengths[scale], self.n_mels[scale]
            basis = slaney_mel_filterbank(self.sampling_rate, w, nm, self.mel_fmin[scale], self.mel_fmax[scale])
            self._plans[key] = MelPlan(w, torch.hann_window(w, dtype=torch.float64), basis, device)
        return self._plans[key]

    def log_mels(self, wav: torch.Tensor, scale: int) -> torch.Tensor:
        """[B, C, T] -> [B, C, n_mels, frames]: log10(clamp(mel, clamp_eps)) of one scale (loss.py:195-200)."""
        B, C, T = wav.shape
        w = self.window_lengths[scale]
        out = logmel(wav.reshape(B * C, T), self._plan(scale, wav.device), w // 4, w // 2, AFA_MEL_PAD_REFLECT, mag_eps=0.0,
                     clamp_eps=self.clamp_eps, log_scale=1.0 / math.log(10.0))
        return out.view(B, C, out.shape[1], out.shape[2])

    def forward(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if type(self.loss_fn) is torch.nn.L1Loss and self.loss_fn.reduction == "mean" and x.shape == y.shape and x.dim() == 3 \
                and x.dtype == y.dtype == torch.float32:
            return _MultiScaleL1Fn.apply(x, y, self)          # the default loss_fn: L1 fused into the kernels
        losses = []
        for s in range(len(self.window_lengths)):
            lx, ly = self.log_mels(x, s), self.log_mels(y, s)
            losses.append(self.log_weight * self.loss_fn(lx, ly))
            losses.append(self.mag_weight * self.loss_fn(lx, ly))     # the reference compares the log mels in both terms (:206-207)
        return sum(losses)
