"""Torch-op restatement of the reference's CPU path (TEST INFRASTRUCTURE ONLY).

The reference computes Activation1d with six ATen ops.  /root/reference does not
exist on the GPU box, so this module states the same op sequence functionally, for two uses only:

  * tests/ and __graft_entry__.smoke(): the fp32 "torch oracle" the CUDA kernel is compared with
    on identical inputs (runs on CPU or on the GPU box's device, fp32 or fp64);
  * bench.py's `cpu_baseline` leg and `--impl reference`: the reference's torch CPU path timed on
    the box's host cores (kind = "port").

Sequence restated (file:line in /root/reference/BigVGAN):
    alias_free_activation/resample.py:32   F.pad(x, (5, 5), "replicate")
    alias_free_activation/resample.py:33-35  2 * F.conv_transpose1d(x, taps.expand(C,1,12), stride=2, groups=C)
    alias_free_activation/resample.py:36   crop [15:-15]
    activations.py:119-124                 exp(alpha), exp(beta); x + 1/(beta+1e-9) * sin(x*alpha)**2
    alias_free_activation/filter.py:98     F.pad(x, (5, 6), "replicate")
    alias_free_activation/filter.py:99     F.conv1d(x, taps.expand(C,1,12), stride=2, groups=C)

tests/test_oracle.py checks this file against the golden vectors produced by the unmodified
reference (tests/golden/make_golden.py).  The product path never imports it.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 0.000000001  # activations.py:49 / :111 `no_div_by_zero`


def make_taps(dtype=torch.float32) -> torch.Tensor:
    """12-tap Kaiser-windowed sinc, cutoff 0.25, half-width 0.3 (filter.py:30-62 via resample.py:23-25)."""
    import math

    k, cutoff, half_width = 12, 0.25, 0.3
    amp = 2.285 * (k // 2 - 1) * math.pi * (4 * half_width) + 7.95
    if amp > 50.0:
        kaiser_beta = 0.1102 * (amp - 8.7)
    elif amp >= 21.0:
        kaiser_beta = 0.5842 * (amp - 21) ** 0.4 + 0.07886 * (amp - 21.0)
    else:
        kaiser_beta = 0.0
    win = torch.kaiser_window(k, beta=kaiser_beta, periodic=False)
    t = torch.arange(-(k // 2), k // 2) + 0.5
    taps = 2 * cutoff * win * torch.sinc(2 * cutoff * t)
    taps = taps / taps.sum()
    return taps.to(dtype).view(1, 1, k)


def activation1d_torch(x, alpha, beta, logscale: bool, taps_up, taps_down):
    """down(act(up(x))) with torch ops. x [B,C,T]; alpha/beta raw [C] (beta None => Snake)."""
    C = x.shape[1]
    u = F.pad(x, (5, 5), mode="replicate")
    u = 2 * F.conv_transpose1d(u, taps_up.expand(C, -1, -1), stride=2, groups=C)
    u = u[..., 15:-15]
    a = alpha.unsqueeze(0).unsqueeze(-1)
    b = a if beta is None else beta.unsqueeze(0).unsqueeze(-1)
    if logscale:
        a = torch.exp(a)
        b = a if beta is None else torch.exp(b)
    s = u + (1.0 / (b + EPS)) * torch.pow(torch.sin(u * a), 2)
    s = F.pad(s, (5, 6), mode="replicate")
    return F.conv1d(s, taps_down.expand(C, -1, -1), stride=2, groups=C)


def activation1d_torch_grads(x, gy, alpha, beta, logscale: bool, taps_up, taps_down):
    """Autograd of the op sequence above: returns (gx, galpha, gbeta|None)."""
    x = x.detach().clone().requires_grad_(True)
    alpha = alpha.detach().clone().requires_grad_(True)
    beta_ = None if beta is None else beta.detach().clone().requires_grad_(True)
    y = activation1d_torch(x, alpha, beta_, logscale, taps_up, taps_down)
    y.backward(gy)
    return x.grad, alpha.grad, (None if beta_ is None else beta_.grad)


# --------------------------------------------------------------------------------------
# log-mel spectrogram: the reference's torch-op chains, restated
# --------------------------------------------------------------------------------------
def mel_spectrogram_torch(y, mel_basis, n_fft: int, hop_size: int, win_size: int):
    """meldataset.py:95-118 for a [B, T] waveform (reflect pad, hann, center=False); mel_basis [n_mels, n_fft/2+1]
    is what the reference caches at :89-92."""
    pad = (n_fft - hop_size) // 2
    y = F.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=torch.hann_window(win_size).to(y.device),
                      center=False, pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    spec = torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1) + 1e-9)
    return torch.log(torch.clamp(torch.matmul(mel_basis, spec), min=1e-5))


def msmsl_logmels_torch(wav, mel_basis, window_length: int, clamp_eps: float = 1e-5):
    """loss.py:110-167 + :195-197 for one scale (match_stride=False): [B, C, T] -> log10 mels [B, C, n_mels, frames]."""
    B, C, T = wav.shape
    window = torch.hann_window(window_length, periodic=True, dtype=torch.float64).float().to(wav.device)
    stft = torch.stft(wav.reshape(-1, T), n_fft=window_length, hop_length=window_length // 4, window=window,
                      return_complex=True, center=True)
    mag = torch.abs(stft).reshape(B, C, stft.shape[1], stft.shape[2])
    mels = (mag.transpose(2, -1) @ mel_basis.T).transpose(-1, 2)
    return torch.log(mels.clamp(min=clamp_eps)) / torch.log(torch.tensor(10.0))
