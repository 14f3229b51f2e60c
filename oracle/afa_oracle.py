"""CPU oracle for the anti-aliased activation path (TEST INFRASTRUCTURE ONLY).

This file is a plain numpy restatement of the reference's Activation1d path:

    UpSample1d (2x, 12-tap Kaiser-sinc, replicate pad)     BigVGAN/alias_free_activation/resample.py:10-38
    Snake / SnakeBeta                                      BigVGAN/activations.py:51-62, 113-126
    DownSample1d -> LowPassFilter1d (stride 2)             BigVGAN/alias_free_activation/resample.py:41-58,
                                                           BigVGAN/alias_free_activation/filter.py:65-101
    Activation1d.forward = down(act(up(x)))                BigVGAN/alias_free_activation/act.py:25-30

The arithmetic itself lives in PyTorch (third-party, reference pin `torch>=1.13.0`,
requirements.txt:2; this image has torch 2.11.0): F.pad(replicate), F.conv_transpose1d,
F.conv1d, torch.sin/pow/exp, torch.kaiser_window, torch.sinc.  Their published definitions are
restated here with numpy in float64 (or any dtype handed in).

Pinning: the reference ships NO tests and NO golden vectors for this path, so
parity is pinned on outputs of the reference itself run in the build container:
tests/golden/make_golden.py imports the unmodified reference modules, runs them (forward and
autograd backward) and commits the vectors under tests/golden/*.npz; tests/test_oracle.py checks
this file against every one of them.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product path (the CUDA library) never does.
"""
from __future__ import annotations

import math

import numpy as np

NO_DIV_BY_ZERO = 0.000000001  # activations.py:49, :111


# --------------------------------------------------------------------------------------
# filter design                                                         filter.py:9-62
# --------------------------------------------------------------------------------------
def _bessel_i0(x: np.ndarray) -> np.ndarray:
    """Modified Bessel function I0 by its power series (what torch.kaiser_window evaluates)."""
    x = np.asarray(x, dtype=np.float64)
    q = (x / 2.0) ** 2
    term = np.ones_like(x)
    acc = np.ones_like(x)
    for k in range(1, 64):
        term = term * q / (k * k)
        acc = acc + term
    return acc


def kaiser_window(n: int, beta: float) -> np.ndarray:
    """torch.kaiser_window(n, periodic=False, beta): I0(beta*sqrt(1-((2i/(n-1))-1)^2))/I0(beta)."""
    if n == 1:
        return np.ones(1)
    i = np.arange(n, dtype=np.float64)
    r = 2.0 * i / (n - 1) - 1.0
    return _bessel_i0(beta * np.sqrt(np.maximum(0.0, 1.0 - r * r))) / _bessel_i0(np.float64(beta))


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> np.ndarray:
    """filter.py:30-62 -> float64 taps [kernel_size], normalised to sum 1."""
    even = kernel_size % 2 == 0
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95          # filter.py:37-38
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = kaiser_window(kernel_size, beta)
    if even:
        time = np.arange(-half_size, half_size) + 0.5                # filter.py:48-49
    else:
        time = np.arange(kernel_size) - half_size
    if cutoff == 0:
        return np.zeros(kernel_size)
    filt = 2 * cutoff * window * np.sinc(2 * cutoff * time)          # np.sinc == torch.sinc
    return filt / filt.sum()                                         # filter.py:59


def default_taps(ratio: int = 2, kernel_size: int = 12) -> np.ndarray:
    """Taps both UpSample1d and DownSample1d build (resample.py:23-25, :48-53)."""
    return kaiser_sinc_filter1d(0.5 / ratio, 0.6 / ratio, kernel_size)


# --------------------------------------------------------------------------------------
# the three stages, stated the way the torch ops are defined
# --------------------------------------------------------------------------------------
def _pad_replicate(x: np.ndarray, left: int, right: int) -> np.ndarray:
    return np.concatenate(
        [np.repeat(x[..., :1], left, axis=-1), x, np.repeat(x[..., -1:], right, axis=-1)], axis=-1
    )


def upsample2x(x: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """resample.py:29-38 with ratio=2, kernel 12: pad 5/5, depthwise conv_transpose1d stride 2, x2, crop 15/15."""
    K = taps.shape[0]
    ratio, stride = 2, 2
    pad = K // ratio - 1
    pad_left = pad * stride + (K - stride) // 2
    pad_right = pad * stride + (K - stride + 1) // 2
    xp = _pad_replicate(x, pad, pad)
    Tp = xp.shape[-1]
    out = np.zeros(x.shape[:-1] + ((Tp - 1) * stride + K,), dtype=x.dtype)
    for k in range(K):                                   # conv_transpose1d: out[i*stride + k] += in[i]*w[k]
        out[..., k : k + stride * Tp : stride] += xp * taps[k]
    out = ratio * out
    return out[..., pad_left : out.shape[-1] - pad_right]


def snake(u: np.ndarray, alpha_eff: np.ndarray, beta_eff: np.ndarray) -> np.ndarray:
    """activations.py:124 (SnakeBeta) / :60 (Snake, beta_eff == alpha_eff). alpha/beta broadcast over [C,1]."""
    return u + (1.0 / (beta_eff + NO_DIV_BY_ZERO)) * np.sin(u * alpha_eff) ** 2


def downsample2x(s: np.ndarray, taps: np.ndarray) -> np.ndarray:
    """filter.py:94-101 with stride 2, kernel 12: pad 5/6 replicate, depthwise conv1d (cross-correlation)."""
    K = taps.shape[0]
    even = K % 2 == 0
    pad_left = K // 2 - int(even)
    pad_right = K // 2
    sp = _pad_replicate(s, pad_left, pad_right)
    n_out = (sp.shape[-1] - K) // 2 + 1
    y = np.zeros(s.shape[:-1] + (n_out,), dtype=s.dtype)
    for k in range(K):
        y += taps[k] * sp[..., k : k + 2 * n_out : 2]
    return y


def effective_params(alpha, beta, logscale: bool, dtype=np.float64):
    """activations.py:119-123 (exp if alpha_logscale). beta=None means Snake (beta := alpha, :57-60)."""
    a = np.asarray(alpha, dtype=dtype)
    b = a if beta is None else np.asarray(beta, dtype=dtype)
    if logscale:
        a, b = np.exp(a), np.exp(b)
    return a, b


def activation1d_forward(x, alpha, beta=None, logscale=True, taps_up=None, taps_down=None):
    """act.py:25-30. x [B,C,T]; alpha/beta raw parameters [C]; returns float64 [B,C,T]."""
    x = np.asarray(x, dtype=np.float64)
    taps_up = default_taps() if taps_up is None else np.asarray(taps_up, dtype=np.float64).reshape(-1)
    taps_down = default_taps() if taps_down is None else np.asarray(taps_down, dtype=np.float64).reshape(-1)
    a, b = effective_params(alpha, beta, logscale)
    u = upsample2x(x, taps_up)
    s = snake(u, a[None, :, None], b[None, :, None])
    return downsample2x(s, taps_down)


# --------------------------------------------------------------------------------------
# backward: exact adjoints of the three stages (what autograd replays at train_binaural_mel.py:787)
# --------------------------------------------------------------------------------------
def _pad_replicate_adjoint(g: np.ndarray, left: int, right: int) -> np.ndarray:
    T = g.shape[-1] - left - right
    out = g[..., left : left + T].copy()
    out[..., 0] += g[..., :left].sum(axis=-1)
    out[..., -1] += g[..., left + T :].sum(axis=-1)
    return out


def activation1d_backward(x, gy, alpha, beta=None, logscale=True, taps_up=None, taps_down=None):
    """Gradients of sum(y*gy) w.r.t. x and the RAW parameters. Returns (gx, galpha, gbeta|None)."""
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    taps_up = default_taps() if taps_up is None else np.asarray(taps_up, dtype=np.float64).reshape(-1)
    taps_down = default_taps() if taps_down is None else np.asarray(taps_down, dtype=np.float64).reshape(-1)
    K = 12
    a, b = effective_params(alpha, beta, logscale)
    a3, b3 = a[None, :, None], b[None, :, None]
    u = upsample2x(x, taps_up)
    T = x.shape[-1]

    # adjoint of downsample2x: scatter taps, then fold the replicate pad (5 left, 6 right)
    gsp = np.zeros(u.shape[:-1] + (2 * T + 11,), dtype=np.float64)
    for k in range(K):
        gsp[..., k : k + 2 * T : 2] += taps_down[k] * gy
    gs = _pad_replicate_adjoint(gsp, 5, 6)

    ib = 1.0 / (b3 + NO_DIV_BY_ZERO)
    sin_au = np.sin(u * a3)
    gu = gs * (1.0 + ib * a3 * np.sin(2.0 * u * a3))
    ga_eff = (gs * ib * u * np.sin(2.0 * u * a3)).sum(axis=(0, 2))
    gb_eff = (-gs * sin_au**2 * ib**2).sum(axis=(0, 2))

    # adjoint of upsample2x: un-crop (15/15), x2, gather conv_transpose taps, fold the pad (5/5)
    Tp = T + 10
    gfull = np.zeros(u.shape[:-1] + ((Tp - 1) * 2 + K,), dtype=np.float64)
    gfull[..., 15 : 15 + 2 * T] = 2.0 * gu
    gxp = np.zeros(x.shape[:-1] + (Tp,), dtype=np.float64)
    for k in range(K):
        gxp += taps_up[k] * gfull[..., k : k + 2 * Tp : 2]
    gx = _pad_replicate_adjoint(gxp, 5, 5)

    if logscale:                                   # d exp(p)/dp = exp(p)        activations.py:121-123
        ga_raw, gb_raw = ga_eff * a, gb_eff * b
    else:
        ga_raw, gb_raw = ga_eff, gb_eff
    if beta is None:                               # Snake: beta aliases alpha   activations.py:57-60
        return gx, ga_raw + gb_raw, None
    return gx, ga_raw, gb_raw


def param_grad_mass(x, gy, alpha, beta=None, logscale=True, taps_up=None, taps_down=None):
    """Sum of |terms| of the two parameter-gradient reductions (per channel, chain rule applied).

    The reductions sum signed terms that can cancel almost completely (e.g. one short row), so an
    fp32 reduction's error scales with this mass, not with the result; tests use it as a second
    normaliser next to max|g|.
    """
    x = np.asarray(x, dtype=np.float64)
    gy = np.asarray(gy, dtype=np.float64)
    taps_up = default_taps() if taps_up is None else np.asarray(taps_up, dtype=np.float64).reshape(-1)
    taps_down = default_taps() if taps_down is None else np.asarray(taps_down, dtype=np.float64).reshape(-1)
    a, b = effective_params(alpha, beta, logscale)
    a3, b3 = a[None, :, None], b[None, :, None]
    u = upsample2x(x, taps_up)
    T = x.shape[-1]
    gsp = np.zeros(u.shape[:-1] + (2 * T + 11,), dtype=np.float64)
    for k in range(12):
        gsp[..., k : k + 2 * T : 2] += taps_down[k] * gy
    gs = _pad_replicate_adjoint(gsp, 5, 6)
    ib = 1.0 / (b3 + NO_DIV_BY_ZERO)
    ma = np.abs(gs * ib * u * np.sin(2.0 * u * a3)).sum(axis=(0, 2))
    mb = np.abs(gs * np.sin(u * a3) ** 2 * ib**2).sum(axis=(0, 2))
    if logscale:
        ma, mb = ma * a, mb * b
    if beta is None:
        return ma + mb, None
    return ma, mb


def max_normalised_error(y, y_ref) -> float:
    """E = max|y - y_ref| / max|y_ref|  (pointwise relative error is ill-posed at zero crossings)."""
    y = np.asarray(y, dtype=np.float64)
    y_ref = np.asarray(y_ref, dtype=np.float64)
    denom = float(np.max(np.abs(y_ref))) if y_ref.size else 0.0
    if y.size == 0:
        return 0.0
    return float(np.max(np.abs(y - y_ref))) / (denom if denom > 0 else 1.0)
