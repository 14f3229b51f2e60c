"""CPU oracle for the log-mel spectrogram row -- TEST INFRASTRUCTURE ONLY.

Plain numpy float64 restatement, op by op, of

    mel_spectrogram                                   BigVGAN/meldataset.py:51-123
      F.pad(reflect | constant)                       :95-101
      torch.stft(hann, center=False, onesided)        :103-114   (third-party PyTorch, pin torch>=1.13.0)
      sqrt(re^2 + im^2 + 1e-9)                        :115
      mel_basis @ spec, log(clamp(., 1e-5))           :117-118, :31-32
    MultiScaleMelSpectrogramLoss.mel_spectrogram      BigVGAN/loss.py:110-167  (center=True STFT, |.|, basis)
    MultiScaleMelSpectrogramLoss.forward              BigVGAN/loss.py:169-211  (log10 of the clamped mels, L1 per scale)

and of the filterbank both take from a THIRD-PARTY dependency that is absent from /root/reference and from this
image: `librosa.filters.mel` (reference pin: `librosa>=0.8.1`, requirements.txt:9; defaults htk=False,
norm='slaney').  `slaney_mel_filterbank` restates librosa's published algorithm (Slaney's Auditory Toolbox mel scale:
linear below 1 kHz, log above; triangles in Hz; area normalisation 2 / (f[m+2] - f[m])); tests/test_oracle_mel.py
pins it on an independent implementation of the same published algorithm that IS installed
(`transformers.audio_utils.mel_filter_bank(norm='slaney', mel_scale='slaney')`, the one Whisper's feature extractor
uses in place of librosa) through tests/golden/mel_golden.npz.  Against librosa's own array the filterbank is therefore
PARITY UNPINNED (librosa cannot be installed here); a caller who has librosa passes its basis in (`mel_basis=`).  The STFT / magnitude / log part is pinned on outputs
of the reference's own `mel_spectrogram` and `MultiScaleMelSpectrogramLoss`, run unmodified by
tests/golden/make_golden_mel.py with that filterbank standing in for librosa's.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np


# --------------------------------------------------------------------------------------
# librosa.filters.mel (htk=False, norm='slaney'), restated
# --------------------------------------------------------------------------------------
_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = math.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    lin = f / _F_SP
    log = _MIN_LOG_MEL + np.log(np.maximum(f, _MIN_LOG_HZ) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(f >= _MIN_LOG_HZ, log, lin)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    lin = _F_SP * m
    log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (np.maximum(m, _MIN_LOG_MEL) - _MIN_LOG_MEL))
    return np.where(m >= _MIN_LOG_MEL, log, lin)


def slaney_mel_filterbank(sr: float, n_fft: int, n_mels: int, fmin: float = 0.0, fmax=None) -> np.ndarray:
    """float32 [n_mels, 1 + n_fft // 2], as librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=) returns it."""
    if fmax is None:
        fmax = float(sr) / 2.0
    fftfreqs = np.linspace(0.0, float(sr) / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    return (weights * enorm[:, None]).astype(np.float32)


# --------------------------------------------------------------------------------------
# windows
# --------------------------------------------------------------------------------------
def hann_periodic(n: int) -> np.ndarray:
    """torch.hann_window(n) (periodic=True, meldataset.py:93) == scipy.signal.get_window('hann', n) (loss.py:100)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n, dtype=np.float64) / n)


# --------------------------------------------------------------------------------------
# STFT magnitudes -> mel
# --------------------------------------------------------------------------------------
def pad_rows(y: np.ndarray, pad: int, mode: str) -> np.ndarray:
    if pad == 0:
        return y
    return np.pad(y, ((0, 0), (pad, pad)), mode="reflect" if mode == "reflect" else "constant")


def stft_mag(y: np.ndarray, n_fft: int, hop: int, window: np.ndarray, pad: int, pad_mode: str, mag_eps: float) -> np.ndarray:
    """[rows, T] -> [rows, n_fft // 2 + 1, n_frames]: sqrt(re^2 + im^2 + mag_eps) of the framed, windowed rfft."""
    y = pad_rows(np.asarray(y, dtype=np.float64), pad, pad_mode)
    rows, Tp = y.shape
    n_frames = 1 + (Tp - n_fft) // hop if Tp >= n_fft else 0
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    frames = y[:, idx] * window[None, None, :]
    spec = np.fft.rfft(frames, axis=-1)
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2 + mag_eps)
    return np.transpose(mag, (0, 2, 1))


def mel_spectrogram(y: np.ndarray, n_fft: int, num_mels: int, sampling_rate: int, hop_size: int, win_size: int,
                    fmin: float, fmax=None, mel_basis: np.ndarray | None = None) -> np.ndarray:
    """meldataset.py:51-123 (center=False).  y: [B, T] (reflect pad) or [T] (zero pad, returns [1, n_mels, frames])."""
    assert win_size == n_fft, "the reference's configs use win_size == n_fft"
    if mel_basis is None:
        mel_basis = slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax)
    pad = (n_fft - hop_size) // 2
    if y.ndim == 1:
        y2, mode = y[None, :], "constant"
    else:
        y2, mode = y, "reflect"
    mag = stft_mag(y2, n_fft, hop_size, hann_periodic(win_size), pad, mode, 1e-9)
    mel = np.einsum("mk,rkf->rmf", mel_basis.astype(np.float64), mag)
    return np.log(np.maximum(mel, 1e-5))


def msmsl_mels(wav: np.ndarray, sampling_rate: int, n_mels: int, window_length: int, fmin: float = 0.0, fmax=None,
               mel_basis: np.ndarray | None = None) -> np.ndarray:
    """loss.py:110-167 with match_stride=False: [B, C, T] -> [B, C, n_mels, frames] (raw mel magnitudes)."""
    B, C, T = wav.shape
    hop = window_length // 4
    if mel_basis is None:
        mel_basis = slaney_mel_filterbank(sampling_rate, window_length, n_mels, fmin, fmax)
    mag = stft_mag(wav.reshape(-1, T), window_length, hop, hann_periodic(window_length), window_length // 2, "reflect", 0.0)
    mel = np.einsum("mk,rkf->rmf", mel_basis.astype(np.float64), mag)
    return mel.reshape(B, C, n_mels, -1)


def logmel_backward(y: np.ndarray, gout: np.ndarray, mel_basis: np.ndarray, n_fft: int, hop: int, pad: int, pad_mode: str,
                    mag_eps: float, clamp_eps: float, log_scale: float, raw: bool = False) -> np.ndarray:
    """d <out, gout> / d y for out = log(max(basis @ sqrt(|STFT(y)|^2 + mag_eps), clamp_eps)) * log_scale (or the raw mels):
    the derivative autograd takes through meldataset.py:95-118 / loss.py:131-167, 195-197, stated op by op in reverse.
    torch.clamp passes the gradient where its input >= min; torch.abs has gradient 0 at 0.  y [rows, T], gout
    [rows, n_mels, n_frames] -> [rows, T] float64."""
    y = np.asarray(y, dtype=np.float64)
    rows, T = y.shape
    window = hann_periodic(n_fft)
    yp = pad_rows(y, pad, pad_mode)
    Tp = yp.shape[1]
    n_frames = 1 + (Tp - n_fft) // hop
    idx = np.arange(n_frames)[:, None] * hop + np.arange(n_fft)[None, :]
    spec = np.fft.rfft(yp[:, idx] * window[None, None, :], axis=-1)                  # [rows, frames, bins]
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2 + mag_eps)
    basis = mel_basis.astype(np.float64)
    g = np.transpose(np.asarray(gout, dtype=np.float64), (0, 2, 1))                  # [rows, frames, mels]
    if not raw:
        mel = mag @ basis.T
        g = np.where(mel >= clamp_eps, g * log_scale / np.maximum(mel, 1e-300), 0.0)
    gmag = g @ basis                                                                 # [rows, frames, bins]
    scale = np.where(mag > 0, gmag / np.where(mag > 0, mag, 1.0), 0.0)
    G = scale * spec                                                                 # d / d re + i d / d im
    # adjoint of rfft: g[n] = Re sum_{k=0..N/2} G_k e^{+2 pi i k n / N} = N * irfft(H), H_0 = Re G_0, H_{N/2} = Re G_{N/2}, else G_k / 2
    H = G * 0.5
    H[..., 0] = G[..., 0].real
    H[..., -1] = G[..., -1].real
    gframes = np.fft.irfft(H, n=n_fft, axis=-1) * n_fft * window[None, None, :]
    gp = np.zeros((rows, Tp))
    for r in range(rows):
        np.add.at(gp[r], idx, gframes[r])
    if pad == 0:
        return gp
    gy = gp[:, pad:pad + T].copy()
    if pad_mode == "reflect":                                                        # padded[pad - j] = y[j], padded[pad + T - 1 + j] = y[T - 1 - j]
        j = np.arange(1, pad + 1)
        gy[:, j] += gp[:, pad - j]
        gy[:, T - 1 - j] += gp[:, pad + T - 1 + j]
    return gy


def msmsl_loss_backward(x: np.ndarray, y: np.ndarray, sampling_rate: int, n_mels=None, window_lengths=None,
                        clamp_eps: float = 1e-5, log_weight: float = 1.0, mag_weight: float = 0.0) -> np.ndarray:
    """d msmsl_loss / d x ([B, C, T]): L1 mean's subgradient sign(lx - ly) / numel through every scale."""
    n_mels = MSMSL_N_MELS if n_mels is None else n_mels
    window_lengths = MSMSL_WINDOWS if window_lengths is None else window_lengths
    B, C, T = x.shape
    gx = np.zeros((B * C, T))
    for nm, w in zip(n_mels, window_lengths):
        basis = slaney_mel_filterbank(sampling_rate, w, nm, 0.0, None)
        lx = np.log(np.maximum(msmsl_mels(x, sampling_rate, nm, w), clamp_eps)) / math.log(10.0)
        ly = np.log(np.maximum(msmsl_mels(y, sampling_rate, nm, w), clamp_eps)) / math.log(10.0)
        gout = (log_weight + mag_weight) * np.sign(lx - ly) / lx.size
        gx += logmel_backward(x.reshape(B * C, T), gout.reshape(B * C, nm, -1), basis, w, w // 4, w // 2, "reflect", 0.0,
                              clamp_eps, 1.0 / math.log(10.0))
    return gx.reshape(B, C, T)


MSMSL_N_MELS = (5, 10, 20, 40, 80, 160, 320)          # loss.py:56
MSMSL_WINDOWS = (32, 64, 128, 256, 512, 1024, 2048)   # loss.py:57


def msmsl_loss(x: np.ndarray, y: np.ndarray, sampling_rate: int, n_mels=MSMSL_N_MELS, window_lengths=MSMSL_WINDOWS,
               clamp_eps: float = 1e-5, log_weight: float = 1.0, mag_weight: float = 0.0) -> float:
    """loss.py:169-211 with the defaults train_binaural_mel.py:458-460 uses (L1, pow = 1)."""
    total = 0.0
    for nm, w in zip(n_mels, window_lengths):
        lx = np.log(np.maximum(msmsl_mels(x, sampling_rate, nm, w), clamp_eps)) / math.log(10.0)
        ly = np.log(np.maximum(msmsl_mels(y, sampling_rate, nm, w), clamp_eps)) / math.log(10.0)
        l1 = float(np.mean(np.abs(lx - ly)))
        total += log_weight * l1 + mag_weight * l1     # both terms compare the LOG mels in the reference (:206-207)
    return total
