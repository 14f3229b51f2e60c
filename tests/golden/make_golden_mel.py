#!/usr/bin/env python
"""Generate tests/golden/mel_golden.npz by running the UNMODIFIED reference `mel_spectrogram`
(BigVGAN/meldataset.py:51-123) and `MultiScaleMelSpectrogramLoss` (BigVGAN/loss.py:23-211) in the build container.

    python tests/golden/make_golden_mel.py        # needs /root/reference (read-only)

Both import `librosa.filters.mel`, a third-party dependency absent from this image (reference pin librosa>=0.8.1,
requirements.txt:9).  It is stood in for by `transformers.audio_utils.mel_filter_bank(norm='slaney',
mel_scale='slaney')` -- an independent, installed implementation of the same published (Slaney) filterbank, the one
Whisper's feature extractor uses instead of librosa -- and the filterbanks it returned are stored in the fixture, so
the oracle's own restatement (oracle/mel_oracle.py) is checked against them too.  Everything else (pad, torch.stft,
magnitude, matmul, log, the seven-scale L1) is the reference's code, executed as it lies.  `loss.py` also imports
`scipy.signal` (installed).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get("AFA_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
BIG = os.path.join(REF, "BigVGAN")


def librosa_mel_standin(sr, n_fft, n_mels, fmin=0.0, fmax=None, **kw):
    from transformers.audio_utils import mel_filter_bank

    if fmax is None:
        fmax = sr / 2.0
    fb = mel_filter_bank(num_frequency_bins=1 + n_fft // 2, num_mel_filters=n_mels, min_frequency=float(fmin),
                         max_frequency=float(fmax), sampling_rate=sr, norm="slaney", mel_scale="slaney")
    return np.ascontiguousarray(fb.T).astype(np.float32)


def load_reference():
    import transformers.audio_utils  # noqa: F401  (before the stub: transformers probes for a real librosa at import)

    for name in ("librosa", "librosa.filters", "librosa.util"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["librosa.filters"].mel = librosa_mel_standin
    sys.modules["librosa"].filters = sys.modules["librosa.filters"]
    sys.path.insert(0, BIG)

    def load(name, file):
        spec = importlib.util.spec_from_file_location(name, os.path.join(BIG, file))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    return load("meldataset", "meldataset.py"), load("loss", "loss.py")


def main():
    import torch

    meldataset, loss = load_reference()
    torch.manual_seed(1234)
    out = {}
    sr, n_fft, num_mels, hop, win, fmin, fmax = 22050, 1024, 80, 256, 1024, 0, None   # bigvgan_binaural_22khz_80band_256x.json
    # broadband (every bin carries energy, so the log is well conditioned) + a tone, inside [-1, 1]
    T = 8192
    t = torch.arange(T, dtype=torch.float32) / sr
    y = 0.25 * torch.randn(3, T) + 0.3 * torch.sin(2 * np.pi * 440.0 * t)[None, :] * torch.tensor([[1.0], [0.0], [0.5]])
    y = y.clamp(-1, 1)
    out["y"] = y.numpy()
    out["mel_2d"] = meldataset.mel_spectrogram(y, n_fft, num_mels, sr, hop, win, fmin, fmax).numpy()      # reflect pad
    out["mel_1d"] = meldataset.mel_spectrogram(y[0], n_fft, num_mels, sr, hop, win, fmin, fmax).numpy()   # zero pad branch
    y_short = y[:, :1500].contiguous()                                                                     # ragged: T % hop != 0
    out["mel_short"] = meldataset.mel_spectrogram(y_short, n_fft, num_mels, sr, hop, win, fmin, fmax).numpy()
    out["basis_80_1024"] = librosa_mel_standin(sr, n_fft, num_mels, fmin, fmax)
    out["basis_80_1024_fmax8000"] = librosa_mel_standin(sr, n_fft, num_mels, 0, 8000)

    msl = loss.MultiScaleMelSpectrogramLoss(sampling_rate=sr)
    x = y[:2, :4096].reshape(2, 1, 4096).contiguous()
    xh = (x + 0.05 * torch.randn(x.shape)).clamp(-1, 1)
    out["msl_x"] = xh.numpy()
    out["msl_y"] = x.numpy()
    out["msl_loss"] = np.float64(msl(xh, x).item())
    for nm, w in zip(msl.n_mels, [s.window_length for s in msl.stft_params]):
        out[f"msl_mels_{w}"] = msl.mel_spectrogram(xh, nm, 0, None, w, w // 4, False, "hann").numpy()
        out[f"basis_{nm}_{w}"] = librosa_mel_standin(sr, w, nm, 0, None)
    # backward (autograd through the reference's own op chains, float32): d/dy of <mel_spectrogram(y), G> for the reflect and
    # the zero-pad branch, and d/dx of the seven-scale loss (what loss_mel.backward() propagates, train_binaural_mel.py:759-787)
    G = torch.randn(3, num_mels, 32)
    yg = y.clone().requires_grad_(True)
    (meldataset.mel_spectrogram(yg, n_fft, num_mels, sr, hop, win, fmin, fmax) * G).sum().backward()
    out["bwd_G"] = G.numpy()
    out["bwd_gy_2d"] = yg.grad.numpy()
    y1 = y[0].clone().requires_grad_(True)
    (meldataset.mel_spectrogram(y1, n_fft, num_mels, sr, hop, win, fmin, fmax) * G[0]).sum().backward()
    out["bwd_gy_1d"] = y1.grad.numpy()
    xg = xh.clone().requires_grad_(True)
    msl(xg, x).backward()
    out["msl_gx"] = xg.grad.numpy()
    path = os.path.join(HERE, "mel_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
