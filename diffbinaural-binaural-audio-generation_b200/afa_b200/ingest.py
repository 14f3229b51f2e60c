"""Data formats either side of the vocoder: mel `.npy` in, zero-frame handling, int16 WAV out.

Reference (BigVGAN/inference_e2e.py, paths relative to the reference root):
    np.load(left/right mel)  float32 [num_mels, T_mel]                                   :140-141
    detect_and_exclude_zero_frames  (frames whose |.|-sum <= 1e-10 are dropped)           :38-74
    generator(x_left), generator(x_right)                                                :174-175
    reconstruct_audio_with_silence  (generated hops go back to their frames, rest = 0)   :77-111
    np.stack([L, R]) * 32767 -> astype(int16) -> .T -> scipy.io.wavfile.write            :189-205

Here the frame bookkeeping is index work on the host (tiny: T_mel integers per channel); the restoration itself is not
a separate pass: the fused tail kernel scatters every generated hop straight to its frame of the zero-initialised int16
PCM (`frame_map` of `afa_tail_fwd_cl`).  Left and right may keep different numbers of frames; they share a launch when
the numbers agree and run as two batch-1 launches otherwise (padding would change the samples near the end of the
shorter one, which the reference does not do).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch

MAX_WAV_VALUE = 32767.0          # meldataset.py:20


def load_mel_npy(path: str) -> np.ndarray:
    """float32 [num_mels, T_mel] as test_realBinaural.py:266-278 saves it."""
    mel = np.load(path)
    if mel.ndim == 3 and mel.shape[0] == 1:
        mel = mel[0]
    if mel.ndim != 2:
        raise ValueError(f"{path}: expected a [num_mels, T_mel] array, got {mel.shape}")
    return np.ascontiguousarray(mel, dtype=np.float32)


def detect_zero_frames(mel: np.ndarray, zero_threshold: float = 1e-10) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """inference_e2e.py:38-74 -> (filtered_mel [num_mels, n_kept], zero_mask [T_mel] bool, nonzero_indices [n_kept])."""
    frame_sums = np.sum(np.abs(mel), axis=0)
    zero_mask = frame_sums <= zero_threshold
    if not np.any(zero_mask):
        return mel, zero_mask, np.arange(mel.shape[1])
    nonzero_indices = np.where(~zero_mask)[0]
    return mel[:, nonzero_indices], zero_mask, nonzero_indices


def restore_silence_host(filtered_audio: np.ndarray, nonzero_indices: np.ndarray, hop_size: int, original_length: int) -> np.ndarray:
    """inference_e2e.py:77-111 as one vectorised scatter (host-side twin of the kernel's frame_map; used by tests)."""
    out = np.zeros(original_length, dtype=filtered_audio.dtype)
    n = min(len(nonzero_indices), len(filtered_audio) // hop_size)
    idx = np.asarray(nonzero_indices[:n], dtype=np.int64)
    keep = (idx + 1) * hop_size <= original_length
    frames = out[: (original_length // hop_size) * hop_size].reshape(-1, hop_size)
    frames[idx[keep]] = filtered_audio[: n * hop_size].reshape(n, hop_size)[keep]
    return out


def vocode_binaural(engine, mel_left: np.ndarray, mel_right: np.ndarray, interpolate_zero_frames: bool = True,
                    device: Optional[torch.device] = None) -> torch.Tensor:
    """One clip of inference_e2e.py's loop body (:140-201) on the channels-last engine: two [num_mels, T_mel] mels ->
    interleaved int16 stereo PCM [T_mel * hop, 2] on the device (what `scipy.io.wavfile.write` takes after `.cpu().numpy()`)."""
    device = device or engine.device
    if mel_left.shape[1] != mel_right.shape[1]:
        raise ValueError("left and right mels of a clip have the same number of frames (inference_e2e.py:140-171)")
    t_mel = mel_left.shape[1]
    hop = 1
    for u in engine.h["upsample_rates"]:
        hop *= u
    t_out = t_mel * hop
    if interpolate_zero_frames:
        fl, _, il = detect_zero_frames(mel_left)
        fr, _, ir = detect_zero_frames(mel_right)
    else:
        fl, il, fr, ir = mel_left, np.arange(t_mel), mel_right, np.arange(t_mel)
    pcm = torch.zeros(1, t_out, 2, dtype=torch.int16, device=device)
    if fl.shape[1] == fr.shape[1]:
        if fl.shape[1] == 0:
            return pcm[0]
        mel = torch.from_numpy(np.stack([fl, fr])).to(device, non_blocking=True)
        fmap = torch.from_numpy(np.stack([il, ir]).astype(np.int32)).to(device, non_blocking=True)
        _, out = engine(mel, want_pcm=True, pcm_interleave=2, frame_map=fmap, t_out=t_out)
        return out[0]
    # different numbers of kept frames: one launch per channel, each scattering into its half of the interleaved PCM
    for ch, (f, idx) in enumerate(((fl, il), (fr, ir))):
        if f.shape[1] == 0:
            continue
        mel = torch.from_numpy(f[None]).to(device, non_blocking=True)
        fmap = torch.from_numpy(idx.astype(np.int32)[None]).to(device, non_blocking=True)
        _, mono = engine(mel, want_pcm=True, pcm_interleave=1, frame_map=fmap, t_out=t_out)
        pcm[0, :, ch] = mono[0, :, 0]
    return pcm[0]


def write_wav(path: str, sampling_rate: int, pcm_stereo) -> None:
    """inference_e2e.py:205: scipy.io.wavfile.write(output_file, h.sampling_rate, stereo_audio [T, 2] int16)."""
    from scipy.io.wavfile import write

    if isinstance(pcm_stereo, torch.Tensor):
        pcm_stereo = pcm_stereo.cpu().numpy()
    write(path, sampling_rate, np.ascontiguousarray(pcm_stereo, dtype=np.int16))
