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


# ------------------------------------------------------------------------------------------------
# batched path: many clips per launch, zero-frame compaction on the device
# ------------------------------------------------------------------------------------------------
def compact_zero_frames(mel: torch.Tensor, zero_threshold: float = 1e-10):
    """detect_and_exclude_zero_frames (inference_e2e.py:38-74) for a whole batch in one launch (afa_compact_zero_frames).
    mel: float32 [rows, n_mels, T] on the device.  Returns (packed [rows, n_mels, T] with the kept frames left-packed,
    frame_map int32 [rows, T] (original frame of each kept frame, -1 behind them), n_kept int32 [rows])."""
    from . import _lib

    if not mel.is_cuda or mel.dtype != torch.float32 or mel.dim() != 3 or not mel.is_contiguous():
        raise RuntimeError("compact_zero_frames takes a contiguous float32 [rows, n_mels, T] CUDA tensor (there is no CPU fallback)")
    rows, n_mels, T = mel.shape
    packed = torch.empty_like(mel)
    frame_map = torch.empty(rows, T, dtype=torch.int32, device=mel.device)
    n_kept = torch.empty(rows, dtype=torch.int32, device=mel.device)
    with torch.cuda.device_of(mel):
        rc = _lib.load_library().afa_compact_zero_frames(mel.data_ptr(), packed.data_ptr(), frame_map.data_ptr(), n_kept.data_ptr(),
                                                         rows, n_mels, T, float(zero_threshold),
                                                         torch.cuda.current_stream(mel.device).cuda_stream)
    _lib.check(rc, "afa_compact_zero_frames")
    return packed, frame_map, n_kept


class BatchedVocoder:
    """inference_e2e.py:129-205 for BATCHES of clips: pinned mels -> one H2D -> one compaction launch -> generator ->
    interleaved int16 stereo PCM.  Rows (clip x channel) that keep every frame -- the normal case -- run through a CUDA-graph
    replay of the channels-last engine, `clips_per_batch` clips (2 x that many rows) at a time; rows that lost frames run
    eagerly at their own length, grouped by length (padding a shorter row would change its last samples, which the reference
    does not do), and are scattered back through their frame map by the tail kernel.  The only host synchronisation per batch
    is the read of `n_kept` (rows x 4 bytes) that decides the routing."""

    def __init__(self, engine, clips_per_batch: int, t_mel: int):
        from .engine import GraphedEngine

        self.engine, self.clips, self.t_mel = engine, clips_per_batch, t_mel
        self.hop = 1
        for u in engine.h["upsample_rates"]:
            self.hop *= u
        self.graphed = GraphedEngine(engine, 2 * clips_per_batch, t_mel, want_pcm=True, pcm_interleave=2)

    def __call__(self, mel_batch: torch.Tensor) -> torch.Tensor:
        """mel_batch: float32 [clips, 2, n_mels, t_mel] on the device -> int16 PCM [clips, t_mel * hop, 2] (a fresh tensor)."""
        clips, two, n_mels, T = mel_batch.shape
        if (clips, two, T) != (self.clips, 2, self.t_mel):
            raise ValueError(f"expected [{self.clips}, 2, n_mels, {self.t_mel}] mels, got {tuple(mel_batch.shape)}")
        rows = mel_batch.reshape(2 * clips, n_mels, T)
        packed, fmap, n_kept = compact_zero_frames(rows)
        kept = n_kept.cpu().tolist()                               # the one synchronising read: routes the rows
        t_out = T * self.hop
        if all(k == T for k in kept):
            _, pcm = self.graphed(rows)                            # nothing dropped: the batch as it came
            return pcm.clone()
        pcm = torch.zeros(clips, t_out, 2, dtype=torch.int16, device=rows.device)
        by_len = {}
        for r, k in enumerate(kept):
            by_len.setdefault(k, []).append(r)
        flat = pcm.view(clips, t_out, 2)
        # rare lengths run eagerly: no cuDNN autotuning for a shape that will not come back (it costs seconds per new length)
        # (the context manager resets EVERY cuDNN flag to its own default -- allow_tf32=True among them -- unless told otherwise:
        # the caller's precision and determinism settings are passed through)
        with torch.backends.cudnn.flags(enabled=torch.backends.cudnn.enabled, benchmark=False,
                                        deterministic=torch.backends.cudnn.deterministic,
                                        allow_tf32=torch.backends.cudnn.allow_tf32):
            for k, rs in sorted(by_len.items()):
                if k == 0:
                    continue                                       # a silent channel stays silence
                idx = torch.tensor(rs, device=rows.device)
                _, mono = self.engine(packed.index_select(0, idx)[:, :, :k].contiguous(), want_pcm=True, pcm_interleave=1,
                                      frame_map=fmap.index_select(0, idx)[:, :k].contiguous(), t_out=t_out)
                for j, r in enumerate(rs):
                    flat[r // 2, :, r % 2] = mono[j, :, 0]
        return pcm
