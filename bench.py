#!/usr/bin/env python
"""bench.py -- the hot path's headline metric: Activation1d HBM GB/s on the BigVGAN AMP shapes.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (the reference's torch CPU path, host cores)
    python bench.py --table                                  (per-shape table, fp32/bf16, fwd/bwd; extra)

A STEP is one pass of the hot path over one batch of synthetic input: the 109 Activation1d forwards of
one `bigvgan_binaural_22khz_80band_256x` generator pass (reference: BigVGAN/bigvgan.py:361-387 with
configs/bigvgan_binaural_22khz_80band_256x.json: channels 768..24, T = T_mel x 4..256, 18 calls per
stage + activation_post) for `--clips` 10-second binaural clips per GPU (B = 2 x clips rows per channel,
T_mel = 861), fp32, through the C ABI (afa_activation1d_fwd).  `value` is algorithmic GB/s
(8 B per element: read x, write y) with inputs resident in HBM; every call's tensors are larger than
L2 (and rotate), so launches are L2-cold.  `e2e` is the same metric through the nn.Module mirror with
HOST buffers (pinned H2D of every call's input and D2H of every call's output inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")
for _p in (REPO, PKG_ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# (channels, T / T_mel, calls per generator pass)   bigvgan.py:302-328, :345; SURVEY.md section 8a
AMP_STAGES = [(768, 4, 18), (384, 16, 18), (192, 32, 18), (96, 64, 18), (48, 128, 18), (24, 256, 19)]
T_MEL_10S = 861
METRIC = "activation1d_hbm_gbps"
UNIT = "GB/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def host_cores() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown CPU"


def traffic_record(kernel: str, dtype: str, shape):
    """profiles/traffic.json: DRAM bytes of the SAME launch (kernel, dtype, shape) from a committed ncu capture, or None."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            tj = json.load(f)
        for rec in tj.get("launches", []):
            if rec["kernel"] == kernel and rec["dtype"] == dtype and list(rec["shape"]) == list(shape):
                return rec
    except Exception:
        pass
    return None


def stage_shapes(clips: int, t_mel: int):
    return [(2 * clips, c, mult * t_mel, calls) for (c, mult, calls) in AMP_STAGES]


def step_elements(clips: int, t_mel: int) -> int:
    return sum(b * c * t * calls for (b, c, t, calls) in stage_shapes(clips, t_mel))


# ----------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed region
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's torch CPU path (oracle/torch_path.py port)
# ----------------------------------------------------------------------------------------------
def cpu_reference_pass(t_mel: int, clips: int = 1, repeats: int = 1, threads: int | None = None):
    """One bounded sample: one Activation1d call per AMP stage shape (6 calls) for `clips` clip(s), fp32, torch CPU."""
    import torch

    from oracle import torch_path as TP

    # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    torch.set_num_threads(threads or host_cores())
    torch.manual_seed(1234)
    taps = TP.make_taps()
    work = []
    for (b, c, t, _calls) in stage_shapes(clips, t_mel):
        work.append((torch.randn(b, c, t), torch.randn(c) * 0.5, torch.randn(c) * 0.5))
    elems = sum(x.numel() for x, _, _ in work)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            for x, a, b_ in work:
                TP.activation1d_torch(x, a, b_, True, taps, taps)
            times.append(time.perf_counter() - t0)
    return elems, times


def cpu_generator_pass(t_mel: int, repeats: int = 1):
    """The second half of the metric on the host: the whole generator (bigvgan.py:361-387 restated by afa_b200/vocoder.py --
    torch.nn convolutions, the reference's checkpoint layout) with the reference's torch-op Activation1d (oracle/torch_path.py)
    on CPU, fp32, B = 2 (L, R), random init, all host threads, `t_mel` mel frames (bounded: a 10 s clip takes ~45 s)."""
    import torch
    import torch.nn as nn

    from afa_b200.modules import DownSample1d, UpSample1d
    from afa_b200.vocoder import BigVGANGenerator
    from oracle import torch_path as TP

    class TorchOpActivation1d(nn.Module):
        def __init__(self, activation):
            super().__init__()
            self.act, self.upsample, self.downsample = activation, UpSample1d(2, 12), DownSample1d(2, 12)

        def forward(self, x):
            beta = getattr(self.act, "beta", None)
            return TP.activation1d_torch(x, self.act.alpha, beta, bool(self.act.alpha_logscale), self.upsample.filter,
                                         self.downsample.lowpass.filter)

    torch.set_num_threads(host_cores())
    torch.manual_seed(1234)
    gen = BigVGANGenerator(activation_factory=TorchOpActivation1d).eval()
    mel = torch.rand(2, 80, t_mel) * 14.5 - 12.0
    times = []
    with torch.no_grad():
        gen(mel[:, :, : min(8, t_mel)])                      # warm the thread pool / oneDNN primitives
        for _ in range(repeats):
            t0 = time.perf_counter()
            gen(mel)
            times.append(time.perf_counter() - t0)
    audio_s = t_mel * 256 / 22050.0
    return {"audio_sec_per_sec": round(audio_s / min(times), 4), "unit": "binaural audio-s per wall-s", "t_mel": t_mel,
            "seconds_per_pass": round(min(times), 3), "cores": host_cores(), "cpu": cpu_model(), "kind": "port",
            "sample": (f"full generator (112 M parameters, random init), B=2 (L, R), T_mel={t_mel} ({audio_s:.2f} s of audio), fp32, torch CPU "
                       f"ops incl. the reference's Activation1d op chain, best of {repeats}; compare with vocoder.audio_sec_per_sec of the GPU arm")}


def run_reference(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    t_mel = args.t_mel
    sample = (f"one Activation1d call per AMP stage shape (6 calls, B=2 i.e. one binaural clip, T_mel={t_mel}, fp32) "
              f"per step, torch CPU ops, {cores} threads on {cpu_model()}; the GPU arm times all 109 calls for 8 clips -- both arms "
              f"are normalised to GB/s of the same per-element byte count (same_config: shapes equal per stage, batch differs)")
    elems, _ = cpu_reference_pass(t_mel, 1, repeats=max(1, args.warmup))
    elems, times = cpu_reference_pass(t_mel, 1, repeats=args.steps)
    total_s = sum(times)
    value = elems * 8 * len(times) / total_s / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * total_s / len(times), 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"bigvgan_binaural_22khz_80band_256x AMP Activation1d shapes, T_mel={t_mel}",
                   "arm": "reference torch CPU path (oracle/torch_path.py port of alias_free_activation/*.py)"},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    try:
        line["vocoder_cpu"] = cpu_generator_pass(64, repeats=1)
    except Exception as exc:  # noqa: BLE001
        line["vocoder_cpu"] = {"error": repr(exc)[:200]}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class Workload:
    """Device tensors + modules for one generator pass worth of Activation1d calls."""

    def __init__(self, dev, clips: int, t_mel: int, dtype, rotate: int = 2):
        import torch

        from afa_b200 import Activation1d
        from afa_b200.activations import SnakeBeta

        self.dev, self.dtype = dev, dtype
        self.stages = []
        g = torch.Generator(device="cpu").manual_seed(1234)
        for (b, c, t, calls) in stage_shapes(clips, t_mel):
            act = SnakeBeta(c, alpha_logscale=True)                       # configs/...256x.json:20-21
            with torch.no_grad():
                act.alpha.copy_(torch.randn(c, generator=g) * 0.5)
                act.beta.copy_(torch.randn(c, generator=g) * 0.5)
            mod = Activation1d(activation=act).to(dev)
            xs = [torch.randn(b, c, t, device=dev).to(dtype) for _ in range(rotate)]
            ys = [torch.empty_like(xs[0]) for _ in range(rotate)]
            self.stages.append({"shape": (b, c, t), "calls": calls, "mod": mod, "xs": xs, "ys": ys})
        self.elements = sum(s["shape"][0] * s["shape"][1] * s["shape"][2] * s["calls"] for s in self.stages)
        self.launches = sum(s["calls"] for s in self.stages)

    def step_device(self):
        """109 calls through the C ABI on resident tensors (the library launches on the current stream)."""
        from afa_b200 import functional as Fn

        for s in self.stages:
            m = s["mod"]
            tu, td = m._host_taps()
            a, b = m.act.alpha.detach(), m.act.beta.detach()
            for k in range(s["calls"]):
                Fn.activation1d_forward_raw(s["xs"][k % len(s["xs"])], a, b, tu, td, True, out=s["ys"][k % len(s["ys"])])


def run_e2e(wl: Workload, steps: int, warmup: int):
    """Same step through the nn.Module mirror with HOST buffers: H2D of every call's input, D2H of every output."""
    import torch

    dev = wl.dev
    host = []
    for s in wl.stages:
        b, c, t = s["shape"]
        hx = torch.randn(b, c, t).to(wl.dtype).pin_memory()
        hy = torch.empty(b, c, t, dtype=wl.dtype).pin_memory()
        host.append((hx, hy))
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    cur = torch.cuda.current_stream(dev)
    h2d = sum(hx.numel() * hx.element_size() * s["calls"] for (hx, _), s in zip(host, wl.stages))

    def one_step():
        for (hx, hy), s in zip(host, wl.stages):
            m = s["mod"]
            n = len(s["xs"])
            ev_in = [None] * n
            ev_done = [None] * n
            ev_out = [None] * n
            for k in range(s["calls"]):
                i = k % n
                with torch.cuda.stream(s_in):
                    if ev_done[i] is not None:
                        s_in.wait_event(ev_done[i])           # device input slot free again
                    s["xs"][i].copy_(hx, non_blocking=True)
                    ev_in[i] = torch.cuda.Event()
                    ev_in[i].record(s_in)
                cur.wait_event(ev_in[i])
                if ev_out[i] is not None:
                    cur.wait_event(ev_out[i])                 # previous result in this slot already copied out
                with torch.no_grad():
                    y = m(s["xs"][i])                         # public API: Activation1d.forward
                s["ys"][i] = y
                ev_done[i] = torch.cuda.Event()
                ev_done[i].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_done[i])
                    hy.copy_(y, non_blocking=True)
                    ev_out[i] = torch.cuda.Event()
                    ev_out[i].record(s_out)
                    y.record_stream(s_out)

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0

    # the host-link roofline of this step: the same pinned buffers, device slots and copy streams, no kernels in between
    def copies_only():
        for (hx, hy), s in zip(host, wl.stages):
            n = len(s["xs"])
            for k in range(s["calls"]):
                i = k % n
                with torch.cuda.stream(s_in):
                    s["xs"][i].copy_(hx, non_blocking=True)
                with torch.cuda.stream(s_out):
                    hy.copy_(s["ys"][i], non_blocking=True)

    copies_only()
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    copies_only()
    torch.cuda.synchronize(dev)
    dt_copy = time.perf_counter() - t1
    return dt / steps, h2d, h2d, dt_copy


def run_vocoder(dev, world: int, rank: int, total_clips: int, clips_per_batch: int, t_mel: int, reps: int = 1,
                compare_ncw: bool = True):
    """audio-seconds per second of the whole generator (BASELINE configs 3/4): `total_clips` 10-second binaural
    clips dealt round-robin over ranks; per batch: pinned H2D of the mels -> CUDA-graph replay of the bf16
    channels-last engine (afa_b200/engine.py: cuDNN NHWC convolutions + the fused AMP kernels, tail fused down to
    interleaved int16 stereo PCM, which is what inference_e2e.py:193-205 writes) on L and R -> D2H of the PCM.
    Random-init weights.  `compare_ncw` also times the [B, C, T] harness (reference layout) on one batch."""
    import torch
    import torch.distributed as dist

    from afa_b200 import shard_indices
    from afa_b200.engine import ChannelsLastVocoder, GraphedEngine
    from afa_b200.vocoder import BigVGANGenerator, GraphedVocoder

    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)
    gen = BigVGANGenerator().to(dev)
    with torch.no_grad():
        for n, p in gen.named_parameters():
            if n.endswith("alpha") or n.endswith("beta"):
                p.normal_(0, 0.5)
    gen = gen.bfloat16().eval()
    B = 2 * clips_per_batch
    eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
    ge = GraphedEngine(eng, B, t_mel, want_pcm=True, pcm_interleave=2)
    mine = shard_indices(total_clips, rank, world)
    n_batches = (len(mine) + clips_per_batch - 1) // clips_per_batch
    # one pinned mel buffer and one pinned PCM buffer PER BATCH (distinct inputs, every result kept): U(-12, 2.5) is
    # DiffBinaural's mel clamp range
    n_host = max(1, n_batches)
    mel_hosts = [(torch.rand(B, 80, t_mel) * 14.5 - 12.0).pin_memory() for _ in range(n_host)]
    pcm_hosts = [torch.empty(clips_per_batch, t_mel * gen.hop, 2, dtype=torch.int16).pin_memory() for _ in range(n_host)]
    for _ in range(2):
        pcm_hosts[0].copy_(ge(mel_hosts[0].to(dev, non_blocking=True))[1], non_blocking=True)
    # the only collective of the inference path: gather finished PCM; NCCL has no int16, an interleaved stereo sample pair
    # travels as one int32.  One device buffer per batch so that the gather runs once, after the last batch.
    pcm_dev = [torch.empty_like(ge.static_pcm) for _ in range(n_host)] if world > 1 else None
    gathered = [torch.empty(world, *ge.static_pcm.view(torch.int32).shape, dtype=torch.int32, device=dev) for _ in range(n_host)] if world > 1 else None
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        for bi in range(n_batches):
            _, pcm = ge(mel_hosts[bi].to(dev, non_blocking=True))
            pcm_hosts[bi].copy_(pcm, non_blocking=True)
            if world > 1:
                pcm_dev[bi].copy_(pcm)
    e1.record()
    if world > 1:
        for bi in range(n_batches):
            dist.all_gather_into_tensor(gathered[bi], pcm_dev[bi].view(torch.int32))
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - wall0) / reps
    t = torch.tensor([e0.elapsed_time(e1) / reps, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, wall_ms = float(t[0].item()), float(t[1].item())
    out = {
        "audio_sec_per_sec": round(total_clips * 10.0 / (ms * 1e-3), 1), "unit": "binaural audio-s per wall-s",
        "e2e_wall_audio_sec_per_sec": round(total_clips * 10.0 / (wall_ms * 1e-3), 1),
        "e2e_wall_ms_total": round(wall_ms, 2),
        "e2e_wall_includes": ("host wall clock, max over ranks: pinned H2D of every batch's own mel buffer, graph replay, D2H of every "
                              "batch's PCM into its own pinned buffer, and (N > 1) the NCCL all_gather of all finished PCM, then a device sync"),
        "total_clips": total_clips, "clips_per_batch": clips_per_batch, "ms_total": round(ms, 2),
        "ms_per_clip_per_gpu": round(ms / max(1, len(mine)), 3),
        "dtype": "bf16 generator (channels-last engine), fp32 math inside the fused AMP kernels, int16 PCM out",
        "includes": "pinned H2D of mels, CUDA-graph replay (cuDNN NHWC convs + fused AMP kernels incl. tail), D2H of stereo PCM",
        "params": sum(p.numel() for p in gen.parameters()),
    }
    if rank == 0:
        # BASELINE config 3 as stated: ONE 10-second binaural clip (B = 2: L, R), host mel in -> host PCM out, wall clock per clip
        try:
            ge1 = GraphedEngine(eng, 2, t_mel, want_pcm=True, pcm_interleave=2)
            mel1 = (torch.rand(2, 80, t_mel) * 14.5 - 12.0).pin_memory()
            pcm1 = torch.empty(1, t_mel * gen.hop, 2, dtype=torch.int16).pin_memory()
            lat = []
            for it in range(13):
                torch.cuda.synchronize(dev)
                w0 = time.perf_counter()
                pcm1.copy_(ge1(mel1.to(dev, non_blocking=True))[1], non_blocking=True)
                torch.cuda.synchronize(dev)
                if it >= 3:
                    lat.append((time.perf_counter() - w0) * 1e3)
            lat.sort()
            out["one_clip"] = {"latency_ms_median": round(lat[len(lat) // 2], 3), "latency_ms_min": round(lat[0], 3),
                               "audio_sec_per_sec": round(10.0 / (lat[len(lat) // 2] * 1e-3), 1),
                               "what": "BASELINE config 3: one 10 s binaural clip (B = 2), pinned host mel -> H2D -> CUDA-graph replay of the "
                                       "bf16 channels-last engine -> D2H of the int16 stereo PCM -> sync; host wall clock per clip, 10 clips after 3 warm-ups"}
            del ge1
        except Exception as exc:  # noqa: BLE001
            out["one_clip_error"] = repr(exc)[:120]
    if compare_ncw and rank == 0:
        try:
            del ge, pcm_dev, gathered
            torch.cuda.empty_cache()
            gv = GraphedVocoder(gen, B, t_mel, dtype=torch.bfloat16, device=dev)
            mel_dev = mel_hosts[0].to(dev)
            for _ in range(2):
                gv(mel_dev)
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(3):
                gv(mel_dev)
            e1.record()
            torch.cuda.synchronize(dev)
            ms_ncw = e0.elapsed_time(e1) / 3
            out["ncw_harness_ms_per_clip"] = round(ms_ncw / clips_per_batch, 3)
            out["ncw_harness_audio_sec_per_sec_one_gpu"] = round(clips_per_batch * 10.0 / (ms_ncw * 1e-3), 1)
        except Exception as exc:  # noqa: BLE001
            out["ncw_harness_error"] = repr(exc)[:120]
    return out


def run_vocoder_files(dev, world: int, rank: int, total_clips: int, clips_per_batch: int, t_mel: int, zero_every: int = 8):
    """inference_e2e.py:129-205 as a measured, file-to-file path: `total_clips` synthetic clips as
    left/right mel `.npy` files on tmpfs (every `zero_every`-th clip carries zero frames, as DiffBinaural's padded silence
    does) -> loader thread (np.load into pinned staging buffers, a batch ahead) -> one H2D per batch -> zero-frame
    compaction on the device (afa_compact_zero_frames) -> bf16 channels-last generator (CUDA-graph replay for full-length
    rows, eager per-length launches for rows that lost frames) -> D2H of int16 stereo PCM -> writer thread
    (scipy.io.wavfile.write to tmpfs, off the critical path).  Wall clock, max over ranks."""
    import queue
    import shutil

    import numpy as np
    import torch
    import torch.distributed as dist

    from afa_b200 import ingest, shard_indices
    from afa_b200.engine import ChannelsLastVocoder
    from afa_b200.vocoder import BigVGANGenerator

    root = os.environ.get("AFA_BENCH_TMP", "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")
    work = os.path.join(root, f"afa_bench_npy_{os.environ.get('MASTER_PORT', '0')}")
    if rank == 0:
        shutil.rmtree(work, ignore_errors=True)
        for d in ("left", "right", "wav"):
            os.makedirs(os.path.join(work, d))
        rng = np.random.default_rng(1234)
        for i in range(total_clips):
            for side in ("left", "right"):
                mel = (rng.random((80, t_mel)) * 14.5 - 12.0).astype(np.float32)        # U(-12, 2.5): DiffBinaural's clamp range
                if zero_every and i % zero_every == zero_every - 1:
                    a = int(rng.integers(0, t_mel - 40))
                    mel[:, a : a + int(rng.integers(5, 40))] = 0.0                      # a run of silent (zero) frames
                    if side == "left":
                        mel[:, -7:] = 0.0                                                # left and right keep different counts
                np.save(os.path.join(work, side, f"clip_{i:04d}.npy"), mel)
    if world > 1:
        dist.barrier()
    torch.manual_seed(1234)
    gen = BigVGANGenerator().to(dev)
    with torch.no_grad():
        for n, p in gen.named_parameters():
            if n.endswith("alpha") or n.endswith("beta"):
                p.normal_(0, 0.5)
    gen = gen.bfloat16().eval()
    eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
    bv = ingest.BatchedVocoder(eng, clips_per_batch, t_mel)
    mine = shard_indices(total_clips, rank, world)
    batches = [mine[i : i + clips_per_batch] for i in range(0, len(mine) - len(mine) % clips_per_batch, clips_per_batch)]
    n_stage = 3
    stage = [torch.empty(clips_per_batch, 2, 80, t_mel).pin_memory() for _ in range(n_stage)]
    pcm_host = [torch.empty(clips_per_batch, t_mel * gen.hop, 2, dtype=torch.int16).pin_memory() for _ in range(n_stage)]
    free_in, ready_in, to_write, free_out = queue.Queue(), queue.Queue(), queue.Queue(), queue.Queue()
    for i in range(n_stage):
        free_in.put(i)
        free_out.put(i)

    def loader():
        for b in batches:
            slot = free_in.get()
            buf = stage[slot].numpy()
            for j, ci in enumerate(b):
                buf[j, 0] = ingest.load_mel_npy(os.path.join(work, "left", f"clip_{ci:04d}.npy"))      # inference_e2e.py:140-141
                buf[j, 1] = ingest.load_mel_npy(os.path.join(work, "right", f"clip_{ci:04d}.npy"))
            ready_in.put((slot, b))
        ready_in.put(None)

    def writer():
        while True:
            item = to_write.get()
            if item is None:
                return
            slot, b, ev = item
            ev.synchronize()
            for j, ci in enumerate(b):
                ingest.write_wav(os.path.join(work, "wav", f"clip_{ci:04d}.wav"), 22050, pcm_host[slot][j].numpy())     # :205
            free_out.put(slot)

    # warm-up: one full-length batch (graph replay path) outside the timed region
    bv(torch.zeros(clips_per_batch, 2, 80, t_mel, device=dev) - 5.0)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    def one_pass():
        tl, tw = threading.Thread(target=loader), threading.Thread(target=writer)
        t0 = time.perf_counter()
        tl.start()
        tw.start()
        ragged_rows = 0
        while True:
            item = ready_in.get()
            if item is None:
                break
            slot, b = item
            mel_dev = stage[slot].to(dev, non_blocking=True)
            ev_in = torch.cuda.Event()
            ev_in.record()
            pcm = bv(mel_dev)                                   # synchronises on n_kept: the staging buffer is free after this
            ev_in.synchronize()
            free_in.put(slot)
            out_slot = free_out.get()
            pcm_host[out_slot].copy_(pcm, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            to_write.put((out_slot, b, ev))
        torch.cuda.synchronize(dev)
        t_pcm = time.perf_counter() - t0
        to_write.put(None)
        tl.join()
        tw.join()
        t_wav = time.perf_counter() - t0
        return t_pcm, t_wav

    cold = one_pass()                                   # first sight of the ragged lengths: cuDNN builds plans for them
    if world > 1:
        dist.barrier()
    t_pcm, t_wav = one_pass()                           # the same files again: what a running service sees
    n_done = sum(len(b) for b in batches)
    t = torch.tensor([t_pcm, t_wav], dtype=torch.float64, device=dev)
    cnt = torch.tensor([n_done], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    wavs = len(os.listdir(os.path.join(work, "wav"))) if rank == 0 else None
    if world > 1:
        dist.barrier()
    if rank == 0:
        shutil.rmtree(work, ignore_errors=True)
    clips_done = int(cnt.item())
    return {"first_pass_audio_sec_per_sec": round(n_done * 10.0 / cold[1], 1),
            "first_pass_note": "this rank, first sight of the ragged lengths: cuDNN builds execution plans for every new (rows, length) of the ~110 convolutions",
            "audio_sec_per_sec_pcm_on_host": round(clips_done * 10.0 / float(t[0].item()), 1),
            "audio_sec_per_sec_wav_written": round(clips_done * 10.0 / float(t[1].item()), 1),
            "clips": clips_done, "clips_with_zero_frames": len([i for i in range(total_clips) if zero_every and i % zero_every == zero_every - 1]),
            "wav_files_written": wavs, "clips_per_batch": clips_per_batch, "t_mel": t_mel, "tmpfs": root,
            "unit": "binaural audio-s per wall-s (10 s per clip)",
            "path": ".npy (tmpfs) -> pinned staging (loader thread) -> H2D -> afa_compact_zero_frames -> bf16 channels-last generator -> "
                    "int16 stereo PCM -> D2H -> .wav (tmpfs, writer thread)"}


def run_cl_step(dev, clips: int, t_mel: int, steps: int = 5):
    """The fused AMP kernels of one channels-last generator pass (bf16, CUDA-graph replay, rotating buffers):
    72 x activation+bias, 36 x activation+bias+residual (writes the new residual stream), 6 x resblock mean,
    1 x tail.  GB/s counts algorithmic bytes: 4 B/element, 8 B/element, 14 B/element, (2 C + 6) B/sample."""
    import torch

    from afa_b200 import _lib, functional as F_afa, functional_cl as FC
    from afa_b200.modules import kaiser_sinc_filter1d

    dt = torch.bfloat16
    h = F_afa.host_taps(kaiser_sinc_filter1d(0.25, 0.3, 12))
    B = 2 * clips
    work, bytes_total, launches = [], 0, 0
    for (c, mult, _calls) in AMP_STAGES:
        T = mult * t_mel
        bufs = [torch.randn(B, T, c, device=dev, dtype=dt) for _ in range(4)]
        outs = [torch.empty(B, T, c, device=dev, dtype=dt) for _ in range(3)]
        alpha, beta, bias = (torch.randn(c, device=dev) * 0.5 for _ in range(3))
        work.append((c, T, bufs, outs, alpha, beta, bias))
        n = B * T * c
        bytes_total += 12 * n * 4 + 6 * n * 8 + n * 14
        launches += 19
    c, T, bufs, outs, alpha, beta, bias = work[-1]
    w_post = torch.randn(c, 7, device=dev) * 0.05
    bytes_total += B * T * (2 * c + 6)
    launches += 1
    wave = torch.empty(B, T, device=dev)
    pcm = torch.empty(clips, T, 2, dtype=torch.int16, device=dev)

    def step():
        for (c, T, bufs, outs, alpha, beta, bias) in work:
            for k in range(12):
                FC.amp_activation1d_cl(bufs[k % 4], T, alpha, beta, h, h, True, bias=bias, out=outs[k % 3])
            for k in range(6):
                FC.amp_activation1d_cl(bufs[k % 4], T, alpha, beta, h, h, True, bias=bias, res=bufs[(k + 1) % 4],
                                       xsum=outs[(k + 1) % 3], out=outs[k % 3])
            FC.resblock_mean(bufs[:3], [bufs[3], bufs[0], bufs[1]], bias, 1.0 / 3.0, out=outs[0])
        c, T, bufs, outs, alpha, beta, bias = work[-1]
        FC.tail_cl(bufs[0], T, alpha, beta, h, h, True, w_post, None, wave=wave, pcm=pcm, want_pcm=True)

    step()
    torch.cuda.synchronize(dev)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    elems = sum(B * T * c * 18 for (c, T, *_r) in work)

    # the activation without residual prologue, per call: channels-last tensor-core kernel (tcgen05, the built-in choice)
    # against the channels-last walk kernel, same buffers, same box
    def plain_us(c, T, bufs, outs, alpha, beta, bias, mode):
        _lib.set_tuning(7, mode, 0)
        try:
            g = torch.cuda.CUDAGraph()
            FC.amp_activation1d_cl(bufs[0], T, alpha, beta, h, h, True, bias=bias, out=outs[0])
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(g):
                for k in range(12):
                    FC.amp_activation1d_cl(bufs[k % 4], T, alpha, beta, h, h, True, bias=bias, out=outs[k % 3])
            g.replay()
            torch.cuda.synchronize(dev)
            e0.record()
            for _ in range(steps):
                g.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            return e0.elapsed_time(e1) * 1e3 / (12 * steps)
        finally:
            _lib.set_tuning(7, 1, 0)

    per_stage = []
    for (c, T, bufs, outs, alpha, beta, bias) in work:
        us_walk, us_tc = plain_us(c, T, bufs, outs, alpha, beta, bias, 0), plain_us(c, T, bufs, outs, alpha, beta, bias, 1)
        n = B * T * c
        per_stage.append({"C": c, "T": T, "B": B, "walk_us": round(us_walk, 2), "tensor_core_us": round(us_tc, 2),
                          "tensor_core_gbps": round(4 * n / us_tc / 1e3, 1)})
    return {"ms_per_pass": round(ms, 3), "gbps": round(bytes_total / (ms * 1e-3) / 1e9, 1), "launches_per_pass": launches,
            "activation_elements_per_pass": elems, "gelem_per_s": round(elems / (ms * 1e-3) / 1e9, 1),
            "dtype": "bf16 I/O, f32 math", "layout": "[B, T, C]", "clips": clips,
            "kernels": "afa_tc_cl_fwd_kernel (tcgen05) x72 (no residual prologue), afa_cl_fwd_kernel<bf16,true> x36, afa_mean_kernel x6, "
                       "afa_cl_tail_kernel x1",
            "plain_activation_per_call": {"rows": per_stage,
                                          "how": "12 calls per CUDA-graph replay over 4 input / 3 output buffers (4 B/element algorithmic), "
                                                 "afa_set_tuning(7, 0 | 1): channels-last walk kernel against the channels-last tensor-core kernel"}}


def run_mel_step(dev, batch: int = 32, segment: int = 8192, reps: int = 20):
    """Training-side neighbour: the log-mel work of one BASELINE config 5 step -- the single-scale
    mel of a [batch, segment] waveform (train_binaural_mel.py:711-720) and the seven-scale log mels of estimate and target
    (loss.py:185-200) -- as fused afa_logmel_fwd launches, next to the reference's torch-op chains on the same GPU.
    CUDA events; tensors are L2-resident by nature (1 MB of audio per batch)."""
    import torch

    from afa_b200 import mel as P
    sr = 22050
    hann = {w: torch.hann_window(w, device=dev) for w in (32, 64, 128, 256, 512, 1024, 2048)}

    def chain_single(w, basis):       # the op sequence of meldataset.py:95-118 (torch ops on this GPU: the baseline beside the kernel)
        p = torch.nn.functional.pad(w.unsqueeze(1), (384, 384), mode="reflect").squeeze(1)
        spec = torch.stft(p, 1024, hop_length=256, win_length=1024, window=hann[1024], center=False, pad_mode="reflect",
                          normalized=False, onesided=True, return_complex=True)
        spec = torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1) + 1e-9)
        return torch.log(torch.clamp(torch.matmul(basis, spec), min=1e-5))

    def chain_scale(wav, basis, n):   # loss.py:110-167 + :195-197 for one scale
        B, C, T = wav.shape
        stft = torch.stft(wav.reshape(-1, T), n_fft=n, hop_length=n // 4, window=hann[n], return_complex=True, center=True)
        mag = torch.abs(stft).reshape(B, C, stft.shape[1], stft.shape[2])
        mels = (mag.transpose(2, -1) @ basis.T).transpose(-1, 2)
        return torch.log(mels.clamp(min=1e-5)) / torch.log(torch.tensor(10.0))

    torch.manual_seed(1234)
    y = (0.3 * torch.randn(batch, segment, device=dev)).clamp(-1, 1)
    yh = (y + 0.05 * torch.randn_like(y)).clamp(-1, 1)
    plan1 = P.MelPlan(1024, torch.hann_window(1024, dtype=torch.float64), P.slaney_mel_filterbank(sr, 1024, 80), dev)
    msl = P.MultiScaleMelSpectrogramLoss(sr)
    bases = [torch.from_numpy(P.slaney_mel_filterbank(sr, w, nm)).to(dev) for w, nm in zip(msl.window_lengths, msl.n_mels)]
    basis1 = torch.from_numpy(P.slaney_mel_filterbank(sr, 1024, 80)).to(dev)
    y3, yh3 = y.unsqueeze(1), yh.unsqueeze(1)

    def ours_single():
        return P.logmel(yh, plan1, 256, 384)

    def ours_multi():
        return [(msl.log_mels(yh3, s), msl.log_mels(y3, s)) for s in range(7)]

    def torch_single():
        return chain_single(yh, basis1)

    def torch_multi():
        return [(chain_scale(yh3, bases[s], msl.window_lengths[s]), chain_scale(y3, bases[s], msl.window_lengths[s]))
                for s in range(7)]

    l1 = torch.nn.functional.l1_loss

    def ours_loss_step():             # the loss_mel term of the step: forward + backward to d loss / d estimate
        x = yh3.detach().requires_grad_(True)
        with torch.enable_grad():
            msl(x, y3).backward()
        return x.grad

    def torch_loss_step():
        x = yh3.detach().requires_grad_(True)
        with torch.enable_grad():
            sum(l1(chain_scale(x, bases[s], msl.window_lengths[s]), chain_scale(y3, bases[s], msl.window_lengths[s])) for s in range(7)).backward()
        return x.grad

    def timed(fn):
        best = float("inf")
        with torch.no_grad():
            for _ in range(3):
                fn()
            for _trial in range(3):           # best of 3 trials of `reps` calls: the host side of an eager launch is noisy
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    fn()
                e1.record()
                torch.cuda.synchronize(dev)
                best = min(best, e0.elapsed_time(e1) / reps * 1e3)
        return best

    with torch.no_grad():
        err = float((ours_single() - torch_single()).abs().max())
    frames = sum(2 * batch * P.num_frames(segment, w, w // 4, w // 2) for w in msl.window_lengths)
    # inference side (get_mel_spectrogram at BigVGAN/inference_binaural.py:131-132: mel of whole clips): 16 ten-second rows = 8 binaural clips
    clips = (0.3 * torch.randn(16, 220416, device=dev)).clamp(-1, 1)
    clip_us = timed(lambda: P.logmel(clips, plan1, 256, 384))
    clip_torch_us = timed(lambda: chain_single(clips, basis1))
    out = {"shape": [batch, segment], "single_scale_us": round(timed(ours_single), 2), "single_scale_torch_ops_us": round(timed(torch_single), 2),
           "multi_scale_14_launches_us": round(timed(ours_multi), 2), "multi_scale_torch_ops_us": round(timed(torch_multi), 2),
           "multi_scale_frames": frames, "max_abs_diff_vs_torch_ops_log_mel": err,
           "clips_16x220416_us": round(clip_us, 2), "clips_16x220416_torch_ops_us": round(clip_torch_us, 2),
           "clips_algorithmic_gbps": round(16 * (220416 + 80 * 861) * 4 / (clip_us * 1e-6) / 1e9, 1),
           "multi_scale_loss_fwd_bwd_us": round(timed(ours_loss_step), 2), "multi_scale_loss_fwd_bwd_torch_ops_us": round(timed(torch_loss_step), 2),
           "what": "eager calls incl. host overhead, CUDA events, best of 3 x 20; loss step = forward + backward to d loss / d estimate: 14 afa_logmel_fwd + 7 afa_l1_partial_sums + 7 afa_logmel_bwd (2 kernels each); device time per step under ncu: 0.6 ms in ~50 launches vs 1.6 ms in 330 for the torch-op chain (profiles/r01_mel_loss_step_launches_*.txt)"}
    return out


# ----------------------------------------------------------------------------------------------
# per-shape timing (shared by the main line's `roofline_per_shape` and by --table)
# ----------------------------------------------------------------------------------------------
def time_shape(dev, b: int, c: int, t: int, dname: str, which: str, torch_baseline: bool = False, reps: int = 3):
    """One Activation1d forward (`fwd`) or backward (`bwd`) of a [b, c, t] tensor through the C ABI: CUDA-graph replay over
    rotating buffer sets that exceed L2 (L2-cold), CUDA events.  Returns a dict with microseconds per call, algorithmic GB/s
    (fwd 2 x, bwd 3 x element size per element) and the fraction of the measured HBM peak."""
    import torch

    from afa_b200 import Activation1d
    from afa_b200 import functional as Fn
    from afa_b200.activations import SnakeBeta
    from oracle import torch_path as TP

    dtype = torch.float32 if dname == "fp32" else torch.bfloat16
    es = 4 if dname == "fp32" else 2
    peak, _ = measured_peak()
    act = SnakeBeta(c, alpha_logscale=True)
    with torch.no_grad():
        act.alpha.normal_(0, 0.5)
        act.beta.normal_(0, 0.5)
    m = Activation1d(activation=act).to(dev)
    n = b * c * t
    nbuf = max(2, min(24, int(1.0e9 // (n * es * 2))))
    xs = [torch.randn(b, c, t, device=dev).to(dtype) for _ in range(nbuf)]
    ys = [torch.randn(b, c, t, device=dev).to(dtype) for _ in range(nbuf)]
    tu, td = m._host_taps()
    a_, b_ = m.act.alpha.detach(), m.act.beta.detach()
    if which == "fwd":
        def call(i):
            Fn.activation1d_forward_raw(xs[i % nbuf], a_, b_, tu, td, True, out=ys[i % nbuf])
    else:
        def call(i):
            Fn.activation1d_backward_raw(xs[i % nbuf], ys[i % nbuf], a_, b_, tu, td, True)
    iters = 2 * nbuf
    call(0)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            call(i)
    g.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    us = e0.elapsed_time(e1) * 1e3 / (reps * iters)
    bpe = (2 if which == "fwd" else 3) * es
    gbs = n * bpe / us / 1e3
    row = {"dtype": dname, "dir": which, "B": b, "C": c, "T": t, "us": round(us, 2), "GBps": round(gbs, 1),
           "frac_of_measured_peak": round(gbs / peak, 3), "Gelem_per_s": round(n / us / 1e3, 1)}
    if torch_baseline and which == "fwd":
        taps = m.upsample.filter.to(dtype)
        with torch.no_grad():
            TP.activation1d_torch(xs[0], a_.to(dtype), b_.to(dtype), True, taps, taps)
            torch.cuda.synchronize(dev)
            e0.record()
            for i in range(3):
                TP.activation1d_torch(xs[i % nbuf], a_.to(dtype), b_.to(dtype), True, taps, taps)
            e1.record()
            torch.cuda.synchronize(dev)
        row["torch_ops_gpu_us"] = round(e0.elapsed_time(e1) * 1e3 / 3, 1)
    del xs, ys, g
    torch.cuda.empty_cache()
    return row


def per_shape_rooflines(dev, t_mel: int):
    """fp32 + bf16, forward + backward, on: BASELINE config 1 (2, 512, 8192; target <= 12.0 us = 70 % of 8 TB/s for the fp32
    forward), one binaural clip (B = 2), eight clips (B = 16) and the config-5 training shapes (B = 32, T_mel = 32).
    Compact rows: [dtype, dir, B, C, T, us, GB/s, fraction of the measured HBM peak]."""
    shapes = [("cfg1", 2, 512, 8192)]
    for clips, tag in ((1, "clip1"), (8, "clips8")):
        shapes += [(tag, b, c, t) for (b, c, t, _) in stage_shapes(clips, t_mel)]
    shapes += [("train", 32, c, mult * 32) for (c, mult, _) in AMP_STAGES]
    rows = []
    for dname in ("fp32", "bf16"):
        for which in ("fwd", "bwd"):
            for (tag, b, c, t) in shapes:
                r = time_shape(dev, b, c, t, dname, which)
                rows.append([tag, dname, which, b, c, t, r["us"], r["GBps"], r["frac_of_measured_peak"]])
    def worst(tag, dname, which):
        v = [r[8] for r in rows if r[0] == tag and r[1] == dname and r[2] == which]
        return [min(v), round(sum(v) / len(v), 3), max(v)] if v else None
    summary = {f"{dname}_{which}_{tag}": worst(tag, dname, which) for dname in ("fp32", "bf16") for which in ("fwd", "bwd")
               for tag in ("cfg1", "clip1", "clips8", "train")}
    cfg1 = next(r for r in rows if r[0] == "cfg1" and r[1] == "fp32" and r[2] == "fwd")
    return {"columns": ["set", "dtype", "dir", "B", "C", "T", "us", "GBps", "frac_of_measured_peak"], "rows": rows,
            "min_mean_max_frac": summary, "cfg1_fp32_fwd_us": cfg1[6], "cfg1_target_us": 12.0,
            "how": "one call per row, CUDA-graph replay over rotating buffer sets > L2, CUDA events; algorithmic bytes: forward 2 x, backward 3 x element size"}


def gpu_torch_baseline(dev, clips: int, t_mel: int):
    """SURVEY.md section 8(d) 'GPU baselines': the reference's torch-op chain (oracle/torch_path.py, the same ATen calls as
    alias_free_activation/*.py) on THIS GPU, one call per AMP stage shape of the main workload, fp32, next to the fused kernel."""
    rows, t_torch, t_fused = [], 0.0, 0.0
    for (b, c, t, calls) in stage_shapes(clips, t_mel):
        r = time_shape(dev, b, c, t, "fp32", "fwd", torch_baseline=True)
        rows.append({"shape": [b, c, t], "calls": calls, "fused_us": r["us"], "torch_ops_us": r["torch_ops_gpu_us"]})
        t_torch += calls * r["torch_ops_gpu_us"]
        t_fused += calls * r["us"]
    return {"per_stage": rows, "pass_ms_torch_ops": round(t_torch / 1e3, 3), "pass_ms_fused": round(t_fused / 1e3, 3),
            "speedup": round(t_torch / t_fused, 2), "what": "reference torch-op Activation1d on the same B200, fp32, eager, 109 calls weighted"}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from afa_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device for the GPU arm (there is no CPU fallback); use --impl reference")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")     # the `train` object captures DDP's all-reduce in a CUDA graph
        dist.init_process_group(backend="nccl", device_id=dev)
    _lib.load_library()
    if args.chunks:
        _lib.set_tuning(0, args.chunks, 0)
    dtype = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    esize = 4 if args.dtype == "fp32" else 2

    wl = Workload(dev, args.clips, args.t_mel, dtype)
    bytes_per_step = wl.elements * 2 * esize
    shapes_calls = [list(s["shape"]) + [s["calls"]] for s in wl.stages]
    n_elements, n_launches = wl.elements, wl.launches
    min_call_mb = min(s["shape"][0] * s["shape"][1] * s["shape"][2] for s in wl.stages) * esize * 2 / 1e6

    # warm-up (also fills the host tap caches), then capture the step in a CUDA graph
    for _ in range(max(1, args.warmup if not args.graph else 1)):
        wl.step_device()
    torch.cuda.synchronize(dev)
    graph = None
    if args.graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            wl.step_device()
        for _ in range(args.warmup):
            graph.replay()
        torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                wl.step_device()
        e1.record()
        torch.cuda.synchronize(dev)
    barrier()
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * bytes_per_step / (ms_per_step * 1e-3) / 1e9
    lib_launches = _lib.launch_count() - launches0
    gpu_launches = args.steps * wl.launches if graph is not None else lib_launches

    # roofline of the dominant kernel (every launch in the step is afa_fwd_kernel; average over the timed region)
    peak, peak_src = measured_peak()
    per_launch_bytes = bytes_per_step / wl.launches
    per_launch_us = ms_per_step * 1e3 / wl.launches
    achieved = per_launch_bytes / (per_launch_us * 1e-6) / 1e9
    kb, kc, kt = max((s["shape"] for s in wl.stages), key=lambda sh: sh[2])          # the long-row (dominant) variant
    kinfo = _lib.kernel_info(0, 0 if args.dtype == "fp32" else 1, kt, kb, kc)
    # DRAM traffic of the SAME launch (kernel, dtype, shape) from the committed ncu capture; its algorithmic size beside it
    trec = traffic_record("afa::afa_fwd_kernel", args.dtype, (kb, kc, kt))
    traffic = None if trec is None else trec["dram_bytes"]

    # e2e through the module with host buffers
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        sec, h2d, d2h, sec_copy = run_e2e(wl, e2e_steps, 1)
        t2 = torch.tensor([sec, sec_copy], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        sec, sec_copy = float(t2[0].item()), float(t2[1].item())
        # value counts algorithmic bytes (in + out) = the bytes that cross the host link (h2d + d2h): same unit as the peak
        pcie_peak = world * (h2d + d2h) / sec_copy / 1e9
        e2e = {"value": round(world * bytes_per_step / sec / 1e9, 3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
               "ms_per_step": round(sec * 1e3, 3),
               "pcie_peak_gbs": round(pcie_peak, 3),
               "frac_of_pcie_peak": round(world * bytes_per_step / sec / 1e9 / pcie_peak, 3),
               "pcie_peak_how": ("the step's own pinned buffers, device slots and two copy streams with the kernels removed (H2D and D2H "
                                 "concurrently), all ranks at once, max over ranks: the host link + host memory ceiling this e2e figure "
                                 "is bound by; every intermediate activation crosses PCIe here, which no caller of the op does -- see "
                                 "vocoder.e2e_wall_audio_sec_per_sec for the realistic end-to-end"),
               "api": "afa_b200.Activation1d.forward (ctypes -> afa_activation1d_fwd), pinned host buffers"}

    cpu_base = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cores = host_cores()
        cpu_reference_pass(args.t_mel, 1, repeats=1)
        elems, times = cpu_reference_pass(args.t_mel, 1, repeats=args.cpu_repeats)
        cpu_base = {"value": round(elems * 8 / min(times) / 1e9, 4), "unit": UNIT, "cores": cores, "kind": "port", "cpu": cpu_model(),
                    "sample": (f"one Activation1d call per AMP stage shape (6 calls, B=2, T_mel={args.t_mel}, fp32, "
                               f"{elems} elements), torch CPU ops (oracle/torch_path.py), {cores} threads on {cpu_model()}, best of {args.cpu_repeats}")}

    used_graph = graph is not None
    vocoder = None
    if not args.no_vocoder:
        del wl, graph
        torch.cuda.empty_cache()
        try:
            # weak-scaling companion number: `--clips` clips per GPU through the whole generator
            vocoder = run_vocoder(dev, world, rank, args.clips * world, min(4, args.clips), args.t_mel)
        except Exception as exc:  # noqa: BLE001  (the headline metric must still print)
            vocoder = {"error": repr(exc)[:200]}
        try:
            torch.cuda.empty_cache()
            vocoder["files"] = run_vocoder_files(dev, world, rank, 2 * args.clips * world, min(4, args.clips), args.t_mel)
        except Exception as exc:  # noqa: BLE001
            vocoder["files"] = {"error": repr(exc)[:200]}
        wl = None
    channels_last = None
    if not args.no_vocoder and rank == 0 and world == 1:
        try:
            torch.cuda.empty_cache()
            channels_last = run_cl_step(dev, args.clips, args.t_mel)
        except Exception as exc:  # noqa: BLE001
            channels_last = {"error": repr(exc)[:200]}

    mel = None
    if not args.no_vocoder and rank == 0 and world == 1:
        try:
            mel = run_mel_step(dev)
        except Exception as exc:  # noqa: BLE001
            mel = {"error": repr(exc)[:200]}

    per_shape = torch_gpu = None
    if rank == 0 and world == 1 and not args.no_per_shape:
        try:
            torch.cuda.empty_cache()
            per_shape = per_shape_rooflines(dev, args.t_mel)
            torch_gpu = gpu_torch_baseline(dev, args.clips, args.t_mel)
        except Exception as exc:  # noqa: BLE001
            per_shape = {"error": repr(exc)[:200]}
    bf16_step = None
    if rank == 0 and world == 1 and not args.no_per_shape and args.dtype == "fp32":
        # the same 109-call step with bf16 tensors (BASELINE config 2 names fp32 + bf16): the tensor-core forward
        # wherever it is eligible, the register-walk kernel elsewhere; same timing rules as the headline
        try:
            torch.cuda.empty_cache()
            wl16 = Workload(dev, args.clips, args.t_mel, torch.bfloat16)
            wl16.step_device()
            torch.cuda.synchronize(dev)
            g16 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g16):
                wl16.step_device()
            for _ in range(max(3, args.warmup)):
                g16.replay()
            torch.cuda.synchronize(dev)
            with ClockSampler(local_rank) as clocks16:
                e0.record()
                for _ in range(args.steps):
                    g16.replay()
                e1.record()
                torch.cuda.synchronize(dev)
            ms16 = e0.elapsed_time(e1) / args.steps
            by16 = wl16.elements * 4
            n_tc = sum(s["calls"] for s in wl16.stages if (s["shape"][2] % 8 == 0 or (s["shape"][2] % 4 == 0 and (s["shape"][0] * s["shape"][1]) % 2 == 0))
                       and s["shape"][2] >= 256 and s["shape"][0] * s["shape"][1] * s["shape"][2] >= (6 << 20))
            bf16_step = {"metric": METRIC, "value": round(by16 / (ms16 * 1e-3) / 1e9, 3), "unit": UNIT, "dtype": "bf16 I/O, f32 math",
                         "ms_per_step": round(ms16, 4), "steps": args.steps, "frac_of_measured_peak": round(by16 / (ms16 * 1e-3) / 1e9 / peak, 4),
                         "algorithmic_bytes_per_step": by16, "launches_per_step": wl16.launches,
                         "kernels": f"{n_tc} x afa_tc::afa_tc_fwd_kernel (tcgen05), {wl16.launches - n_tc} x afa::afa_fwd_kernel",
                         "clocks": clocks16.summary()}
            del wl16, g16
        except Exception as exc:  # noqa: BLE001
            bf16_step = {"error": repr(exc)[:200]}
    train = None
    if not args.no_train:
        try:
            torch.cuda.empty_cache()
            train = train_step_bench(args, dev, world, rank, local_rank, min(5, max(1, args.steps)), 3, profile_share=True)
        except Exception as exc:  # noqa: BLE001
            train = {"error": repr(exc)[:200]}

    if world > 1 and wl is not None:
        # the only collective: gather one checksum per rank (stands in for gathering finished waveforms)
        chk = torch.stack([s["ys"][0].float().abs().mean() for s in wl.stages]).sum().reshape(1)
        outs = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(outs, chk)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if args.dtype == "fp32" else "bf16 I/O, f32 math", "data": "synthetic",
            "config": {
                "workload": (f"bigvgan_binaural_22khz_80band_256x: all 109 AMP Activation1d forwards of one generator pass, "
                             f"{args.clips} x 10 s binaural clips per GPU (B={2 * args.clips}, T_mel={args.t_mel}), SnakeBeta logscale"),
                "shapes_BCT_calls": shapes_calls,
                "elements_per_step": n_elements, "algorithmic_bytes_per_step": bytes_per_step,
                "l2": "every call's tensors exceed L2 (>=%.0f MB in + out per call) and rotate over 2 buffer sets" % min_call_mb,
                "cuda_graph": used_graph, "parallelism": f"clip-sharded dp{world}, no collective in the path",
            },
            "roofline": {"bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                         "kernel": "afa::afa_fwd_kernel", "avg_launch_us": round(per_launch_us, 3),
                         "algorithmic_bytes_per_launch": int(per_launch_bytes), "kernel_info": kinfo,
                         "traffic_launch": None if trec is None else {
                             "shape": trec["shape"], "algorithmic_bytes": trec["algorithmic_bytes"], "dram_read_bytes": trec["dram_read_bytes"],
                             "traffic_over_algorithmic": round(trec["dram_bytes"] / trec["algorithmic_bytes"], 3), "source": trec["source"],
                             "note": "per launch of the largest stage shape of this workload (not the 109-launch average printed above)"},
                         "tensor_core_forward": traffic_record("afa_tc::afa_tc_fwd_kernel", "bf16", (16, 384, 13776)),
                         "tensor_core_forward_channels_last": traffic_record("afa_tc::afa_tc_cl_fwd_kernel", "bf16", (16, 55104, 96))},
            "cpu_baseline": cpu_base,
            "e2e": e2e,
            "gpu_launches": int(gpu_launches),
            "clocks": clocks.summary(),
            "audio_sec_per_sec_activation_only": round(world * args.clips * 10.0 / (ms_per_step * 1e-3), 2),
            "vocoder": vocoder,
            "channels_last_amp_kernels": channels_last,
            "log_mel": mel,
            "roofline_per_shape": per_shape,
            "bf16_step": bf16_step,
            "gpu_torch_baseline": torch_gpu,
            "train": train,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_vocoder_mode(args):
    """`--mode vocoder`: BASELINE config 3/4 as the main line -- `--total-clips` clips, strong scaling over ranks."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group(backend="nccl", device_id=dev)
    with ClockSampler(local_rank) as clocks:
        v = run_vocoder(dev, world, rank, args.total_clips, min(4, args.total_clips), args.t_mel, reps=max(1, args.steps))
        if args.from_npy:
            torch.cuda.empty_cache()
            v["files"] = run_vocoder_files(dev, world, rank, args.total_clips, min(4, args.total_clips), args.t_mel)
    if rank == 0:
        print(json.dumps({
            "metric": "vocoded_audio_sec_per_sec", "value": v["audio_sec_per_sec"], "unit": "audio-s/s", "n_gpus": world,
            "steps": max(1, args.steps), "warmup": 2, "ms_per_step": v["ms_total"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"full BigVGAN generator (bigvgan_binaural_22khz_80band_256x, random init), "
                                   f"{args.total_clips} x 10 s binaural clips sharded by clip over {world} GPU(s)",
                       "parallelism": f"clip-sharded dp{world}, NCCL only to gather waveforms"},
            "vocoder": v, "clocks": clocks.summary(),
        }))
    if world > 1:
        dist.destroy_process_group()
    return 0


def train_step_bench(args, dev, world: int, rank: int, local_rank: int, steps: int, warmup: int, profile_share: bool = True):
    """BASELINE config 5 -- the generator's forward + backward with the fused Activation1d forward and
    backward kernels in it (train_binaural_mel.py:431-543, 787-791: BigVGAN(h) under DDP, AdamW, clip_grad_norm_), at
    the training segment (8192 samples = 32 mel frames, `--t-mel 32`) and `--batch` items per GPU.  The discriminators
    and losses of the reference are not importable here (nnAudio, librosa, pesq, auraloss absent: SURVEY.md section 8c), so
    the step is generator forward -> synthetic scalar loss -> backward -> clip -> AdamW, fp32 as the reference trains.
    `--torch-baseline` also times the same step with the reference's torch-op activation on the same GPU."""
    import torch
    import torch.distributed as dist
    import torch.nn as nn

    from afa_b200 import _lib
    from afa_b200.vocoder import BigVGANGenerator

    torch.backends.cudnn.benchmark = True
    t_mel = args.t_mel if args.t_mel != T_MEL_10S else 32
    B = args.batch
    ddp_opts = {"gradient_as_bucket_view": True, "bucket_cap_mb": int(os.environ.get("AFA_DDP_BUCKET_MB", "100")),
                "static_graph": os.environ.get("AFA_DDP_STATIC", "1") == "1",
                # DDP's default, as train_binaural_mel.py:541 runs it.  (Round 1 lost 20 ms per step here: the in-place re-broadcast of
                # the constant filter buffers made each of the 109 fused modules re-read its taps with a synchronising device-to-host
                # copy; the tap cache now follows reloads, not version bumps -- profiles/r02_ddp_probe_n2.log.)
                "broadcast_buffers": os.environ.get("AFA_DDP_BROADCAST_BUFFERS", "1") == "1"}
    # Whole-step CUDA graph.  One GPU: only with --graph-step (the eager step is device-bound there).  Under DDP it is the
    # default: DDP's per-parameter autograd hooks and bucket bookkeeping make the eager step HOST-bound (profiles/
    # r02_ddp_probe_n2.log: 41 ms of kernels, 1 ms of NCCL, 58 ms per step), and a captured step replays with no host work.
    # Recipe of the CUDA-graphs note: DDP built and warmed up (>= 11 steps) on a side stream, NCCL async error handling off.
    want_graph = bool(args.graph_step) or (world > 1 and os.environ.get("AFA_TRAIN_GRAPH", "1") == "1")

    def build(factory=None):
        torch.manual_seed(1234)                                    # configs/bigvgan_binaural_22khz_80band_256x.json:9
        gen = BigVGANGenerator() if factory is None else BigVGANGenerator(activation_factory=factory)
        with torch.no_grad():
            for n, p in gen.named_parameters():
                if n.endswith("alpha") or n.endswith("beta"):
                    p.normal_(0, 0.5)
        gen = gen.to(dev).train()
        # train_binaural_mel.py:540-543 wraps the generator in plain DDP; bucket views avoid one gradient copy per step
        if world > 1:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                model = nn.parallel.DistributedDataParallel(gen, device_ids=[local_rank], **ddp_opts)
            torch.cuda.current_stream(dev).wait_stream(side)
        else:
            model = gen
        if world > 1 and os.environ.get("AFA_DDP_BF16_HOOK", "0") == "1":
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks

            model.register_comm_hook(None, default_hooks.bf16_compress_hook)
        opt = torch.optim.AdamW(gen.parameters(), 5e-5, betas=(0.8, 0.99),       # train_binaural_mel.py:548-551, config lr / betas
                                capturable=want_graph)
        return gen, model, opt

    def make_loss(fused: bool):
        """`--mel-loss`: the reference's generator-side mel term, loss_mel = MultiScaleMelSpectrogramLoss(y, y_g_hat) * 60
        (train_binaural_mel.py:759, lambda_melloss = 60) plus the single-scale mel of the estimate it computes every step
        (:711-720), on a synthetic target -- through the fused kernels (fused=True) or the reference's torch-op chain."""
        if not args.mel_loss:
            return lambda y: y.abs().mean()
        from afa_b200 import mel as P

        torch.manual_seed(4321)
        target = (0.3 * torch.randn(B, 1, t_mel * 256, device=dev)).clamp(-1, 1)
        if fused:
            msl = P.MultiScaleMelSpectrogramLoss(22050)

            def loss_of(y):
                P.mel_spectrogram(y.detach().squeeze(1), 1024, 80, 22050, 256, 1024, 0, None, check_range=False)
                return msl(target, y) * 60.0
            return loss_of
        wins, nms = (32, 64, 128, 256, 512, 1024, 2048), (5, 10, 20, 40, 80, 160, 320)
        bases = [torch.from_numpy(P.slaney_mel_filterbank(22050, w, nm)).to(dev) for w, nm in zip(wins, nms)]
        basis1 = torch.from_numpy(P.slaney_mel_filterbank(22050, 1024, 80)).to(dev)
        hann = {w: torch.hann_window(w, device=dev) for w in wins}
        log10 = torch.log(torch.tensor(10.0))

        def logmels(wav, basis, n):       # loss.py:110-167 + :195-197
            stft = torch.stft(wav.reshape(-1, wav.shape[-1]), n_fft=n, hop_length=n // 4, window=hann[n], return_complex=True, center=True)
            mels = (torch.abs(stft).transpose(1, 2) @ basis.T).transpose(1, 2)
            return torch.log(mels.clamp(min=1e-5)) / log10

        def loss_of(y):
            yp = torch.nn.functional.pad(y.detach(), (384, 384), mode="reflect").squeeze(1)       # meldataset.py:95-118
            spec = torch.stft(yp, 1024, hop_length=256, win_length=1024, window=hann[1024], center=False, return_complex=True)
            torch.log(torch.clamp(basis1 @ torch.sqrt(torch.view_as_real(spec).pow(2).sum(-1) + 1e-9), min=1e-5))
            return sum(torch.nn.functional.l1_loss(logmels(target, b, w), logmels(y, b, w)) for b, w in zip(bases, wins)) * 60.0
        return loss_of

    def time_steps(model, gen, opt, steps, warmup, fused=True):
        mel = torch.rand(B, 80, t_mel, device=dev) * 14.5 - 12.0
        loss_of = make_loss(fused)

        def one():
            opt.zero_grad(set_to_none=True)
            y = model(mel)
            loss = loss_of(y)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(gen.parameters(), 500.0)     # config clip_grad_norm, train_binaural_mel.py:788-790
            opt.step()
            return loss

        graph = None
        if want_graph and fused:
            # the whole step (forward, fused backward kernels, DDP's bucketed all-reduce, clip, AdamW) captured once and replayed
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(max(warmup, 11 if world > 1 else warmup)):
                    one()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            opt.zero_grad(set_to_none=False)
            with torch.cuda.graph(graph):
                y = model(mel)
                static_loss = loss_of(y)
                static_loss.backward()
                torch.nn.utils.clip_grad_norm_(gen.parameters(), 500.0)
                opt.step()
                opt.zero_grad(set_to_none=False)
            for _ in range(2):
                graph.replay()
        else:
            for _ in range(warmup):
                one()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            if graph is not None:
                graph.replay()
                loss = static_loss
            else:
                loss = one()
        e1.record()
        torch.cuda.synchronize(dev)
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        per_step = (_lib.launch_count() - l0) // steps if graph is None else 109 * 3     # 109 forward + 109 backward + 109 finalize launches
        return float(t.item()), float(loss.item()), per_step

    with ClockSampler(local_rank) as clocks:
        gen, model, opt = build()
        n_params = sum(p.numel() for p in gen.parameters())
        ms, loss, launches = time_steps(model, gen, opt, max(1, steps), max(3, warmup))
    # share of the fused activation kernels in the step (device time, one profiled step on rank 0)
    act_ms = tot_ms = nccl_ms = None
    if profile_share:
        # one extra step on every rank (DDP's all-reduce needs all of them); rank 0 records it with the kernel profiler
        try:
            import contextlib

            from torch.profiler import ProfilerActivity, profile

            mel = torch.rand(B, 80, t_mel, device=dev) * 14.5 - 12.0
            ctx = profile(activities=[ProfilerActivity.CUDA]) if rank == 0 else contextlib.nullcontext()
            with ctx as prof:
                opt.zero_grad(set_to_none=True)
                model(mel).abs().mean().backward()
                torch.cuda.synchronize(dev)
            if rank == 0:
                ka = prof.key_averages()
                act_ms = sum(r.device_time_total for r in ka if "afa" in r.key) / 1e3
                nccl_ms = sum(r.device_time_total for r in ka if "nccl" in r.key.lower()) / 1e3
                tot_ms = sum(r.device_time_total for r in ka) / 1e3
        except Exception:  # noqa: BLE001
            act_ms = tot_ms = nccl_ms = None
    base = None
    if args.torch_baseline:
        del gen, model, opt
        torch.cuda.empty_cache()
        from afa_b200.modules import DownSample1d, UpSample1d
        from oracle import torch_path as TP                       # the reference's op sequence (optional diagnostic leg only)

        class TorchOpActivation1d(nn.Module):
            def __init__(self, activation):
                super().__init__()
                self.act, self.upsample, self.downsample = activation, UpSample1d(2, 12), DownSample1d(2, 12)

            def forward(self, x):
                beta = getattr(self.act, "beta", None)
                return TP.activation1d_torch(x, self.act.alpha, beta, bool(self.act.alpha_logscale), self.upsample.filter,
                                             self.downsample.lowpass.filter)

        gen_t, model_t, opt_t = build(TorchOpActivation1d)
        ms_t, loss_t, _ = time_steps(model_t, gen_t, opt_t, max(1, steps), 3, fused=False)
        base = {"ms_per_step": round(ms_t, 3), "loss": loss_t, "what": "same step, reference torch-op Activation1d on the same GPU"}
    samples = world * B * t_mel * 256
    result = {
        "metric": "train_step_generator_fwd_bwd_ms", "value": round(ms, 3), "unit": "ms per step", "n_gpus": world,
        "steps": max(1, steps), "warmup": max(3, warmup), "ms_per_step": round(ms, 3), "higher_is_better": False,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"BASELINE config 5: BigVGAN generator (bigvgan_binaural_22khz_80band_256x, random init) forward + "
                                f"backward + clip_grad_norm + AdamW, batch {B} per GPU, segment {t_mel * 256} samples (T_mel={t_mel}), "
                                + ("multi-scale mel loss x 60 + single-scale mel of the estimate (fused log-mel kernels, forward + adjoint)"
                                   if args.mel_loss else "synthetic scalar loss")
                                + "; fused Activation1d forward and backward ([B, C, T] kernels)"),
                   "parallelism": f"DDP dp{world}" if world > 1 else "single GPU",
                   "ddp": ddp_opts if world > 1 else None,
                   "collective": (f"DDP bucketed NCCL all-reduce of the fp32 gradients ({n_params * 4 / 1e6:.0f} MB per step, "
                                  f"{ddp_opts['bucket_cap_mb']} MB buckets, overlapped with backward)") if world > 1 else None,
                   "cuda_graph_step": want_graph},
        "audio_sec_per_sec_trained": round(samples / 22050.0 / (ms * 1e-3), 1),
        "activation_kernels_ms_per_step": None if act_ms is None else round(act_ms, 3),
        "all_kernels_ms_per_step": None if act_ms is None else round(tot_ms, 3),
        "nccl_kernels_ms_per_step": None if nccl_ms is None else round(nccl_ms, 3),
        "gpu_launches": int(launches) * max(1, steps), "loss": loss, "torch_op_activation_baseline": base,
        "clocks": clocks.summary(),
    }
    return result


def run_train_mode(args):
    """`--mode train`: BASELINE config 5 as the main line (see train_step_bench)."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")     # required to capture DDP's all-reduce in a CUDA graph
        dist.init_process_group(backend="nccl", device_id=dev)
    result = train_step_bench(args, dev, world, rank, local_rank, max(1, args.steps), max(3, args.warmup))
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ----------------------------------------------------------------------------------------------
# per-shape table (extra; not the driver's contract)
# ----------------------------------------------------------------------------------------------
def run_table(args):
    import torch

    from afa_b200 import Activation1d, _lib
    from afa_b200 import functional as Fn
    from afa_b200.activations import SnakeBeta
    from oracle import torch_path as TP

    dev = torch.device("cuda:0")
    peak, _ = measured_peak()
    rows = []
    shapes = [(2, 512, 8192)]
    for clips in (1, 8):
        shapes += [(b, c, t) for (b, c, t, _) in stage_shapes(clips, args.t_mel)]
    shapes += [(32, c, mult * 32) for (c, mult, _) in AMP_STAGES]
    for dname in ("fp32", "bf16"):
        for which in ("fwd", "bwd"):
            for (b, c, t) in shapes:
                row = time_shape(dev, b, c, t, dname, which, torch_baseline=args.torch_baseline)
                rows.append(row)
                print(json.dumps(row), file=sys.stderr, flush=True)
    out = os.path.join(REPO, "gpurun_out", "shape_table.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    with open(out, "w") as f:
        json.dump(rows, f, indent=1)
    print(json.dumps({"table_rows": len(rows), "written": out}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--clips", type=int, default=8, help="10 s binaural clips per GPU (B = 2 x clips)")
    ap.add_argument("--t-mel", dest="t_mel", type=int, default=T_MEL_10S)
    ap.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32")
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-vocoder", action="store_true", help="skip the whole-generator audio-s/s companion number")
    ap.add_argument("--no-per-shape", action="store_true", help="skip roofline_per_shape / gpu_torch_baseline (one GPU only)")
    ap.add_argument("--no-train", action="store_true", help="skip the BASELINE config 5 training-step object")
    ap.add_argument("--mode", choices=["activation", "vocoder", "train"], default="activation",
                    help="vocoder: BASELINE config 4 (--total-clips clips sharded over the ranks, strong scaling); "
                         "train: BASELINE config 5 (generator forward + backward step, --batch per GPU, segment 8192)")
    ap.add_argument("--batch", type=int, default=32, help="--mode train: items per GPU (config batch_size)")
    ap.add_argument("--mel-loss", action="store_true", help="--mode train: the reference's multi-scale mel loss x 60 instead of the synthetic scalar loss")
    ap.add_argument("--graph-step", action="store_true", help="--mode train, one GPU: capture the whole step in a CUDA graph")
    ap.add_argument("--total-clips", type=int, default=64)
    ap.add_argument("--from-npy", action="store_true", help="--mode vocoder: also run the file-to-file path (.npy mels on tmpfs in, .wav out)")
    ap.add_argument("--cpu-repeats", type=int, default=3)
    ap.add_argument("--chunks", type=int, default=0, help="tuning: 16-byte chunks per thread segment for the forward (0 = library default)")
    ap.add_argument("--table", action="store_true")
    ap.add_argument("--torch-baseline", action="store_true", help="with --table: also time the torch-op path on the GPU")
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = max(args.warmup, 3)      # timing rule: at least 3 untimed warm-up steps
    if args.table:
        return run_table(args)
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "vocoder":
        return run_vocoder_mode(args)
    if args.mode == "train":
        return run_train_mode(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
