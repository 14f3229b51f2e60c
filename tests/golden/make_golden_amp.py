#!/usr/bin/env python
"""Generate tests/golden/amp_golden.npz by running the UNMODIFIED reference generator code (bigvgan.py)
in the build container: AMPBlock1 / AMPBlock2 forwards, the generator tail (activation_post -> conv_post ->
clamp | tanh -> int16 stereo as inference_e2e.py writes it) and one whole (small) generator pass.

    python tests/golden/make_golden_amp.py        # needs /root/reference (read-only)

These vectors pin the rows SURVEY.md section 8(f) ranks 1 and 2: the reference ships no tests for them either.
`bigvgan.py` imports matplotlib / librosa through utils.py / meldataset.py; they are absent here and are
stubbed in sys.modules, exactly as tests/test_vocoder_harness.py does.
Weight norm is removed (as inference_e2e.py:126 does) before the weights are saved, so the vectors
carry plain `weight` / `bias` arrays.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types
import zlib

import numpy as np

REF = os.environ.get("AFA_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
BIG = os.path.join(REF, "BigVGAN")


def load_reference():
    for name in ("matplotlib", "matplotlib.pylab", "librosa", "librosa.filters", "librosa.util"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["librosa.filters"].mel = lambda *a, **k: None
    sys.modules["librosa.util"].normalize = lambda *a, **k: None
    # the reference imports `alias_free_activation.torch.act`, a sub-package it does not ship;
    # alias the flat files under that name, unmodified (same recipe as make_golden.py)
    sys.path.insert(0, HERE)
    import make_golden

    make_golden.load_reference()
    sys.path.insert(0, BIG)
    import bigvgan
    from env import AttrDict

    return bigvgan, AttrDict


def randomise_snake(module, gen, torch):
    with torch.no_grad():
        for n, p in module.named_parameters():
            if n.endswith("alpha") or n.endswith("beta"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.5)


def scale_weights(module, gen, torch, gain):
    """Rescale every convolution to a fan-in-normalised gain, so that the signal neither dies nor explodes through
    the block and an error anywhere would show in the output; biases become non-trivial."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if n.endswith("weight") and p.dim() == 3:
                fan_in = p.shape[1] * p.shape[2] if "ups" not in n else p.shape[0] * p.shape[2] / 4.0
                p.mul_(gain / (float(p.std()) * fan_in ** 0.5))
            if n.endswith("bias"):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)


def main():
    import torch

    bigvgan, AttrDict = load_reference()
    h = AttrDict(json.load(open(os.path.join(BIG, "configs", "bigvgan_binaural_22khz_80band_256x.json"))))
    out = {}
    g = torch.Generator().manual_seed(1234)

    # ---- AMPBlock1 / AMPBlock2 forward                                              bigvgan.py:132-141, 233-236
    for name, cls, C, T, k, dil, act in (
        ("amp1_k3", bigvgan.AMPBlock1, 6, 47, 3, (1, 3, 5), "snakebeta"),
        ("amp1_k7", bigvgan.AMPBlock1, 4, 131, 7, (1, 3, 5), "snakebeta"),
        ("amp1_k11_snake", bigvgan.AMPBlock1, 3, 64, 11, (1, 3, 5), "snake"),
        ("amp2_k3", bigvgan.AMPBlock2, 5, 40, 3, (1, 3), "snakebeta"),
    ):
        torch.manual_seed(zlib.crc32(name.encode()))     # the block's initial weights come from the global RNG: seed it per case
        blk = cls(h, C, k, dil, activation=act)
        blk.remove_weight_norm()
        randomise_snake(blk, g, torch)
        scale_weights(blk, g, torch, 0.8)
        x = torch.randn(2, C, T, generator=g)
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            with torch.no_grad():
                out[f"{name}/y_{tag}"] = blk.to(dt)(x.to(dt)).numpy()
        blk = blk.float()
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/meta"] = np.array([2, C, T, k, len(dil), int(cls is bigvgan.AMPBlock1), int(act == "snakebeta")] + list(dil), dtype=np.int64)
        for n, p in blk.state_dict().items():
            out[f"{name}/sd/{n}"] = p.numpy()

    # ---- generator tail                                             bigvgan.py:379-385, inference_e2e.py:193-201
    for name, C, T, tanh_final, bias_final in (("tail_clamp", 24, 333, False, False), ("tail_tanh_bias", 8, 100, True, True)):
        hh = AttrDict(dict(h))
        hh["upsample_initial_channel"] = C * 64
        hh["use_tanh_at_final"] = tanh_final
        hh["use_bias_at_final"] = bias_final
        torch.manual_seed(7)
        gen = bigvgan.BigVGAN(hh)
        gen.remove_weight_norm()
        randomise_snake(gen.activation_post, g, torch)
        with torch.no_grad():
            gen.conv_post.weight.mul_(0.6 / (float(gen.conv_post.weight.std()) * (7 * C) ** 0.5))   # part of the output reaches the clamp
            if bias_final:
                gen.conv_post.bias.fill_(0.05)
        x = torch.randn(2, C, T, generator=g)        # rows: left, right
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            gen.activation_post.to(dt); gen.conv_post.to(dt)
            with torch.no_grad():
                y = gen.conv_post(gen.activation_post(x.to(dt)))
                y = torch.tanh(y) if gen.use_tanh_at_final else torch.clamp(y, min=-1.0, max=1.0)
            out[f"{name}/wave_{tag}"] = y.numpy()
            if tag == "f32":                         # inference_e2e.py:174-201, verbatim operations
                left, right = y[0:1].squeeze().cpu().numpy(), y[1:2].squeeze().cpu().numpy()
                stereo = np.stack([left, right], axis=0)
                stereo = stereo * 32767.0            # MAX_WAV_VALUE, meldataset.py:20
                stereo = stereo.astype("int16").T
                out[f"{name}/pcm_i16"] = stereo
        gen.activation_post.float(); gen.conv_post.float()
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/meta"] = np.array([2, C, T, int(tanh_final), int(bias_final)], dtype=np.int64)
        for n, p in gen.activation_post.state_dict().items():
            out[f"{name}/sd/activation_post.{n}"] = p.numpy()
        for n, p in gen.conv_post.state_dict().items():
            out[f"{name}/sd/conv_post.{n}"] = p.numpy()

    # ---- one whole (small) generator pass                                                   bigvgan.py:361-387
    for name, resblock in (("gen_small_1", "1"), ("gen_small_2", "2")):
        hh = AttrDict(dict(h))
        hh.update(upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=16, resblock=resblock,
                  resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3 if resblock == "1" else [[1, 3]] * 3)
        torch.manual_seed(11)
        gen = bigvgan.BigVGAN(hh)
        gen.remove_weight_norm()
        randomise_snake(gen, g, torch)
        scale_weights(gen, g, torch, 0.7)
        with torch.no_grad():
            gen.conv_post.weight.mul_(0.25)          # keep most of the waveform out of the clamp
        mel = torch.rand(2, 80, 13, generator=g) * 14.5 - 12.0
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            with torch.no_grad():
                out[f"{name}/y_{tag}"] = gen.to(dt)(mel.to(dt)).numpy()
        gen = gen.float()
        out[f"{name}/mel"] = mel.numpy()
        out[f"{name}/meta"] = np.array([int(resblock)], dtype=np.int64)
        for n, p in gen.state_dict().items():
            out[f"{name}/sd/{n}"] = p.numpy()

    # ---- whole generator with the SHIPPED stage plan (six stages, rates 4,4,2,2,2,2, AMPBlock1 x 3 kernels), narrowed to
    #      upsample_initial_channel = 192 (channels 96 ... 3) and T_mel = 16.  Weights: tests/golden/synth_weights.py
    #      (derived from name + shape on both sides), so only the reference's outputs are stored.     bigvgan.py:244-387
    from synth_weights import synth_state_dict

    hh = AttrDict(dict(h))
    hh.update(upsample_initial_channel=192)
    torch.manual_seed(13)
    gen = bigvgan.BigVGAN(hh)
    gen.remove_weight_norm()
    sd = gen.state_dict()
    new = synth_state_dict({k: tuple(v.shape) for k, v in sd.items()})
    gen.load_state_dict({k: torch.tensor(new[k]) if k in new else v for k, v in sd.items()})
    mel = torch.rand(2, 80, 16, generator=g) * 14.5 - 12.0
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        with torch.no_grad():
            out[f"gen_model_192/y_{tag}"] = gen.to(dt)(mel.to(dt)).numpy()
    out["gen_model_192/mel"] = mel.numpy()
    out["gen_model_192/meta"] = np.array([192, 16, len(sd)], dtype=np.int64)
    out["gen_model_192/param_checksum"] = np.array([float(sum(np.abs(v).sum() for v in new.values()))])

    # ---- zero-frame handling of the inference script                                     inference_e2e.py:38-111
    spec = importlib.util.spec_from_file_location("ref_inference_e2e", os.path.join(BIG, "inference_e2e.py"))
    inf = importlib.util.module_from_spec(spec)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(inf)
        rng = np.random.default_rng(77)
        for name, t_mel, zero_frames in (("zf_some", 40, [0, 1, 7, 8, 9, 23, 39]), ("zf_none", 17, []), ("zf_all_but_one", 9, [0, 1, 2, 3, 5, 6, 7, 8])):
            mel = (rng.random((80, t_mel)) * 14.5 - 12.0).astype(np.float32)
            mel[:, zero_frames] = 0.0
            filt, mask, idx = inf.detect_and_exclude_zero_frames(mel)
            audio = rng.standard_normal(filt.shape[1] * 256).astype(np.float32)
            restored = inf.reconstruct_audio_with_silence(audio, mask, idx, 256, t_mel * 256)
            out[f"{name}/mel"] = mel
            out[f"{name}/filtered"] = np.ascontiguousarray(filt)
            out[f"{name}/zero_mask"] = mask
            out[f"{name}/nonzero_indices"] = np.asarray(idx, dtype=np.int64)
            out[f"{name}/audio"] = audio
            out[f"{name}/restored"] = restored

    path = os.path.join(HERE, "amp_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
