"""Deterministic synthetic weights for whole-generator golden vectors.

A generator with the shipped stage plan has ~2 M parameters even at upsample_initial_channel = 192: too large for a
committed fixture.  Both sides -- tests/golden/make_golden_amp.py (loads them into the UNMODIFIED reference
bigvgan.BigVGAN) and the GPU tests (load them into this repository's generators) -- therefore derive the same
state dict from (name, shape) with numpy's PCG64, and only the reference's OUTPUTS are committed.
"""
from __future__ import annotations

import zlib

import numpy as np


def synth_state_dict(shapes: dict, seed: int = 2024, gain: float = 0.7) -> dict:
    """shapes: {state-dict key: shape} of a generator AFTER remove_weight_norm().  Keys ending in `filter` (the
    Kaiser-sinc buffers of Activation1d) are not produced: they keep the values the modules computed."""
    out = {}
    for name in sorted(shapes):
        shape = tuple(int(v) for v in shapes[name])
        if name.endswith("filter"):
            continue
        rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
        if name.endswith("alpha") or name.endswith("beta"):
            v = rng.standard_normal(shape) * 0.5
        elif name.endswith("bias"):
            v = rng.standard_normal(shape) * 0.3
        elif name.endswith("weight") and len(shape) == 3:
            # Conv1d [out, in, k]: fan-in = in * k.  ConvTranspose1d [in, out, k] (the upsamplers): in * k / stride ~ in * k / 2
            fan_in = shape[1] * shape[2] if ".ups." not in "." + name else shape[0] * shape[2] / 2.0
            v = rng.standard_normal(shape) * (gain / fan_in ** 0.5)
            if name.startswith("conv_post"):
                v *= 0.25                      # keep most of the waveform out of the final clamp
        else:
            raise ValueError(f"unexpected state-dict entry {name} {shape}")
        out[name] = v.astype(np.float32)
    return out
