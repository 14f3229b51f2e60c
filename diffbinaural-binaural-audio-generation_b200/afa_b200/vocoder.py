"""Measurement harness for the second half of the headline metric: vocoded audio-seconds per second.

`BigVGANGenerator` is the CALLER of the hot path, stated as compactly as possible so that the fused
Activation1d can be timed in its real context on a box that does not have the reference tree
(reference: BigVGAN/bigvgan.py:244-387 generator, :31-147 AMPBlock1, :150-241 AMPBlock2).  Everything
that is not the anti-aliased activation stays library code: `torch.nn.Conv1d` / `ConvTranspose1d`
(cuDNN).  Parameter and buffer names equal the reference's after `remove_weight_norm()`, and
`load_reference_state_dict` folds weight-norm checkpoints (`weight_g`/`weight_v` or
`parametrizations.weight.original0/1`), so a reference checkpoint loads unchanged.

This is harness code for SURVEY.md section 8(d) "audio-sec/sec" and section 8(f) rank 2; it adds no kernels.
The anti-aliased activation is always the fused CUDA op; `activation_factory` exists so that tests can
build the same module tree around their oracle (nothing in this package imports the oracle).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .activations import Snake, SnakeBeta
from .modules import Activation1d

# configs/bigvgan_binaural_22khz_80band_256x.json (the shipped binaural vocoder config), generator part
BINAURAL_22KHZ_80BAND_256X = {
    "resblock": "1",
    "upsample_rates": [4, 4, 2, 2, 2, 2],
    "upsample_kernel_sizes": [8, 8, 4, 4, 4, 4],
    "upsample_initial_channel": 1536,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    "use_tanh_at_final": False,
    "use_bias_at_final": False,
    "activation": "snakebeta",
    "snake_logscale": True,
    "num_mels": 80,
    "hop_size": 256,
    "sampling_rate": 22050,
}


def _pad(kernel_size: int, dilation: int = 1) -> int:
    return (kernel_size * dilation - dilation) // 2          # utils.py:79-80


def _make_act(channels: int, h: Dict, factory):
    cls = {"snake": Snake, "snakebeta": SnakeBeta}.get(h["activation"])
    if cls is None:
        raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
    return factory(cls(channels, alpha_logscale=h["snake_logscale"]))


def _fused(activation):
    return Activation1d(activation=activation)


class AMPBlock1(nn.Module):
    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5), factory=_fused):
        super().__init__()
        self.convs1 = nn.ModuleList(
            [nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=_pad(kernel_size, d)) for d in dilation])
        self.convs2 = nn.ModuleList(
            [nn.Conv1d(channels, channels, kernel_size, 1, dilation=1, padding=_pad(kernel_size, 1)) for _ in dilation])
        self.activations = nn.ModuleList([_make_act(channels, h, factory) for _ in range(2 * len(dilation))])

    def forward(self, x):
        for c1, c2, a1, a2 in zip(self.convs1, self.convs2, self.activations[::2], self.activations[1::2]):
            xt = c2(a2(c1(a1(x))))
            x = xt + x
        return x


class AMPBlock2(nn.Module):
    def __init__(self, h, channels, kernel_size=3, dilation=(1, 3, 5), factory=_fused):
        super().__init__()
        self.convs = nn.ModuleList(
            [nn.Conv1d(channels, channels, kernel_size, 1, dilation=d, padding=_pad(kernel_size, d)) for d in dilation])
        self.activations = nn.ModuleList([_make_act(channels, h, factory) for _ in range(len(dilation))])

    def forward(self, x):
        for c, a in zip(self.convs, self.activations):
            x = c(a(x)) + x
        return x


class BigVGANGenerator(nn.Module):
    """mel [B, num_mels, T_mel] -> waveform [B, 1, T_mel * prod(upsample_rates)], weight norm already folded."""

    def __init__(self, h: Dict | None = None, activation_factory=_fused):
        super().__init__()
        h = dict(BINAURAL_22KHZ_80BAND_256X if h is None else h)
        self.h = h
        c0 = h["upsample_initial_channel"]
        self.num_kernels = len(h["resblock_kernel_sizes"])
        self.num_upsamples = len(h["upsample_rates"])
        self.conv_pre = nn.Conv1d(h["num_mels"], c0, 7, 1, padding=3)
        block = {"1": AMPBlock1, "2": AMPBlock2}[h["resblock"]]
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch = c0
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            self.ups.append(nn.ModuleList([nn.ConvTranspose1d(c0 // (2 ** i), c0 // (2 ** (i + 1)), k, u, padding=(k - u) // 2)]))
            ch = c0 // (2 ** (i + 1))
            for ks, d in zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]):
                self.resblocks.append(block(h, ch, ks, tuple(d), factory=activation_factory))
        self.activation_post = _make_act(ch, h, activation_factory)
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3, bias=h.get("use_bias_at_final", True))
        self.use_tanh_at_final = h.get("use_tanh_at_final", True)
        for m in self.modules():                                  # utils.py:67-70 init_weights
            if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d)):
                m.weight.data.normal_(0.0, 0.01)

    @property
    def hop(self) -> int:
        out = 1
        for u in self.h["upsample_rates"]:
            out *= u
        return out

    def forward(self, x):                                         # bigvgan.py:361-387
        x = self.conv_pre(x)
        for i in range(self.num_upsamples):
            for up in self.ups[i]:
                x = up(x)
            xs = None
            for j in range(self.num_kernels):
                y = self.resblocks[i * self.num_kernels + j](x)
                xs = y if xs is None else xs + y
            x = xs / self.num_kernels
        x = self.activation_post(x)
        x = self.conv_post(x)
        return torch.tanh(x) if self.use_tanh_at_final else torch.clamp(x, min=-1.0, max=1.0)

    def load_reference_state_dict(self, sd: Dict[str, torch.Tensor]):
        """Load a reference checkpoint's `generator` dict, folding weight norm if it is still there."""
        folded: Dict[str, torch.Tensor] = {}
        for k, v in sd.items():
            if k.endswith(".weight_v") or k.endswith(".parametrizations.weight.original1"):
                base = k[: -len(".weight_v")] if k.endswith(".weight_v") else k[: -len(".parametrizations.weight.original1")]
                g = sd[base + ".weight_g"] if k.endswith(".weight_v") else sd[base + ".parametrizations.weight.original0"]
                norm = v.flatten(1).norm(dim=1).view(-1, *([1] * (v.dim() - 1)))
                folded[base + ".weight"] = v * (g / norm)
            elif k.endswith(".weight_g") or k.endswith(".parametrizations.weight.original0"):
                continue
            else:
                folded[k] = v
        return self.load_state_dict(folded)


class GraphedVocoder:
    """CUDA-graph replay of the generator for a fixed mel shape: no per-call launch overhead
    (109 fused-activation launches + ~150 cuDNN launches per pass are captured once)."""

    def __init__(self, gen: BigVGANGenerator, batch: int, t_mel: int, dtype=torch.bfloat16, device="cuda"):
        self.gen = gen
        self.static_in = torch.zeros(batch, gen.h["num_mels"], t_mel, device=device, dtype=dtype)
        with torch.no_grad():
            side = torch.cuda.Stream(device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):
                for _ in range(2):                               # warm-up: cuDNN algorithm picks, tap caches
                    gen(self.static_in)
            torch.cuda.current_stream(device).wait_stream(side)
            torch.cuda.synchronize(device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = gen(self.static_in)

    def __call__(self, mel: torch.Tensor) -> torch.Tensor:
        self.static_in.copy_(mel, non_blocking=True)
        self.graph.replay()
        return self.static_out
