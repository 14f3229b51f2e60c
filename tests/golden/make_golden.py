#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

The reference ships no tests or golden vectors for the Activation1d path, and it
cannot travel to the GPU box, so the pin for parity is: outputs of the reference itself, produced
here once and committed.  Run from the repo root:

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference tree is import-broken as shipped (its flat files import a sub-package
`alias_free_activation.torch` that does not exist).  They are loaded unmodified under
the module names they expect, via sys.modules aliasing.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("AFA_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    big = os.path.join(REF, "BigVGAN")
    for name in ("alias_free_activation", "alias_free_activation.torch"):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    flt = load("alias_free_activation.torch.filter", os.path.join(big, "alias_free_activation", "filter.py"))
    load("alias_free_activation.torch.resample", os.path.join(big, "alias_free_activation", "resample.py"))
    act = load("alias_free_activation.torch.act", os.path.join(big, "alias_free_activation", "act.py"))
    activations = load("activations", os.path.join(big, "activations.py"))
    return flt, act, activations


CASES = [
    # name, B, C, T, kind, logscale, x_scale
    ("t1", 2, 3, 1, "snakebeta", True, 1.0),
    ("t2", 1, 2, 2, "snakebeta", True, 1.0),
    ("t3", 2, 2, 3, "snake", True, 1.0),
    ("t5", 2, 3, 5, "snakebeta", False, 1.0),
    ("t11", 1, 4, 11, "snake", False, 1.0),
    ("t37", 2, 2, 37, "snakebeta", True, 1.0),
    ("t127", 1, 3, 127, "snakebeta", True, 1.0),
    ("t128", 2, 2, 128, "snake", True, 1.0),
    ("t129", 2, 4, 129, "snakebeta", True, 10.0),
    ("t300", 1, 2, 300, "snakebeta", True, 1.0),
    ("t1000_zero_init", 1, 2, 1000, "snakebeta", True, 1.0),
]


def main():
    flt, act_mod, activations = load_reference()
    out = {}
    taps = flt.kaiser_sinc_filter1d(cutoff=0.25, half_width=0.3, kernel_size=12)
    out["taps_f32"] = taps.reshape(-1).numpy()

    g = torch.Generator().manual_seed(1234)          # config seed, configs/bigvgan_binaural_22khz_80band_256x.json:9
    for name, B, C, T, kind, logscale, xs in CASES:
        x = (torch.randn(B, C, T, generator=g) * xs).float()
        gy = torch.randn(B, C, T, generator=g).float()
        cls = activations.SnakeBeta if kind == "snakebeta" else activations.Snake
        res = {}
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            a = cls(C, alpha_logscale=logscale)
            if "zero_init" not in name:
                gp = torch.Generator().manual_seed(len(name) * 131 + T)
                with torch.no_grad():
                    if logscale:
                        a.alpha.copy_(torch.randn(C, generator=gp) * 0.5)
                        if kind == "snakebeta":
                            a.beta.copy_(torch.randn(C, generator=gp) * 0.5)
                    else:
                        a.alpha.copy_(torch.rand(C, generator=gp) * 2.0 + 0.25)
                        if kind == "snakebeta":
                            a.beta.copy_(torch.rand(C, generator=gp) * 2.0 + 0.25)
            res["alpha"] = a.alpha.detach().numpy().copy()
            if kind == "snakebeta":
                res["beta"] = a.beta.detach().numpy().copy()
            m = act_mod.Activation1d(activation=a).to(dt)
            xi = x.to(dt).clone().requires_grad_(True)
            y = m(xi)
            y.backward(gy.to(dt))
            res["y_" + tag] = y.detach().numpy()
            res["gx_" + tag] = xi.grad.numpy()
            res["galpha_" + tag] = m.act.alpha.grad.numpy()
            if kind == "snakebeta":
                res["gbeta_" + tag] = m.act.beta.grad.numpy()
            if tag == "f32":
                assert sorted(m.state_dict().keys()) == sorted(
                    ["act.alpha", "upsample.filter", "downsample.lowpass.filter"]
                    + (["act.beta"] if kind == "snakebeta" else [])
                )
        res["x"] = x.numpy()
        res["gy"] = gy.numpy()
        res["meta"] = np.array([B, C, T, int(kind == "snakebeta"), int(logscale)], dtype=np.int64)
        for k, v in res.items():
            out[f"{name}/{k}"] = v

    # composite up->down impulse response with the activation's periodic term switched off
    # (SURVEY.md Appendix B): alpha -> tiny so sin^2 vanishes.
    a = activations.SnakeBeta(1, alpha_logscale=False)
    with torch.no_grad():
        a.alpha.fill_(0.0)
        a.beta.fill_(1.0)
    m = act_mod.Activation1d(activation=a).double()
    imp = torch.zeros(1, 1, 41, dtype=torch.float64)
    imp[0, 0, 20] = 1.0
    out["impulse_response_f64"] = m(imp).detach().numpy().reshape(-1)

    path = os.path.join(HERE, "activation1d_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
