"""The generator harness (afa_b200/vocoder.py): structure parity with the unmodified reference generator on CPU
(convolutions + the torch-op oracle activation), and fused-vs-oracle parity of the whole generator on the GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.nn as nn

from oracle import torch_path as TP

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")
REFERENCE = os.environ.get("AFA_REFERENCE_ROOT", "/root/reference")


class OracleActivation1d(nn.Module):
    """Test-side stand-in with the same module tree: the activation is evaluated by the torch-op oracle."""

    def __init__(self, activation):
        super().__init__()
        from afa_b200.modules import DownSample1d, UpSample1d

        self.act = activation
        self.upsample = UpSample1d(2, 12)
        self.downsample = DownSample1d(2, 12)

    def forward(self, x):
        beta = getattr(self.act, "beta", None)
        dt = x.dtype
        return TP.activation1d_torch(x, self.act.alpha.to(dt), None if beta is None else beta.to(dt),
                                     bool(self.act.alpha_logscale), self.upsample.filter.to(dt),
                                     self.downsample.lowpass.filter.to(dt))


def small_config(resblock="1", activation="snakebeta"):
    from afa_b200.vocoder import BINAURAL_22KHZ_80BAND_256X

    h = dict(BINAURAL_22KHZ_80BAND_256X)
    h.update(upsample_initial_channel=64, resblock=resblock, activation=activation)
    return h


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "BigVGAN")), reason="reference tree not mounted")
def test_harness_matches_reference_generator_on_cpu(tmp_path):
    """Same checkpoint (with weight norm), same mel -> same waveform as the reference's own generator (torch path)."""
    script = r'''
import sys, types, json, os
pkg_root, big, repo, out = sys.argv[1:5]
for name in ("matplotlib", "matplotlib.pylab", "librosa", "librosa.filters", "librosa.util"):
    m = types.ModuleType(name); m.__path__ = []
    sys.modules[name] = m
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["librosa.filters"].mel = lambda *a, **k: None
sys.modules["librosa.util"].normalize = lambda *a, **k: None
sys.path.insert(0, big); sys.path.insert(0, pkg_root); sys.path.insert(0, repo)
import torch
import bigvgan
from env import AttrDict
h = AttrDict(json.load(open(os.path.join(big, "configs", "bigvgan_binaural_22khz_80band_256x.json"))))
h["upsample_initial_channel"] = 64
torch.manual_seed(1234)
g = bigvgan.BigVGAN(h)                      # torch path, weight norm on
with torch.no_grad():
    for n, p in g.named_parameters():
        if n.endswith("alpha") or n.endswith("beta"):
            p.normal_(0, 0.5)
mel = torch.rand(2, 80, 12) * 14.5 - 12.0
with torch.no_grad():
    y = g(mel)
torch.save({"sd": g.state_dict(), "mel": mel, "y": y}, out)
'''
    out = os.path.join(str(tmp_path), "ref.pt")
    r = subprocess.run([sys.executable, "-c", script, PKG_ROOT, os.path.join(REFERENCE, "BigVGAN"), REPO, out],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    blob = torch.load(out)
    from afa_b200.vocoder import BigVGANGenerator

    mine = BigVGANGenerator(small_config(), activation_factory=OracleActivation1d)
    mine.load_reference_state_dict(blob["sd"])
    with torch.no_grad():
        y = mine(blob["mel"])
    assert y.shape == blob["y"].shape == (2, 1, 12 * 256)
    assert (y - blob["y"]).abs().max().item() <= 1e-6 * max(1.0, blob["y"].abs().max().item())
    # names after weight-norm removal are the reference's
    assert "conv_pre.weight" in mine.state_dict() and "resblocks.0.activations.0.act.alpha" in mine.state_dict()
    assert "activation_post.downsample.lowpass.filter" in mine.state_dict()


@pytest.fixture
def true_fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.gpu
@pytest.mark.parametrize("resblock,activation", [("1", "snakebeta"), ("2", "snake")])
def test_generator_fused_vs_oracle_on_gpu(resblock, activation, true_fp32_convs):
    from afa_b200.vocoder import BigVGANGenerator, GraphedVocoder

    dev = torch.device("cuda:0")
    h = small_config(resblock, activation)
    torch.manual_seed(1234)
    # cuDNN's default TF32 convolutions round the (1e-7-different) activations to 10 mantissa bits, which turns
    # rounding-boundary crossings into 1e-3 differences; compare in true fp32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fused = BigVGANGenerator(h).to(dev)
    with torch.no_grad():
        for n, p in fused.named_parameters():
            if n.endswith("alpha") or n.endswith("beta"):
                p.normal_(0, 0.5)
            elif n.endswith("weight"):
                p.mul_(8.0)                                     # make the signal non-trivial through 6 stages
    oracle = BigVGANGenerator(h, activation_factory=OracleActivation1d).to(dev)
    oracle.load_state_dict(fused.state_dict())
    mel = (torch.rand(2, 80, 40, device=dev) * 14.5 - 12.0)
    with torch.no_grad():
        y_f = fused(mel)
        y_o = oracle(mel)
    assert y_f.shape == (2, 1, 40 * 256)
    scale = max(y_o.abs().max().item(), 1e-6)
    err = (y_f - y_o).abs().max().item() / scale
    assert err <= 2e-4, err
    # CUDA-graph replay of the whole generator reproduces the eager result bit for bit
    gv = GraphedVocoder(fused, 2, 40, dtype=torch.float32, device=dev)
    y_g = gv(mel).clone()
    assert torch.equal(y_g, y_f)
    # bf16 generator (cfg 3: generator.bfloat16()) stays close to the fp32 one
    with torch.no_grad():
        y_b = fused.bfloat16()(mel.bfloat16()).float()
    assert (y_b - y_o).abs().max().item() / scale <= 0.1
