"""CPU-side tests: the C ABI loads and exports what include/afa_b200.h declares, argument checking,
the nn.Module mirror's checkpoint contract, the drop-in seam against the unmodified reference
(when /root/reference is mounted), and clip sharding over a 2-rank gloo group."""
import ctypes
import os
import re
import socket
import sys
import types

import numpy as np
import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")
REFERENCE = os.environ.get("AFA_REFERENCE_ROOT", "/root/reference")


@pytest.fixture(scope="module")
def lib():
    from afa_b200.build import build_library
    from afa_b200._lib import load_library

    build_library()
    return load_library()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(REPO, "include", "afa_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(afa_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    from afa_b200._lib import EXPORTED_SYMBOLS

    assert sorted(EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.afa_version() == 152


def test_argument_errors_need_no_gpu(lib):
    F12 = ctypes.c_float * 12
    taps = F12(*([0.0] * 12))
    vp = ctypes.c_void_p
    assert lib.afa_activation1d_fwd(None, None, None, None, taps, taps, 1, 1, 8, 0, 0, None) == -1
    assert b"null" in lib.afa_last_error()
    assert lib.afa_activation1d_fwd(vp(16), vp(32), vp(64), vp(64), taps, taps, 1, 1, 8, 5, 0, None) == -2
    assert lib.afa_activation1d_fwd(vp(16), vp(32), vp(64), None, taps, taps, 1, 1, 8, 0, 0, None) == -1   # beta required
    assert lib.afa_activation1d_fwd(vp(16), vp(16), vp(64), vp(64), taps, taps, 1, 1, 8, 0, 0, None) == -1  # aliasing
    assert lib.afa_activation1d_fwd(vp(16), vp(32), vp(64), vp(64), taps, taps, -1, 1, 8, 0, 0, None) == -1
    assert lib.afa_activation1d_fwd(vp(18), vp(32), vp(64), vp(64), taps, taps, 1, 1, 8, 0, 0, None) == -5  # 2-byte aligned fp32
    assert lib.afa_activation1d_fwd(vp(16), vp(32), vp(64), vp(64), taps, taps, 1 << 20, 1 << 12, 1 << 20, 0, 0, None) == -3
    assert lib.afa_set_tuning(0, 4, 0) == -1 and lib.afa_set_tuning(10, 0, 0) == -1 and lib.afa_set_tuning(9, 2, 0) == -1 and lib.afa_set_tuning(9, 1, 0) == 0 and lib.afa_set_tuning(3, 2, 0) == -1 and lib.afa_set_tuning(5, 3, 0) == -1
    assert lib.afa_set_tuning(5, 1, 7) == -1 and lib.afa_set_tuning(6, 8, 0) == -1 and lib.afa_set_tuning(5, 1, 0) == 0 and lib.afa_set_tuning(6, -1, 0) == 0
    assert lib.afa_set_tuning(0, 9, 0) == 0 and lib.afa_set_tuning(0, 0, 0) == 0
    assert lib.afa_set_tuning(2, 4, 0) == 0 and lib.afa_set_tuning(2, 0, 0) == 0 and lib.afa_set_tuning(2, -1, 0) == -1
    # channels-last tensor-core forward (7: mode 0..3, forced blocks per CTA a multiple of 4) and the tail's walk length (8)
    assert lib.afa_set_tuning(7, 4, 0) == -1 and lib.afa_set_tuning(7, 1, 6) == -1 and lib.afa_set_tuning(7, 2, 8) == 0 and lib.afa_set_tuning(7, 1, 0) == 0
    assert lib.afa_set_tuning(8, -1, 0) == -1 and lib.afa_set_tuning(8, 3, 0) == 0 and lib.afa_set_tuning(8, 0, 0) == 0
    # channels-last AMP entry points: argument checks come before any CUDA call
    i64 = ctypes.c_int64
    cl = lib.afa_amp_activation1d_fwd_cl
    assert cl(None, 0, None, 0, None, None, 0, None, 0, 0, None, None, taps, taps, 1, 4, 8, 0, 0, None) == -1
    assert cl(vp(16), 32, None, 0, None, vp(48), 32, vp(32), 32, 0, vp(64), vp(64), taps, taps, 1, 4, 8, 0, 0, None) == -1   # xsum without res (they come together)
    assert cl(vp(16), 32, None, 0, None, None, 0, vp(16), 32, 0, vp(64), vp(64), taps, taps, 1, 4, 8, 0, 0, None) == -1    # y aliases x
    assert cl(vp(16), 31, None, 0, None, None, 0, vp(32), 32, 0, vp(64), vp(64), taps, taps, 2, 4, 8, 0, 0, None) == -1    # batch stride too small
    assert cl(vp(16), 32, None, 0, None, None, 0, vp(32), 32, 7, vp(64), vp(64), taps, taps, 1, 4, 8, 0, 0, None) == -1    # y_tpad < T
    assert cl(vp(16), 32, None, 0, None, None, 0, vp(32), 32, 0, vp(64), vp(64), taps, taps, 1, 4, 8, 3, 0, None) == -2
    tail = lib.afa_tail_fwd_cl
    assert tail(vp(16), 320, vp(64), vp(64), taps, taps, vp(64), None, 0, None, None, 2, 32767.0, None, 0, 0, 2, 40, 8, 0, 0, None) == -1   # no output
    assert tail(vp(16), 320, vp(64), vp(64), taps, taps, vp(64), None, 0, vp(128), None, 2, 32767.0, None, 0, 0, 2, 40, 8, 0, 0, None) == -1  # channels > 32
    assert tail(vp(16), 64, vp(64), vp(64), taps, taps, vp(64), None, 0, None, vp(128), 2, 32767.0, None, 0, 0, 3, 8, 8, 0, 0, None) == -1    # batch % interleave
    assert tail(vp(16), 64, vp(64), vp(64), taps, taps, vp(64), None, 0, vp(128), None, 2, 32767.0, vp(256), 3, 8, 2, 8, 8, 0, 0, None) == -1   # T % hop
    arr = (vp * 2)(vp(16), vp(32))
    assert lib.afa_resblock_mean(arr, arr, 5, None, 1.0, vp(64), 4, 4, 0, None) == -1
    assert lib.afa_resblock_mean(arr, arr, 2, None, 1.0, vp(64), 0, 4, 0, None) == 0                                               # empty: no-op
    ws = lib.afa_bwd_workspace_bytes(2, 3, 1000, 0)
    assert ws >= 2 * 4 * 2 * 3 and lib.afa_bwd_workspace_bytes(4, 3, 1000, 0) > ws
    assert lib.afa_bwd_workspace_bytes(2, 3, 0, 0) <= 16
    # the bound covers the shortest compiled segment (5 chunks of 4 fp32) plus the slice sums and counters of the split finalize
    assert ws >= 2 * 3 * 50 * 2 * 4 + 3 * (32 * 2 * 4 + 4)


def test_module_checkpoint_contract(golden):
    from afa_b200 import Activation1d
    from afa_b200.activations import Snake, SnakeBeta

    m = Activation1d(activation=SnakeBeta(7, alpha_logscale=True))
    sd = m.state_dict()
    assert sorted(sd) == ["act.alpha", "act.beta", "downsample.lowpass.filter", "upsample.filter"]
    assert sd["act.alpha"].shape == (7,) and sd["upsample.filter"].shape == (1, 1, 12)
    assert sd["downsample.lowpass.filter"].shape == (1, 1, 12)
    np.testing.assert_array_equal(sd["upsample.filter"].reshape(-1).numpy(), golden["taps_f32"])
    np.testing.assert_array_equal(sd["downsample.lowpass.filter"].reshape(-1).numpy(), golden["taps_f32"])
    assert torch.all(sd["act.alpha"] == 0) and torch.all(sd["act.beta"] == 0)       # logscale init (activations.py:101-103)
    assert torch.all(Snake(3).alpha == 1)
    assert sorted(Activation1d(activation=Snake(3)).state_dict()) == ["act.alpha", "downsample.lowpass.filter", "upsample.filter"]
    assert (m.up_ratio, m.down_ratio, m.upsample.pad, m.upsample.pad_left, m.upsample.pad_right) == (2, 2, 5, 15, 15)
    assert (m.downsample.lowpass.pad_left, m.downsample.lowpass.pad_right) == (5, 6)
    # load a state dict shaped like the reference's
    other = {k: torch.randn_like(v) for k, v in sd.items()}
    m.load_state_dict(other)
    assert torch.equal(m.act.beta, other["act.beta"])
    assert [n for n, p in m.named_parameters()] == ["act.alpha", "act.beta"]
    assert all(not b.requires_grad for b in m.buffers())


def test_module_refuses_everything_but_the_fused_path():
    from afa_b200 import Activation1d, LowPassFilter1d, UpSample1d
    from afa_b200.activations import SnakeBeta

    act = SnakeBeta(4)
    with pytest.raises(NotImplementedError):
        Activation1d(activation=act, fused=False)
    with pytest.raises(NotImplementedError):
        Activation1d(activation=act, up_ratio=4)
    with pytest.raises(NotImplementedError):
        Activation1d(activation=act, up_kernel_size=8)
    with pytest.raises(TypeError):
        Activation1d(activation=torch.nn.Identity())
    with pytest.raises(ValueError):
        LowPassFilter1d(cutoff=0.6)
    m = Activation1d(activation=act)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 4, 16))                       # no CPU fallback
    with pytest.raises(ValueError):
        m(torch.randn(4, 16))                          # same failure mode as resample.py:30
    with pytest.raises(RuntimeError):
        UpSample1d()(torch.randn(1, 4, 16))
    with pytest.raises(RuntimeError):
        act(torch.randn(1, 4, 16))


def test_bench_workload_matches_survey_counts():
    sys.path.insert(0, REPO)
    import bench

    assert sum(c for _, _, c in bench.AMP_STAGES) == 109                      # bigvgan.py: 18 blocks x 6 + activation_post
    assert bench.step_elements(1, 1) == 2 * 614400                            # SURVEY.md section 8a, per mel frame, L+R
    assert bench.step_elements(1, 861) == 1057996800
    assert [(c, m) for c, m, _ in bench.AMP_STAGES] == [(768, 4), (384, 16), (192, 32), (96, 64), (48, 128), (24, 256)]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "BigVGAN")), reason="reference tree not mounted")
def test_drop_in_behind_use_cuda_kernel_of_the_unmodified_reference():
    """bigvgan.py:94-102/194-202/272-280: with our directory ahead of BigVGAN/ on sys.path, the reference's
    own `use_cuda_kernel=True` switch builds a generator whose 109 Activation1d modules are ours and whose
    state dict is interchangeable with the torch-path generator's."""
    import json
    import subprocess

    script = r'''
import sys, types, json, os
pkg_root, big = sys.argv[1], sys.argv[2]
for name in ("matplotlib", "matplotlib.pylab", "librosa", "librosa.filters", "librosa.util"):
    m = types.ModuleType(name); m.__path__ = []
    sys.modules[name] = m
sys.modules["matplotlib"].use = lambda *a, **k: None
sys.modules["librosa.filters"].mel = lambda *a, **k: None
sys.modules["librosa.util"].normalize = lambda *a, **k: None
sys.path.insert(0, big)
sys.path.insert(0, pkg_root)
import torch
import alias_free_activation.torch.act as ref_act          # aliased reference files (F1 in SURVEY.md)
assert ref_act.__file__.startswith(big), ref_act.__file__
import bigvgan
from env import AttrDict
h = AttrDict(json.load(open(os.path.join(big, "configs", "bigvgan_binaural_22khz_80band_256x.json"))))
h["upsample_initial_channel"] = 96      # same topology, 16x narrower: keeps this CPU test fast
torch.manual_seed(1234)
g_ref = bigvgan.BigVGAN(h)
g_new = bigvgan.BigVGAN(h, use_cuda_kernel=True)
import afa_b200
ours = [m for m in g_new.modules() if isinstance(m, afa_b200.Activation1d)]
theirs = [m for m in g_ref.modules() if isinstance(m, ref_act.Activation1d)]
assert len(ours) == 109 and len(theirs) == 109, (len(ours), len(theirs))
assert not any(isinstance(m, ref_act.Activation1d) for m in g_new.modules())
sd_ref, sd_new = g_ref.state_dict(), g_new.state_dict()
assert list(sd_ref.keys()) == list(sd_new.keys())
assert all(sd_ref[k].shape == sd_new[k].shape and sd_ref[k].dtype == sd_new[k].dtype for k in sd_ref)
for k in sd_ref:
    if k.endswith("filter"):
        assert torch.equal(sd_ref[k], sd_new[k]), k
g_new.load_state_dict(sd_ref)
g_new.remove_weight_norm()
assert g_new.h["use_cuda_kernel"] is True
try:
    g_new(torch.zeros(1, 80, 8))
except RuntimeError as e:
    assert "CUDA" in str(e), e           # CPU tensors: loud failure, never a silent torch fallback
else:
    raise AssertionError("fused generator ran on CPU tensors")
print("DROPIN_OK", len(sd_ref))
'''
    out = subprocess.run([sys.executable, "-c", script, PKG_ROOT, os.path.join(REFERENCE, "BigVGAN")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert "DROPIN_OK" in out.stdout


def test_shard_indices_partition():
    from afa_b200 import shard_indices

    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 4, 8):
            parts = [shard_indices(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gather_worker(rank, world, port, n_items, out_dir):
    import torch.distributed as dist

    sys.path.insert(0, PKG_ROOT)
    from afa_b200 import gather_waveforms, shard_indices

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    mine = shard_indices(n_items, rank, world)
    # each "clip" i yields a deterministic stereo waveform
    local = torch.stack([torch.full((2, 5), float(i)) + torch.arange(5.0) for i in mine]) if mine else torch.zeros(0, 2, 5)
    full = gather_waveforms(local, n_items, rank, world)
    torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
    # interleaved int16 stereo PCM [n, T, 2] (what the channels-last engine's tail writes) travels as int32 pairs
    pcm = torch.stack([(torch.arange(12).reshape(6, 2) * (i + 1) - 7 * i).to(torch.int16) for i in mine]) if mine \
        else torch.zeros(0, 6, 2, dtype=torch.int16)
    torch.save(gather_waveforms(pcm, n_items, rank, world), os.path.join(out_dir, f"p{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [7, 8])
def test_clip_sharding_gather_two_ranks_gloo(tmp_path, n_items):
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    mp.spawn(_gather_worker, args=(world, port, n_items, str(tmp_path)), nprocs=world, join=True)
    expect = torch.stack([torch.full((2, 5), float(i)) + torch.arange(5.0) for i in range(n_items)])
    expect_pcm = torch.stack([(torch.arange(12).reshape(6, 2) * (i + 1) - 7 * i).to(torch.int16) for i in range(n_items)])
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert torch.equal(got, expect)
        got_pcm = torch.load(os.path.join(str(tmp_path), f"p{r}.pt"))
        assert got_pcm.dtype == torch.int16 and torch.equal(got_pcm, expect_pcm)


def test_filter_tap_cache_follows_reloads_not_rebroadcasts():
    """The host copy of the 12 filter taps is re-read when the buffers are reloaded or replaced (`load_state_dict`, `.to()`,
    `refresh_filters()`), and NOT when the same values are written in place -- what DistributedDataParallel's
    broadcast_buffers=True (the default the reference's trainer uses, train_binaural_mel.py:541) does before every forward:
    each re-read would be a synchronising device-to-host copy per module per step."""
    from afa_b200 import Activation1d
    from afa_b200.activations import SnakeBeta

    m = Activation1d(activation=SnakeBeta(4, alpha_logscale=True))
    t0 = m._host_taps()
    first = list(t0[0])
    with torch.no_grad():
        m.upsample.filter.copy_(m.upsample.filter.clone())          # in-place rewrite, same values (bumps the version counter)
    assert m._host_taps() is t0
    sd = m.state_dict()
    sd["upsample.filter"] = sd["upsample.filter"] * 0.5
    m.load_state_dict(sd)
    t1 = m._host_taps()
    assert t1 is not t0 and np.allclose(list(t1[0]), np.array(first) * 0.5)
    with torch.no_grad():
        m.downsample.lowpass.filter.mul_(2.0)
    assert m._host_taps() is t1                                      # in-place edit with new values: the caller must say so
    m.refresh_filters()
    t2 = m._host_taps()
    assert np.allclose(list(t2[1]), np.array(list(t1[1])) * 2.0)
    m.double()
    assert m._host_taps() is not t2                                  # new storage
