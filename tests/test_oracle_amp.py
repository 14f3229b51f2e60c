"""The AMP-block oracle (oracle/amp_oracle.py) against the vectors the unmodified reference generator code produced
(tests/golden/amp_golden.npz <- tests/golden/make_golden_amp.py): AMPBlock1/2 forwards, the tail incl. the int16
stereo PCM of inference_e2e.py, a whole small generator pass -- and the decomposition the channels-last kernels
implement (bias / residual / mean folded into the activation) against the same vectors."""
import os

import numpy as np
import pytest

from oracle import afa_oracle as O
from oracle import amp_oracle as A

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def amp_golden():
    g = dict(np.load(os.path.join(REPO, "tests", "golden", "amp_golden.npz")))
    cases = {}
    for k, v in g.items():
        name, rest = k.split("/", 1)
        c = cases.setdefault(name, {"sd": {}})
        if rest.startswith("sd/"):
            c["sd"][rest[3:]] = v
        else:
            c[rest] = v
    return cases


def _sd64(sd):
    return {k: v.astype(np.float64) for k, v in sd.items()}


SMALL_H = dict(upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], resblock_kernel_sizes=[3, 7, 11],
               use_tanh_at_final=False)


def test_ampblocks_match_reference(amp_golden):
    for name in ("amp1_k3", "amp1_k7", "amp1_k11_snake", "amp2_k3"):
        c = amp_golden[name]
        B, C, T, k, nd, is1, is_beta = [int(v) for v in c["meta"][:7]]
        dil = [int(v) for v in c["meta"][7:7 + nd]]
        fn = A.ampblock1_forward if is1 else A.ampblock2_forward
        y = fn(_sd64(c["sd"]), "", c["x"].astype(np.float64), k, dil)
        assert O.max_normalised_error(y, c["y_f64"]) <= 1e-12, name
        assert O.max_normalised_error(y, c["y_f32"]) <= 1e-5, name


def test_decomposition_matches_reference(amp_golden):
    """bias / residual folded into the activation (what the engine schedules) == the reference block."""
    for name in ("amp1_k3", "amp1_k7", "amp1_k11_snake", "amp2_k3"):
        c = amp_golden[name]
        B, C, T, k, nd, is1, is_beta = [int(v) for v in c["meta"][:7]]
        dil = [int(v) for v in c["meta"][7:7 + nd]]
        sd = _sd64(c["sd"])
        t, tb, r, rb = A.ampblock_decomposed(sd, "", c["x"].astype(np.float64), k, dil, bool(is1))
        y = A.resblock_mean([t], [r], tb + rb, 1.0).transpose(0, 2, 1)
        assert O.max_normalised_error(y, c["y_f64"]) <= 1e-12, name


def test_tail_and_pcm_match_reference(amp_golden):
    for name in ("tail_clamp", "tail_tanh_bias"):
        c = amp_golden[name]
        B, C, T, use_tanh, has_bias = [int(v) for v in c["meta"]]
        sd = _sd64(c["sd"])
        y = A.tail_forward(sd, c["x"].astype(np.float64), bool(use_tanh))
        assert O.max_normalised_error(y, c["wave_f64"]) <= 1e-12, name
        # channels-last formulation
        w = A.tail_cl(c["x"].transpose(0, 2, 1), sd["activation_post.act.alpha"], sd.get("activation_post.act.beta"), True,
                      sd["conv_post.weight"].reshape(C, 7), sd.get("conv_post.bias"), bool(use_tanh),
                      sd["activation_post.upsample.filter"].reshape(-1), sd["activation_post.downsample.lowpass.filter"].reshape(-1))
        assert O.max_normalised_error(w, c["wave_f64"][:, 0, :]) <= 1e-12, name
        # int16 stereo exactly as inference_e2e.py writes it, from the reference's own fp32 waveform: bit-exact
        pcm = A.pcm_stereo(c["wave_f32"][:, 0, :])
        assert pcm.dtype == np.int16 and np.array_equal(pcm, c["pcm_i16"]), name
        assert np.array_equal(A.pcm_interleave(c["wave_f32"][:, 0, :], 2)[0], c["pcm_i16"]), name
        # from the fp64 oracle waveform the truncation may flip where wave*32767 sits within 1e-3 of an integer
        diff = np.abs(A.pcm_stereo(w).astype(np.int32) - c["pcm_i16"].astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() <= 0.01, name


def test_generator_matches_reference(amp_golden):
    for name in ("gen_small_1", "gen_small_2"):
        c = amp_golden[name]
        rb = str(int(c["meta"][0]))
        h = dict(SMALL_H, resblock=rb, resblock_dilation_sizes=[[1, 3, 5]] * 3 if rb == "1" else [[1, 3]] * 3)
        y = A.generator_forward(_sd64(c["sd"]), c["mel"].astype(np.float64), h)
        assert y.shape == c["y_f64"].shape == (2, 1, 13 * 8)
        assert O.max_normalised_error(y, c["y_f64"]) <= 1e-12, name
        assert O.max_normalised_error(y, c["y_f32"]) <= 1e-5, name


def test_conv_restatements_against_torch():
    """conv1d / conv_transpose1d restatements agree with the torch ops the reference calls (float64)."""
    import torch
    import torch.nn.functional as F

    rng = np.random.default_rng(0)
    for (ci, co, k, d, T) in ((3, 5, 3, 1, 17), (4, 4, 7, 3, 40), (2, 3, 11, 5, 64)):
        x = rng.standard_normal((2, ci, T))
        w = rng.standard_normal((co, ci, k))
        b = rng.standard_normal(co)
        p = A.get_padding(k, d)
        ref = F.conv1d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), 1, p, d).numpy()
        assert np.abs(A.conv1d(x, w, b, p, d) - ref).max() <= 1e-12
    for (ci, co, k, u, T) in ((4, 2, 8, 4, 9), (6, 3, 4, 2, 21)):
        x = rng.standard_normal((2, ci, T))
        w = rng.standard_normal((ci, co, k))
        b = rng.standard_normal(co)
        ref = F.conv_transpose1d(torch.from_numpy(x), torch.from_numpy(w), torch.from_numpy(b), u, (k - u) // 2).numpy()
        assert np.abs(A.conv_transpose1d(x, w, b, u, (k - u) // 2) - ref).max() <= 1e-12


def test_zero_frame_handling_matches_reference(amp_golden):
    """afa_b200/ingest.py against the reference's own detect_and_exclude_zero_frames / reconstruct_audio_with_silence
    (inference_e2e.py:38-111) outputs: index and copy work, so bit-exact."""
    from afa_b200 import ingest

    for name in ("zf_some", "zf_none", "zf_all_but_one"):
        c = amp_golden[name]
        filt, mask, idx = ingest.detect_zero_frames(c["mel"])
        assert np.array_equal(filt, c["filtered"]) and np.array_equal(mask, c["zero_mask"]), name
        assert np.array_equal(idx, c["nonzero_indices"]), name
        restored = ingest.restore_silence_host(c["audio"], idx, 256, c["mel"].shape[1] * 256)
        assert restored.dtype == c["restored"].dtype and np.array_equal(restored, c["restored"]), name


def test_wav_roundtrip(tmp_path):
    from scipy.io import wavfile

    from afa_b200 import ingest

    pcm = (np.random.default_rng(1).integers(-32767, 32767, size=(1000, 2))).astype(np.int16)
    path = str(tmp_path / "x.wav")
    ingest.write_wav(path, 22050, pcm)
    sr, back = wavfile.read(path)
    assert sr == 22050 and back.dtype == np.int16 and np.array_equal(back, pcm)
    np.save(str(tmp_path / "m.npy"), np.zeros((1, 80, 7), dtype=np.float64))
    assert ingest.load_mel_npy(str(tmp_path / "m.npy")).shape == (80, 7)
