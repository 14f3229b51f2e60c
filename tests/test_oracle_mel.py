"""The log-mel oracle (oracle/mel_oracle.py) against vectors produced by the reference's own `mel_spectrogram` and
`MultiScaleMelSpectrogramLoss` (tests/golden/mel_golden.npz <- tests/golden/make_golden_mel.py), the Slaney
filterbank restatements (oracle and product) against the independent implementation stored in the same fixture,
and the host-side logic of afa_b200/mel.py that needs no GPU."""
import os

import numpy as np
import pytest

from oracle import mel_oracle as M

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SR = 22050


@pytest.fixture(scope="module")
def mel_golden():
    return dict(np.load(os.path.join(REPO, "tests", "golden", "mel_golden.npz")))


def _basis_keys(g):
    for k in g:
        if k.startswith("basis_"):
            parts = k.split("_")
            yield k, int(parts[1]), int(parts[2]), (8000 if k.endswith("fmax8000") else None)


def test_filterbank_restatements_match_independent_implementation(mel_golden):
    from afa_b200.mel import slaney_mel_filterbank

    n = 0
    for k, n_mels, n_fft, fmax in _basis_keys(mel_golden):
        ref = mel_golden[k]
        scale = np.abs(ref).max()
        assert np.abs(M.slaney_mel_filterbank(SR, n_fft, n_mels, 0, fmax) - ref).max() <= 1e-7 * scale, k
        assert np.abs(slaney_mel_filterbank(SR, n_fft, n_mels, 0, fmax) - ref).max() <= 1e-6 * scale, k
        n += 1
    assert n == 9


def test_filterbank_known_properties():
    """Slaney normalisation: every filter integrates to ~1 over Hz; supports are contiguous and ordered."""
    fb = M.slaney_mel_filterbank(SR, 1024, 80, 0, None).astype(np.float64)
    df = SR / 1024
    area = fb.sum(axis=1) * df
    assert np.all(np.abs(area[10:] - 1.0) < 0.05)            # the narrow low bands sample their triangle coarsely
    peaks = fb.argmax(axis=1)
    assert np.all(np.diff(peaks) > 0)
    assert M.hz_to_mel(1000.0) == pytest.approx(15.0) and M.mel_to_hz(15.0) == pytest.approx(1000.0)
    assert M.mel_to_hz(M.hz_to_mel(np.array([50.0, 999.0, 4000.0, 11025.0]))) == pytest.approx([50.0, 999.0, 4000.0, 11025.0])


def test_mel_spectrogram_matches_reference(mel_golden):
    g = mel_golden
    y = g["y"]
    assert np.abs(M.mel_spectrogram(y, 1024, 80, SR, 256, 1024, 0, None) - g["mel_2d"]).max() <= 5e-6
    assert np.abs(M.mel_spectrogram(y[0], 1024, 80, SR, 256, 1024, 0, None)[0] - g["mel_1d"]).max() <= 5e-6
    short = M.mel_spectrogram(y[:, :1500], 1024, 80, SR, 256, 1024, 0, None)
    assert short.shape == g["mel_short"].shape == (3, 80, 5)
    assert np.abs(short - g["mel_short"]).max() <= 5e-6
    # frames: segment_size / hop (the training invariant, train_binaural_mel.py segment 8192 -> 32 frames)
    assert g["mel_2d"].shape == (3, 80, 8192 // 256)


def test_multiscale_mels_and_loss_match_reference(mel_golden):
    g = mel_golden
    for w, nm in zip(M.MSMSL_WINDOWS, M.MSMSL_N_MELS):
        ref = g[f"msl_mels_{w}"]
        mine = M.msmsl_mels(g["msl_x"], SR, nm, w)
        assert mine.shape == ref.shape
        assert np.abs(mine - ref).max() <= 1e-6 * np.abs(ref).max(), w
    assert M.msmsl_loss(g["msl_x"], g["msl_y"], SR) == pytest.approx(float(g["msl_loss"]), rel=1e-6)


def test_banded_form_is_exact(mel_golden):
    from afa_b200.mel import banded

    for k, n_mels, n_fft, fmax in _basis_keys(mel_golden):
        basis = mel_golden[k]
        starts, lens, offs, w = banded(basis)
        dense = np.zeros_like(basis)
        for m in range(n_mels):
            dense[m, starts[m]:starts[m] + lens[m]] = w[offs[m]:offs[m] + lens[m]]
        assert np.array_equal(dense, basis), k
        assert lens.sum() <= w.size


def test_num_frames_and_argument_errors_without_gpu():
    """Host-side entry points answer without a device; compute entry points refuse CPU tensors (no fallback)."""
    import torch

    from afa_b200 import mel as P

    assert P.num_frames(8192, 1024, 256, 384) == 32          # center=False, pad (n_fft - hop) / 2
    assert P.num_frames(1500, 1024, 256, 384) == 5
    assert P.num_frames(4096, 2048, 512, 1024) == 9          # center=True
    assert P.num_frames(100, 1024, 256, 0) == 0
    from afa_b200._lib import load_library

    lib = load_library()
    for T in (1, 40, 63, 64, 65, 1500, 8192, 220416):
        for n, hop, pad in ((64, 16, 0), (64, 16, 32), (1024, 256, 384), (2048, 512, 1024)):
            assert P.num_frames(T, n, hop, pad) == lib.afa_logmel_num_frames(T, n, hop, pad), (T, n, hop, pad)
    with pytest.raises(ValueError):
        P.MelPlan(1000, torch.hann_window(1000), np.zeros((80, 501), np.float32), "cpu")
    plan = P.MelPlan(64, torch.hann_window(64), M.slaney_mel_filterbank(SR, 64, 10), "cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.logmel(torch.zeros(2, 4096), plan, 16, 32)
    with pytest.raises(NotImplementedError):
        P.MultiScaleMelSpectrogramLoss(SR, match_stride=True)


def test_backward_oracle_matches_reference_autograd(mel_golden):
    """The reverse-mode restatement against gradients autograd took through the reference's own functions."""
    g = mel_golden
    b = M.slaney_mel_filterbank(SR, 1024, 80)
    gy = M.logmel_backward(g["y"], g["bwd_G"], b, 1024, 256, 384, "reflect", 1e-9, 1e-5, 1.0)
    assert np.abs(gy - g["bwd_gy_2d"]).max() <= 1e-5 * np.abs(g["bwd_gy_2d"]).max()
    gy1 = M.logmel_backward(g["y"][:1], g["bwd_G"][:1], b, 1024, 256, 384, "constant", 1e-9, 1e-5, 1.0)
    assert np.abs(gy1[0] - g["bwd_gy_1d"]).max() <= 1e-5 * np.abs(g["bwd_gy_1d"]).max()
    # the L1 of the loss is not smooth: where |log mel(x) - log mel(y)| is at float32 rounding level the reference's
    # sign() and the float64 one differ, each flip moving the gradient by 2 / numel of that scale -- hence the budget
    gx = M.msmsl_loss_backward(g["msl_x"], g["msl_y"], SR)
    assert np.abs(gx - g["msl_gx"]).max() <= 1e-2 * np.abs(g["msl_gx"]).max()


def test_backward_oracle_is_the_derivative_of_the_forward_oracle():
    """Directional finite difference in float64 (independent of any fixture)."""
    rng = np.random.default_rng(0)
    y = 0.3 * rng.standard_normal((2, 700))
    G = rng.standard_normal((2, 10, M.stft_mag(y, 64, 16, M.hann_periodic(64), 32, "reflect", 0.0).shape[2]))
    basis = M.slaney_mel_filterbank(SR, 64, 10)
    d = rng.standard_normal(y.shape)

    def f(v):
        mel = np.einsum("mk,rkf->rmf", basis.astype(np.float64), M.stft_mag(v, 64, 16, M.hann_periodic(64), 32, "reflect", 1e-9))
        return float((np.log(np.maximum(mel, 1e-5)) * G).sum())

    gy = M.logmel_backward(y, G, basis, 64, 16, 32, "reflect", 1e-9, 1e-5, 1.0)
    h = 1e-6
    fd = (f(y + h * d) - f(y - h * d)) / (2 * h)
    assert fd == pytest.approx(float((gy * d).sum()), rel=1e-6)


def test_bin_cover_bounds_every_filter(mel_golden):
    from afa_b200.mel import banded, bin_cover

    for k, n_mels, n_fft, fmax in _basis_keys(mel_golden):
        basis = mel_golden[k]
        starts, lens, _, _ = banded(basis)
        lo, hi = bin_cover(starts, lens, basis.shape[1])
        for b in range(basis.shape[1]):
            ms = np.flatnonzero(basis[:, b])
            if ms.size:
                assert lo[b] <= ms[0] and ms[-1] < hi[b], (k, b)


@pytest.mark.parametrize("T,n_fft,hop,n_mels", [(4096, 1024, 256, 80), (1500, 1024, 256, 80), (777, 256, 64, 40), (9000, 2048, 512, 128),
                                               (300, 64, 16, 10), (4099, 512, 128, 80)])
def test_oracle_matches_torch_chain_on_cpu(T, n_fft, hop, n_mels):
    """The numpy restatement against the torch-op restatement of the same reference lines (oracle/torch_path.py), executed
    on the CPU for shapes outside the committed fixture: ragged lengths, every pad / hop relation the reference can produce."""
    import torch

    from oracle import torch_path as TP

    rng = np.random.default_rng(T + n_fft)
    y = (0.3 * rng.standard_normal((2, T))).clip(-1, 1).astype(np.float32)
    basis = M.slaney_mel_filterbank(SR, n_fft, n_mels)
    ref = TP.mel_spectrogram_torch(torch.from_numpy(y).double(), torch.from_numpy(basis).double(), n_fft, hop, n_fft).numpy()
    mine = M.mel_spectrogram(y, n_fft, n_mels, SR, hop, n_fft, 0, None, mel_basis=basis)
    assert mine.shape == ref.shape
    assert np.abs(mine - ref).max() <= 5e-6          # the torch chain's hann window is float32 (meldataset.py:93)
    # the loss's variant (center=True, |.|, log10) for the same window
    ref2 = TP.msmsl_logmels_torch(torch.from_numpy(y).double()[:, None, :], torch.from_numpy(basis).double(), n_fft).numpy()
    mels = M.msmsl_mels(y[:, None, :].astype(np.float64), SR, n_mels, n_fft, mel_basis=basis)
    mine2 = np.log(np.maximum(mels, 1e-5)) / np.log(10.0)
    assert mine2.shape == ref2.shape
    assert np.abs(mine2 - ref2).max() <= 5e-6       # float32 window; torch divides by a float32 log(10) tensor (loss.py:196)


def test_gan_loss_helpers_match_the_reference_and_do_not_sync():
    """afa_b200/losses.py against BigVGAN/loss.py:213-257 executed from the reference tree (skipped where it is not mounted):
    same values, same gradients; the per-discriminator terms come back as tensors (no `.item()`)."""
    import importlib.util
    import sys
    import types

    import torch

    ref_root = os.environ.get("AFA_REFERENCE_ROOT", "/root/reference")
    path = os.path.join(ref_root, "BigVGAN", "loss.py")
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    from afa_b200 import losses as L

    for name in ("librosa", "librosa.filters"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["librosa.filters"].mel = getattr(sys.modules["librosa.filters"], "mel", lambda *a, **k: None)
    spec = importlib.util.spec_from_file_location("ref_loss_for_gan_helpers", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(5)
    real = [torch.randn(3, 1, n, generator=g, requires_grad=True) for n in (40, 25, 17)]
    fake = [torch.randn(3, 1, n, generator=g, requires_grad=True) for n in (40, 25, 17)]
    fr = [[torch.randn(3, 4, n, generator=g) for n in (9, 5)] for _ in range(3)]
    fg = [[torch.randn(3, 4, n, generator=g, requires_grad=True) for n in (9, 5)] for _ in range(3)]
    l_ref, r_ref, g_ref = ref.discriminator_loss(real, fake)
    l_our, r_our, g_our = L.discriminator_loss(real, fake)
    assert torch.equal(l_ref, l_our)
    assert all(isinstance(t, torch.Tensor) and t.dim() == 0 and not t.requires_grad for t in r_our + g_our)
    assert [float(t) for t in r_our] == r_ref and [float(t) for t in g_our] == g_ref
    gl_ref, gterms_ref = ref.generator_loss(fake)
    gl_our, gterms_our = L.generator_loss(fake)
    assert torch.equal(gl_ref, gl_our) and all(torch.equal(a, b) for a, b in zip(gterms_ref, gterms_our))
    f_ref, f_our = ref.feature_loss(fr, fg), L.feature_loss(fr, fg)
    assert torch.equal(f_ref, f_our)
    g1 = torch.autograd.grad(l_ref + gl_ref + f_ref, fake + [t for d in fg for t in d])
    g2 = torch.autograd.grad(l_our + gl_our + f_our, fake + [t for d in fg for t in d])
    assert all(torch.equal(a, b) for a, b in zip(g1, g2))
