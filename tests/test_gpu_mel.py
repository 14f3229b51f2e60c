"""GPU parity of the fused log-mel kernel (afa_logmel_fwd, through the C ABI / afa_b200.mel) against the vectors the
reference's own `mel_spectrogram` / `MultiScaleMelSpectrogramLoss` produced (tests/golden/mel_golden.npz), the float64
oracle, and size-independent properties at BASELINE config 5's batch.

Tolerances.  Linear mel magnitudes: E = max|m - ref| / max|ref| <= 1e-5 (SURVEY.md 8d's fp32 budget; torch's own
fp32 chain sits at ~1e-7 from the float64 oracle).  Log mels: |diff| <= 1e-4 absolute on broadband inputs (a log
turns the relative error of a bin into an absolute one; the fixtures are broadband so no bin sits at the noise floor).
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle as M
from oracle import torch_path as TP

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEV = "cuda:0"
SR = 22050
TOL_LIN = 1e-5
TOL_LOG = 1e-4


@pytest.fixture(scope="module")
def mel_golden():
    return dict(np.load(os.path.join(REPO, "tests", "golden", "mel_golden.npz")))


def _P():
    from afa_b200 import mel as P

    return P


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_mel_spectrogram_matches_reference_vectors(mel_golden):
    P = _P()
    g = mel_golden
    y = _dev(g["y"])
    out = P.mel_spectrogram(y, 1024, 80, SR, 256, 1024, 0, None)
    assert out.shape == (3, 80, 32) and out.dtype == torch.float32
    assert np.abs(out.cpu().numpy() - g["mel_2d"]).max() <= TOL_LOG
    assert np.abs(out.cpu().numpy() - M.mel_spectrogram(g["y"], 1024, 80, SR, 256, 1024, 0, None)).max() <= TOL_LOG
    one = P.mel_spectrogram(y[0], 1024, 80, SR, 256, 1024, 0, None)           # 1-D branch: zero padding, no batch dim
    assert one.shape == (80, 32)
    assert np.abs(one.cpu().numpy() - g["mel_1d"]).max() <= TOL_LOG
    from types import SimpleNamespace

    h = SimpleNamespace(n_fft=1024, num_mels=80, sampling_rate=SR, hop_size=256, win_size=1024, fmin=0, fmax=None)
    assert torch.equal(P.get_mel_spectrogram(y, h), out)
    short = P.mel_spectrogram(y[:, :1500], 1024, 80, SR, 256, 1024, 0, None)  # pitched rows, T % hop != 0
    assert short.shape == (3, 80, 5)
    assert np.abs(short.cpu().numpy() - g["mel_short"]).max() <= TOL_LOG


def test_basis_argument_and_fmax(mel_golden):
    """A caller-supplied basis (e.g. librosa's own, or a checkpointed one) is used as given."""
    P = _P()
    g = mel_golden
    y = _dev(g["y"])
    basis = g["basis_80_1024_fmax8000"]
    a = P.mel_spectrogram(y, 1024, 80, SR, 256, 1024, 0, 8000)
    b = P.mel_spectrogram(y, 1024, 80, SR, 256, 1024, 0, None, mel_basis=basis)
    ref = M.mel_spectrogram(g["y"], 1024, 80, SR, 256, 1024, 0, 8000, mel_basis=basis)
    assert np.abs(b.cpu().numpy() - ref).max() <= TOL_LOG
    assert np.abs(a.cpu().numpy() - ref).max() <= TOL_LOG


def test_every_multiscale_window_matches_reference_vectors(mel_golden):
    """All seven STFT sizes the kernel is instantiated for (32 ... 2048), raw mel magnitudes."""
    P = _P()
    g = mel_golden
    x = _dev(g["msl_x"])
    B, C, T = x.shape
    for w, nm in zip(M.MSMSL_WINDOWS, M.MSMSL_N_MELS):
        plan = P.MelPlan(w, torch.hann_window(w, dtype=torch.float64), g[f"basis_{nm}_{w}"], DEV)
        raw = P.logmel(x.reshape(B * C, T), plan, w // 4, w // 2, mag_eps=0.0, raw=True).view(B, C, nm, -1).cpu().numpy()
        ref = g[f"msl_mels_{w}"]
        assert raw.shape == ref.shape, w
        assert np.abs(raw - ref).max() <= TOL_LIN * np.abs(ref).max(), w
        ora = M.msmsl_mels(g["msl_x"], SR, nm, w)
        assert np.abs(raw - ora).max() <= TOL_LIN * np.abs(ora).max(), w


def test_multiscale_loss_matches_reference(mel_golden):
    P = _P()
    g = mel_golden
    loss = P.MultiScaleMelSpectrogramLoss(SR).to(DEV)
    with torch.no_grad():
        v = float(loss(_dev(g["msl_x"]), _dev(g["msl_y"])))
    assert v == pytest.approx(float(g["msl_loss"]), rel=1e-4)
    xr = _dev(g["msl_x"]).requires_grad_(True)
    loss(xr, _dev(g["msl_y"])).backward()
    ref = g["msl_gx"]
    # the loss's L1 is not smooth: a |log mel(x) - log mel(y)| at rounding level flips sign() between implementations
    # (the float64 oracle sits at 2e-3 from the reference's float32 autograd for the same reason)
    assert np.abs(xr.grad.cpu().numpy() - ref).max() <= 1e-2 * np.abs(ref).max()


def test_training_batch_against_oracle_and_torch_chain():
    """BASELINE config 5's batch (32 segments of 8192 samples): fused kernel vs float64 oracle vs the torch-op chain
    of the reference on the same device."""
    P = _P()
    torch.manual_seed(1234)
    y = (0.3 * torch.randn(32, 8192, device=DEV)).clamp(-1, 1)
    out = P.mel_spectrogram(y, 1024, 80, SR, 256, 1024, 0, None, check_range=False)
    ref = M.mel_spectrogram(y.cpu().numpy(), 1024, 80, SR, 256, 1024, 0, None)
    assert np.abs(out.cpu().numpy() - ref).max() <= TOL_LOG
    basis = torch.from_numpy(M.slaney_mel_filterbank(SR, 1024, 80, 0, None)).to(DEV)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        chain = TP.mel_spectrogram_torch(y, basis, 1024, 256, 1024)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert (out - chain).abs().max().item() <= TOL_LOG
    e_ours = np.abs(out.cpu().numpy() - ref).max()
    e_torch = np.abs(chain.cpu().numpy() - ref).max()
    print(f"log-mel vs float64 oracle: fused {e_ours:.2e}, torch chain {e_torch:.2e}")


def test_size_independent_properties_at_clip_length():
    """10 s clips (220 416 samples, 861 frames): homogeneity (bit-exact under a power-of-two gain, mag_eps = 0) and
    shift-by-one-hop equivariance of the interior frames (bit-exact: the same samples meet the same arithmetic)."""
    P = _P()
    torch.manual_seed(7)
    T, hop, n_fft = 220416, 256, 1024
    y = 0.2 * torch.randn(4, T, device=DEV)
    plan = P.MelPlan(n_fft, torch.hann_window(n_fft, dtype=torch.float64), M.slaney_mel_filterbank(SR, n_fft, 80), DEV)
    pad = (n_fft - hop) // 2
    a = P.logmel(y, plan, hop, pad, mag_eps=0.0, raw=True)
    assert a.shape == (4, 80, 861)
    b = P.logmel(y * 4.0, plan, hop, pad, mag_eps=0.0, raw=True)
    assert torch.equal(b, a * 4.0)
    shifted = torch.roll(y, -hop, dims=1)
    c = P.logmel(shifted, plan, hop, pad, mag_eps=0.0, raw=True)
    assert torch.equal(c[:, :, 2:-4], a[:, :, 3:-3])
    # frames of a row do not depend on the other rows or on the launch's frame tiling
    d = P.logmel(y[1:2, : 100 * hop], plan, hop, pad, mag_eps=0.0, raw=True)
    assert torch.equal(d[0, :, :95], a[1, :, :95])
    assert torch.isfinite(a).all()


TOL_GRAD = 1e-4


def test_fused_l1_path_equals_generic_autograd_path(mel_golden):
    """The default loss (nn.L1Loss) runs with the L1 inside the kernels (afa_l1_partial_sums forward, sign formed in
    afa_logmel_bwd, scales accumulated in place); any other loss_fn takes the per-scale autograd path.  Same log mels,
    same signs -> same loss and the same gradients for BOTH waveforms, up to summation order."""
    P = _P()
    g = mel_golden
    fused = P.MultiScaleMelSpectrogramLoss(SR)
    generic = P.MultiScaleMelSpectrogramLoss(SR, loss_fn=lambda a, b: (a - b).abs().mean())
    res = []
    for loss in (fused, generic):
        x = _dev(g["msl_x"]).requires_grad_(True)
        y = _dev(g["msl_y"]).requires_grad_(True)
        v = loss(x, y)
        (v * 60.0).backward()                                   # lambda_melloss: the upstream gradient is not 1
        res.append((float(v.detach()), x.grad.clone(), y.grad.clone()))
    assert res[0][0] == pytest.approx(res[1][0], rel=1e-5)
    assert res[0][0] == pytest.approx(float(g["msl_loss"]), rel=1e-4)
    for a, b in ((res[0][1], res[1][1]), (res[0][2], res[1][2])):
        assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item()
    assert np.abs(res[0][1].cpu().numpy() - g["msl_gx"] * 60.0).max() <= 1e-2 * np.abs(g["msl_gx"] * 60.0).max()
    # only the estimate differentiated (the training step): the target's gradient is not computed
    x = _dev(g["msl_x"]).requires_grad_(True)
    fused(x, _dev(g["msl_y"])).backward()
    assert torch.equal(x.grad * 60.0, res[0][1]) or (x.grad * 60.0 - res[0][1]).abs().max().item() <= 1e-6 * res[0][1].abs().max().item()


def test_l1_partial_sums():
    P = _P()
    torch.manual_seed(2)
    for n in (1, 255, 4097, 32 * 80 * 32, 1_000_003):
        a, b = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
        out = torch.full((P.L1_PARTIALS,), float("nan"), device=DEV)
        P.l1_partial_sums(a, b, out)
        ref = (a.double() - b.double()).abs().sum().item()
        assert out.double().sum().item() == pytest.approx(ref, rel=1e-5), n
        again = torch.empty_like(out)
        P.l1_partial_sums(a, b, again)
        assert torch.equal(out, again)
    with pytest.raises(ValueError):
        P.l1_partial_sums(torch.zeros(4, device=DEV), torch.zeros(5, device=DEV), torch.zeros(8, device=DEV))


def test_backward_matches_reference_autograd_vectors(mel_golden):
    """d <mel_spectrogram(y), G> / d y through afa_logmel_bwd vs autograd through the reference's own op chain."""
    P = _P()
    g = mel_golden
    y = _dev(g["y"]).requires_grad_(True)
    G = _dev(g["bwd_G"])
    (P.mel_spectrogram(y, 1024, 80, SR, 256, 1024, 0, None) * G).sum().backward()
    ref = g["bwd_gy_2d"]
    e = np.abs(y.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
    y1 = _dev(g["y"][0]).requires_grad_(True)
    (P.mel_spectrogram(y1, 1024, 80, SR, 256, 1024, 0, None) * G[0]).sum().backward()      # zero-pad branch
    ref1 = g["bwd_gy_1d"]
    e1 = np.abs(y1.grad.cpu().numpy() - ref1).max() / np.abs(ref1).max()
    print(f"log-mel backward vs reference autograd: reflect {e:.2e}, zero-pad {e1:.2e}")
    assert e <= TOL_GRAD and e1 <= TOL_GRAD


def test_backward_every_window_against_oracle(mel_golden):
    """All seven STFT sizes, log10 mels with |.| magnitudes (the loss's configuration), ragged length, vs float64."""
    P = _P()
    torch.manual_seed(11)
    T = 3000
    x = (0.3 * torch.randn(3, T, device=DEV)).clamp(-1, 1)
    for w, nm in zip(M.MSMSL_WINDOWS, M.MSMSL_N_MELS):
        basis = mel_golden[f"basis_{nm}_{w}"]
        plan = P.MelPlan(w, torch.hann_window(w, dtype=torch.float64), basis, DEV)
        xr = x.clone().requires_grad_(True)
        out = P.logmel(xr, plan, w // 4, w // 2, mag_eps=0.0, log_scale=1.0 / math.log(10.0))
        G = torch.randn_like(out)
        (out * G).sum().backward()
        ref = M.logmel_backward(x.cpu().numpy(), G.cpu().numpy(), basis, w, w // 4, w // 2, "reflect", 0.0, 1e-5, 1.0 / math.log(10.0))
        e = np.abs(xr.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert e <= TOL_GRAD, (w, e)
        raw = P.logmel(xr, plan, w // 4, w // 2, mag_eps=0.0, raw=True)
        xr.grad = None
        (raw * G).sum().backward()
        ref = M.logmel_backward(x.cpu().numpy(), G.cpu().numpy(), basis, w, w // 4, w // 2, "reflect", 0.0, 1e-5, 1.0, raw=True)
        e = np.abs(xr.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert e <= TOL_GRAD, ("raw", w, e)


def test_backward_training_batch_properties():
    """BASELINE config 5's batch: bitwise reproducible (the overlap-add is a fixed-order gather), bitwise linear in the
    output gradient under a power-of-two gain, equal to the float64 oracle, and equal to autograd through the torch chain."""
    P = _P()
    torch.manual_seed(5)
    y = (0.3 * torch.randn(32, 8192, device=DEV)).clamp(-1, 1)
    basis = M.slaney_mel_filterbank(SR, 1024, 80)
    plan = P.MelPlan(1024, torch.hann_window(1024, dtype=torch.float64), basis, DEV)
    G = torch.randn(32, 80, 32, device=DEV)
    cfg = (plan, 256, 384, P.AFA_MEL_PAD_REFLECT, 1e-9, 1e-5, 1.0, False)
    a = P.logmel_backward_raw(y, G, *cfg)
    b = P.logmel_backward_raw(y, G, *cfg)
    assert torch.equal(a, b)
    assert torch.equal(P.logmel_backward_raw(y, G * 2.0, *cfg), a * 2.0)
    ref = M.logmel_backward(y.cpu().numpy(), G.cpu().numpy(), basis, 1024, 256, 384, "reflect", 1e-9, 1e-5, 1.0)
    assert np.abs(a.cpu().numpy() - ref).max() <= TOL_GRAD * np.abs(ref).max()
    yt = y.clone().requires_grad_(True)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        (TP.mel_spectrogram_torch(yt, torch.from_numpy(basis).to(DEV), 1024, 256, 1024) * G).sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert (a - yt.grad).abs().max().item() <= TOL_GRAD * yt.grad.abs().max().item()
    # rows shorter than one frame: no frame read them, the gradient is zero
    plan64 = P.MelPlan(64, torch.hann_window(64, dtype=torch.float64), M.slaney_mel_filterbank(SR, 64, 10), DEV)
    z = P.logmel_backward_raw(torch.randn(2, 40, device=DEV), torch.zeros(2, 10, 0, device=DEV), plan64, 16, 0, P.AFA_MEL_PAD_ZERO,
                              0.0, 1e-5, 1.0, False)
    assert z.shape == (2, 40) and not z.any()


def test_edge_cases_and_errors():
    P = _P()
    from afa_b200._lib import AfaError

    plan = P.MelPlan(64, torch.hann_window(64, dtype=torch.float64), M.slaney_mel_filterbank(SR, 64, 10), DEV)
    empty = P.logmel(torch.zeros(0, 4096, device=DEV), plan, 16, 32)
    assert empty.shape == (0, 10, 257)
    too_short = P.logmel(torch.zeros(2, 40, device=DEV), plan, 16, 0)            # T < n_fft without padding: no frames
    assert too_short.shape == (2, 10, 0)
    with pytest.raises(RuntimeError, match="Padding size"):
        P.logmel(torch.zeros(2, 20, device=DEV), plan, 16, 32)
    with pytest.raises(TypeError):
        P.logmel(torch.zeros(2, 4096, device=DEV, dtype=torch.bfloat16), plan, 16, 32)
    silence = P.logmel(torch.zeros(2, 4096, device=DEV), plan, 16, 32)           # clamp floor, as the reference gives
    assert torch.allclose(silence, torch.full_like(silence, math.log(1e-5)))
    # argument checks of the C ABI itself
    import ctypes

    from afa_b200._lib import load_library

    lib = load_library()
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int32)
    x = torch.zeros(1, 4096, device=DEV)
    o = torch.zeros(1, 10, 300, device=DEV)
    c = lambda t, ty: ctypes.cast(t.data_ptr(), ty)  # noqa: E731
    rc = lib.afa_logmel_fwd(c(x, fp), c(o, fp), 1, 4096, 4096, 48, 16, 0, 0, c(plan.window, fp), c(plan.twiddle, fp), 10,
                            c(plan.band_start, ip), c(plan.band_len, ip), c(plan.band_off, ip), c(plan.band_w, fp),
                            0.0, 1e-5, 1.0, 0, None)
    assert rc < 0 and b"power of two" in lib.afa_last_error()
    with pytest.raises(AfaError):
        P.check(rc, "afa_logmel_fwd")


def test_cuda_graph_capture_and_side_stream():
    """No allocation, no synchronisation inside the launch: capturable, and it runs on the caller's stream."""
    P = _P()
    torch.manual_seed(3)
    y = 0.3 * torch.randn(8, 8192, device=DEV)
    plan = P.MelPlan(1024, torch.hann_window(1024, dtype=torch.float64), M.slaney_mel_filterbank(SR, 1024, 80), DEV)
    eager = P.logmel(y, plan, 256, 384)
    out = torch.empty_like(eager)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        P.logmel(y, plan, 256, 384, out=out)
    torch.cuda.current_stream().wait_stream(s)
    assert torch.equal(out, eager)
    out.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        P.logmel(y, plan, 256, 384, out=out)
    y.mul_(0.5)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, P.logmel(y, plan, 256, 384))
