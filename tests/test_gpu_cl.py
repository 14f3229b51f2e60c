"""GPU parity of the channels-last AMP-block kernels (through the C ABI) against the oracle and the vectors the
unmodified reference produced (tests/golden/amp_golden.npz).

Tolerances, E = max|y - ref| / max|ref|: fp32 <= 1e-5; bf16 I/O (fp32 math) <= 1e-2;
int16 PCM: |diff| <= 1 LSB and >= 99 % of the samples exact (the reference truncates a float that our fp32
result reproduces to ~1e-7 relative, so a value within that distance of an integer may truncate differently).
"""
import os

import numpy as np
import pytest
import torch

from oracle import afa_oracle as O
from oracle import amp_oracle as A
from oracle import torch_path as TP

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL_F32 = 1e-5
TOL_BF16 = 1e-2
DEV = "cuda:0"


def _fc():
    from afa_b200 import functional as F_afa
    from afa_b200 import functional_cl as FC

    return F_afa, FC


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


@pytest.fixture
def true_fp32_convs():
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _taps():
    F_afa, _ = _fc()
    t = TP.make_taps()
    h = F_afa.host_taps(t)
    return t, (h, h), t.reshape(-1).double().numpy()


def _padded(x: torch.Tensor, rows: int, fill=float("nan")) -> torch.Tensor:
    """[B, T, C] -> a [B, T, C] view into a [B, rows, C] buffer (batch stride rows*C), padding poisoned."""
    B, T, C = x.shape
    buf = torch.full((B, rows, C), fill, dtype=x.dtype, device=x.device)
    buf[:, :T] = x
    return buf[:, :T]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cl_activation_grid(dtype):
    """Edge grid: T around the segment / ring boundaries, odd channel counts, Snake / SnakeBeta, log-scale on/off,
    with and without bias / residual / xsum, padded batch strides, zero-filled tail rows."""
    _, FC = _fc()
    t32, taps, taps64 = _taps()
    rng = np.random.default_rng(5)
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    Ts = [1, 2, 3, 5, 6, 11, 12, 13, 23, 24, 25, 29, 47, 48, 49, 95, 96, 97, 101, 191, 192, 193, 300, 1000]
    Cs = [1, 3, 8, 24, 32, 33, 48]
    n = 0
    for T in Ts:
        for C in Cs:
            if (T * C) % 3 == 1 and T > 100:
                continue
            B = 1 + (n % 3)
            kind = ("snakebeta", "snake")[n % 2]
            logscale = (n % 4) < 3
            with_bias, with_res = (n % 3) != 0, (n % 5) in (1, 2, 3)
            with_xsum = with_res
            n += 1
            x = torch.tensor(rng.standard_normal((B, T, C)), dtype=dtype, device=DEV)
            res = torch.tensor(rng.standard_normal((B, T, C)), dtype=dtype, device=DEV) if with_res else None
            bias = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV) if with_bias else None
            if logscale:
                alpha = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
                beta = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
            else:
                alpha = torch.tensor(rng.random(C) * 2 + 0.25, dtype=torch.float32, device=DEV)
                beta = torch.tensor(rng.random(C) * 2 + 0.25, dtype=torch.float32, device=DEV)
            if kind == "snake":
                beta = None
            tpad = T + (n % 4) * 2
            xin = _padded(x, T + 3) if n % 2 else x
            rin = _padded(res, T + 1) if (with_res and n % 3 == 0) else res
            out = torch.full((B, tpad + 2, C), 7.0, dtype=dtype, device=DEV)
            xsum = torch.full((B, T, C), 9.0, dtype=dtype, device=DEV) if with_xsum else None
            y = FC.amp_activation1d_cl(xin, T, alpha, beta, taps[0], taps[1], logscale, bias=bias, res=rin, xsum=xsum,
                                       out=out, out_tpad=tpad)
            torch.cuda.synchronize()
            xs_ref, y_ref = A.amp_activation1d_cl(
                x.double().cpu().numpy(), alpha.double().cpu().numpy(), None if beta is None else beta.double().cpu().numpy(),
                logscale, None if bias is None else bias.double().cpu().numpy(), None if res is None else res.double().cpu().numpy(),
                taps64, taps64)
            tag = f"T={T} C={C} B={B} {kind} log={logscale} bias={with_bias} res={with_res} xsum={with_xsum}"
            got = y[:, :T].double().cpu().numpy()
            assert O.max_normalised_error(got, y_ref) <= tol, (tag, O.max_normalised_error(got, y_ref))
            if tpad > T:
                assert torch.all(out[:, T:tpad] == 0), tag             # zero padding for the polyphase convolution
            assert torch.all(out[:, tpad:] == 7.0), tag                # nothing written beyond it
            if with_xsum:
                assert O.max_normalised_error(xsum.double().cpu().numpy(), xs_ref) <= (1e-6 if dtype == torch.float32 else 4e-3), tag


def test_cl_tensor_core_activation_grid():
    """The channels-last tensor-core forward (afa_tc_cl_fwd_kernel; bf16, no residual prologue), forced onto small tensors
    the built-in choice leaves to the walk kernel: T around the 32-output block / 64-step chunk / strip boundaries with both
    residues of T % 8, channel counts around the 32-channel units (four per CTA, each with its own batch entry / strip / channel
    quad, the last CTA partly empty), Snake / SnakeBeta, log-scale on /
    off, bias on / off, padded batch strides, zero rows behind T, guard rows behind those; forced strip lengths.  Against the
    float64 oracle on the bf16 inputs, and against the walk kernel (both round y to bf16 once: they differ by the rounding
    of s to bf16 in front of the down filter)."""
    from afa_b200 import _lib

    _, FC = _fc()
    t32, taps, taps64 = _taps()
    rng = np.random.default_rng(11)
    Ts = [64, 68, 72, 100, 124, 128, 132, 252, 256, 260, 316, 508, 512, 516, 1000, 1028, 2052, 4100]
    Cs = [8, 24, 32, 40, 64, 72, 96, 128, 136, 192, 200, 264]
    n = 0
    try:
        for T in Ts:
            for C in Cs:
                if (n % 3) == 2 and T > 600:
                    n += 1
                    continue
                B = 1 + (n % 3)
                kind = ("snakebeta", "snake")[n % 2]
                logscale = (n % 4) < 3
                with_bias = (n % 3) != 0
                ny = (0, 4, 8, 12)[n % 4]
                n += 1
                x = torch.tensor(rng.standard_normal((B, T, C)), dtype=torch.bfloat16, device=DEV)
                bias = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV) if with_bias else None
                if logscale:
                    alpha = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
                    beta = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
                else:
                    alpha = torch.tensor(rng.random(C) * 2 + 0.25, dtype=torch.float32, device=DEV)
                    beta = torch.tensor(rng.random(C) * 2 + 0.25, dtype=torch.float32, device=DEV)
                if kind == "snake":
                    beta = None
                tpad = T + (n % 4) * 2
                xin = _padded(x, T + 8) if n % 2 else x
                tag = f"T={T} C={C} B={B} {kind} log={logscale} bias={with_bias} ny={ny} tpad={tpad}"
                out = torch.full((B, tpad + 4, C), 7.0, dtype=torch.bfloat16, device=DEV)
                _lib.set_tuning(7, 2, ny)
                before = _lib.launch_count()
                y = FC.amp_activation1d_cl(xin, T, alpha, beta, taps[0], taps[1], logscale, bias=bias, out=out, out_tpad=tpad)
                torch.cuda.synchronize()
                assert _lib.launch_count() == before + 1, tag
                _lib.set_tuning(7, 0, 0)
                yw = FC.amp_activation1d_cl(xin, T, alpha, beta, taps[0], taps[1], logscale, bias=bias, out_tpad=tpad)
                torch.cuda.synchronize()
                _, y_ref = A.amp_activation1d_cl(
                    x.double().cpu().numpy(), alpha.double().cpu().numpy(), None if beta is None else beta.double().cpu().numpy(),
                    logscale, None if bias is None else bias.double().cpu().numpy(), None, taps64, taps64)
                got = y[:, :T].double().cpu().numpy()
                assert O.max_normalised_error(got, y_ref) <= TOL_BF16, (tag, O.max_normalised_error(got, y_ref))
                assert O.max_normalised_error(got, yw[:, :T].double().cpu().numpy()) <= TOL_BF16, tag
                if tpad > T:
                    assert torch.all(out[:, T:tpad] == 0), tag
                assert torch.all(out[:, tpad:] == 7.0), tag
    finally:
        _lib.set_tuning(7, 1, 0)
    # the built-in choice takes it on the engine's wide stages and leaves residual calls, fp32 and narrow tensors alone
    info = _lib.kernel_info(7, 1, 8192)
    assert info["threads"] == 320 and info["registers"] <= 102, info


def test_cl_activation_matches_reference_golden(golden_cases):
    """The reference's own Activation1d outputs (tests/golden/activation1d_golden.npz), fed channels-last."""
    _, FC = _fc()
    for name, c in golden_cases.items():
        B, C, T, is_beta, logscale = [int(v) for v in c["meta"]]
        F_afa, _ = _fc()
        h = F_afa.host_taps(TP.make_taps())
        taps = (h, h)
        x = torch.tensor(c["x"], device=DEV).transpose(1, 2).contiguous()
        alpha = torch.tensor(c["alpha"], device=DEV)
        beta = torch.tensor(c["beta"], device=DEV) if is_beta else None
        y = FC.amp_activation1d_cl(x, T, alpha, beta, taps[0], taps[1], bool(logscale))
        got = y.transpose(1, 2).double().cpu().numpy()
        assert O.max_normalised_error(got, c["y_f64"]) <= TOL_F32, name
        assert O.max_normalised_error(got, c["y_f32"]) <= TOL_F32, name


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cl_activation_full_size_vs_torch_oracle(dtype):
    """BASELINE shapes (8 clips: B=16) against the torch-op oracle on the device; segment-length invariance (bitwise)."""
    from afa_b200 import _lib

    _, FC = _fc()
    t32, taps, _ = _taps()
    torch.manual_seed(3)
    for C, T in ((768, 3444), (96, 55104), (24, 220416)):
        B = 16 if dtype == torch.bfloat16 else 4
        x = torch.randn(B, T, C, device=DEV, dtype=dtype)
        res = torch.randn(B, T, C, device=DEV, dtype=dtype)
        bias = torch.randn(C, device=DEV) * 0.3
        alpha = torch.randn(C, device=DEV) * 0.5
        beta = torch.randn(C, device=DEV) * 0.5
        xsum = torch.empty_like(x)
        y = FC.amp_activation1d_cl(x, T, alpha, beta, taps[0], taps[1], True, bias=bias, res=res, xsum=xsum)
        xs = x.float() + res.float()
        ref = TP.activation1d_torch((xs + bias).transpose(1, 2).contiguous(), alpha, beta, True, t32.to(DEV), t32.to(DEV)).transpose(1, 2)
        scale = ref.abs().max().item()
        tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
        assert (y.float() - ref).abs().max().item() / scale <= tol, (C, T)
        assert (xsum.float() - xs).abs().max().item() / xs.abs().max().item() <= (1e-6 if dtype == torch.float32 else 4e-3)
        # plain variant (bias folded into the pending pairs), and bitwise invariance against the segment length
        y0 = FC.amp_activation1d_cl(x, T, alpha, beta, taps[0], taps[1], True, bias=bias)
        ref0 = TP.activation1d_torch((x.float() + bias).transpose(1, 2).contiguous(), alpha, beta, True, t32.to(DEV), t32.to(DEV)).transpose(1, 2)
        assert (y0.float() - ref0).abs().max().item() / ref0.abs().max().item() <= tol, (C, T)
        try:
            _lib.set_tuning(2, 2)
            y1 = FC.amp_activation1d_cl(x, T, alpha, beta, taps[0], taps[1], True, bias=bias)
            _lib.set_tuning(2, 11)
            y2 = FC.amp_activation1d_cl(x, T, alpha, beta, taps[0], taps[1], True, bias=bias)
        finally:
            _lib.set_tuning(2, 0)
        assert torch.equal(y0, y1) and torch.equal(y0, y2), (C, T)
        del x, res, xsum, y, ref, xs, y0, y1, y2, ref0
        torch.cuda.empty_cache()


def test_resblock_mean():
    _, FC = _fc()
    rng = np.random.default_rng(9)
    for dtype, tol in ((torch.float32, 1e-6), (torch.bfloat16, 4e-3)):
        for (B, T, C, K) in ((2, 37, 24, 3), (1, 5, 7, 3), (3, 16, 48, 1), (2, 33, 8, 4), (1, 1, 1, 2)):
            xts = [torch.tensor(rng.standard_normal((B, T, C)), dtype=dtype, device=DEV) for _ in range(K)]
            xrs = [torch.tensor(rng.standard_normal((B, T, C)), dtype=dtype, device=DEV) for _ in range(K)]
            bs = torch.tensor(rng.standard_normal(C), dtype=torch.float32, device=DEV)
            out = FC.resblock_mean(xts, xrs, bs)
            ref = A.resblock_mean([t.double().cpu().numpy() for t in xts], [t.double().cpu().numpy() for t in xrs], bs.double().cpu().numpy())
            assert O.max_normalised_error(out.double().cpu().numpy(), ref) <= tol, (dtype, B, T, C, K)
            out2 = FC.resblock_mean(xts, xrs, None, 0.5)
            ref2 = A.resblock_mean([t.double().cpu().numpy() for t in xts], [t.double().cpu().numpy() for t in xrs], None, 0.5)
            assert O.max_normalised_error(out2.double().cpu().numpy(), ref2) <= tol


def _tail_from_golden(c, dtype):
    F_afa, FC = _fc()
    B, C, T, use_tanh, has_bias = [int(v) for v in c["meta"]]
    sd = c["sd"]
    up = F_afa.host_taps(torch.tensor(sd["activation_post.upsample.filter"]))
    dn = F_afa.host_taps(torch.tensor(sd["activation_post.downsample.lowpass.filter"]))
    x = torch.tensor(c["x"], device=DEV).transpose(1, 2).contiguous().to(dtype)
    alpha = torch.tensor(sd["activation_post.act.alpha"], device=DEV)
    beta = torch.tensor(sd["activation_post.act.beta"], device=DEV)
    w = torch.tensor(sd["conv_post.weight"], device=DEV).reshape(C, 7).contiguous()
    b = torch.tensor(sd["conv_post.bias"], device=DEV) if has_bias else None
    return FC.tail_cl(x, T, alpha, beta, up, dn, True, w, b, use_tanh=bool(use_tanh), want_wave=True, want_pcm=True)


def test_tail_matches_reference_golden(amp_golden):
    """activation_post -> conv_post -> clamp | tanh -> int16 stereo, against the reference's own outputs."""
    for name in ("tail_clamp", "tail_tanh_bias"):
        c = amp_golden[name]
        wave, pcm = _tail_from_golden(c, torch.float32)
        got = wave.double().cpu().numpy()
        assert O.max_normalised_error(got, c["wave_f64"][:, 0, :]) <= TOL_F32, name
        assert O.max_normalised_error(got, c["wave_f32"][:, 0, :]) <= TOL_F32, name
        p = pcm.cpu().numpy()
        assert p.shape == (1,) + c["pcm_i16"].shape and p.dtype == np.int16
        diff = np.abs(p[0].astype(np.int32) - c["pcm_i16"].astype(np.int32))
        assert diff.max() <= 1 and (diff == 0).mean() >= 0.99, (name, diff.max(), (diff == 0).mean())
        # the PCM is exactly the truncation of the float wave the same launch wrote
        assert np.array_equal(p, A.pcm_interleave(wave.cpu().numpy(), 2)), name
        wave_b, _ = _tail_from_golden(c, torch.bfloat16)
        assert O.max_normalised_error(wave_b.double().cpu().numpy(), c["wave_f64"][:, 0, :]) <= TOL_BF16, name


def test_tail_edge_grid():
    _, FC = _fc()
    t32, taps, taps64 = _taps()
    rng = np.random.default_rng(21)
    for T in (1, 2, 3, 4, 7, 89, 90, 91, 93, 94, 180, 181, 500):
        for C in (1, 8, 24, 32):
            B = 2
            x = torch.tensor(rng.standard_normal((B, T, C)), dtype=torch.float32, device=DEV)
            alpha = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
            beta = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
            w = torch.tensor(rng.standard_normal((C, 7)) * 0.5 / np.sqrt(7 * C), dtype=torch.float32, device=DEV)
            wave, pcm = FC.tail_cl(x, T, alpha, beta, taps[0], taps[1], True, w, None, use_tanh=False, want_pcm=True)
            ref = A.tail_cl(x.double().cpu().numpy(), alpha.double().cpu().numpy(), beta.double().cpu().numpy(), True,
                            w.double().cpu().numpy(), None, False, taps64, taps64)
            err = np.abs(wave.double().cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-3)
            assert err <= TOL_F32, (T, C, err)
            assert np.array_equal(pcm.cpu().numpy(), A.pcm_interleave(wave.cpu().numpy(), 2)), (T, C)
    # the walk length of a warp segment is a tuning knob (afa_set_tuning(8, n): 12 n + 2 samples): same wave whatever the cut
    from afa_b200 import _lib

    x = torch.tensor(rng.standard_normal((2, 1000, 24)), dtype=torch.float32, device=DEV)
    alpha = torch.tensor(rng.standard_normal(24) * 0.5, dtype=torch.float32, device=DEV)
    beta = torch.tensor(rng.standard_normal(24) * 0.5, dtype=torch.float32, device=DEV)
    w = torch.tensor(rng.standard_normal((24, 7)) * 0.05, dtype=torch.float32, device=DEV)
    base = FC.tail_cl(x, 1000, alpha, beta, taps[0], taps[1], True, w, None, use_tanh=False, want_pcm=True)
    try:
        for n in (1, 3, 20):
            _lib.set_tuning(8, n)
            other = FC.tail_cl(x, 1000, alpha, beta, taps[0], taps[1], True, w, None, use_tanh=False, want_pcm=True)
            assert torch.equal(base[0], other[0]) and torch.equal(base[1], other[1]), n
    finally:
        _lib.set_tuning(8, 0)


def _engine_from_sd(sd, h, dtype):
    from afa_b200.engine import ChannelsLastVocoder
    from afa_b200.vocoder import BINAURAL_22KHZ_80BAND_256X, BigVGANGenerator

    hh = dict(BINAURAL_22KHZ_80BAND_256X)
    hh.update(h)
    gen = BigVGANGenerator(hh)
    gen.load_state_dict({k: torch.tensor(v) for k, v in sd.items()})
    gen = gen.to(DEV).eval()
    return gen, ChannelsLastVocoder(gen.to(dtype) if dtype != torch.float32 else gen, dtype=dtype)


def test_engine_matches_reference_generator_golden(amp_golden, true_fp32_convs):
    """Whole (small) generator pass of the channels-last engine against the reference's own output."""
    for name in ("gen_small_1", "gen_small_2"):
        c = amp_golden[name]
        rb = str(int(c["meta"][0]))
        h = dict(upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=16, resblock=rb,
                 resblock_dilation_sizes=[[1, 3, 5]] * 3 if rb == "1" else [[1, 3]] * 3)
        gen, eng = _engine_from_sd(c["sd"], h, torch.float32)
        mel = torch.tensor(c["mel"], device=DEV)
        wave, pcm = eng(mel, want_pcm=True)
        got = wave.double().cpu().numpy()
        assert got.shape == c["y_f64"].shape
        e64 = O.max_normalised_error(got, c["y_f64"])
        assert e64 <= 2e-5, (name, e64)                      # fp32 convolutions of cuDNN in between: a few 1e-6
        ref_pcm = A.pcm_stereo(c["y_f32"][:, 0, :])
        diff = np.abs(pcm.cpu().numpy()[0].astype(np.int32) - ref_pcm.astype(np.int32))
        assert diff.max() <= 1 and (diff == 0).mean() >= 0.95, (name, diff.max(), (diff == 0).mean())
        # and the [B, C, T] harness around the same weights agrees
        with torch.no_grad():
            y_h = gen(mel)
        assert (y_h - wave).abs().max().item() <= 2e-5


@pytest.mark.parametrize("resblock,activation", [("1", "snakebeta"), ("2", "snake")])
def test_engine_vs_ncw_harness_model_config(resblock, activation, true_fp32_convs):
    """The shipped config's channel plan (1536 -> 24 over six stages): fp32 engine == fp32 [B, C, T] harness; bf16 close;
    CUDA-graph replay reproduces the eager result bit for bit."""
    from afa_b200.engine import ChannelsLastVocoder, GraphedEngine
    from afa_b200.vocoder import BINAURAL_22KHZ_80BAND_256X, BigVGANGenerator

    h = dict(BINAURAL_22KHZ_80BAND_256X)
    h.update(upsample_initial_channel=192, resblock=resblock, activation=activation)
    torch.manual_seed(1234)
    gen = BigVGANGenerator(h).to(DEV).eval()
    with torch.no_grad():
        for n, p in gen.named_parameters():
            if n.endswith("alpha") or n.endswith("beta"):
                p.normal_(0, 0.5)
            elif n.endswith("weight"):
                p.mul_(6.0)
            elif n.endswith("bias"):
                p.normal_(0, 0.1)
    mel = torch.rand(2, 80, 23, device=DEV) * 14.5 - 12.0
    eng = ChannelsLastVocoder(gen, dtype=torch.float32)
    with torch.no_grad():
        y_ref = gen(mel)
    wave, _ = eng(mel)
    scale = max(y_ref.abs().max().item(), 1e-6)
    assert wave.shape == y_ref.shape == (2, 1, 23 * 256)
    err = (wave - y_ref).abs().max().item() / scale
    assert err <= 2e-4, err
    # graph capture picks cuDNN algorithms in benchmark mode: compare against an eager pass made under the same mode
    old_benchmark = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    try:
        wave_b, _ = eng(mel)
        ge = GraphedEngine(eng, 2, 23, want_pcm=True)
        w_g, p_g = ge(mel)
    finally:
        torch.backends.cudnn.benchmark = old_benchmark
    assert torch.equal(w_g, wave_b)
    assert (wave_b - y_ref).abs().max().item() / scale <= 2e-4
    eng_b = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
    w_b, _ = eng_b(mel)
    # bf16 through ~110 layers with 6x-amplified weights: judge by the relative L2 error, bound the worst sample loosely
    assert ((w_b - y_ref).norm() / y_ref.norm()).item() <= 0.05
    assert (w_b - y_ref).abs().max().item() / scale <= 0.3


def test_cl_error_codes():
    from afa_b200 import _lib

    _, FC = _fc()
    t32, taps, _ = _taps()
    x = torch.randn(2, 50, 8, device=DEV)
    a = torch.zeros(8, device=DEV)
    with pytest.raises(_lib.AfaError, match="alias"):
        FC.amp_activation1d_cl(x, 50, a, a, taps[0], taps[1], True, out=x)
    with pytest.raises(RuntimeError, match="xsum"):
        FC.amp_activation1d_cl(x, 50, a, a, taps[0], taps[1], True, xsum=torch.empty_like(x))
    with pytest.raises(RuntimeError, match="contiguous"):
        FC.amp_activation1d_cl(x.transpose(1, 2), 8, a, a, taps[0], taps[1], True)
    with pytest.raises(RuntimeError, match="CUDA"):
        FC.amp_activation1d_cl(x.cpu(), 50, a, a, taps[0], taps[1], True)
    x40 = torch.randn(2, 50, 40, device=DEV)
    a40 = torch.zeros(40, device=DEV)
    with pytest.raises(_lib.AfaError, match="32"):
        FC.tail_cl(x40, 50, a40, a40, taps[0], taps[1], True, torch.zeros(40, 7, device=DEV))
    # empty inputs are legal no-ops
    y = FC.amp_activation1d_cl(torch.empty(0, 5, 8, device=DEV), 5, a, a, taps[0], taps[1], True)
    assert y.shape == (0, 5, 8)


def _bf16_round(a):
    return torch.tensor(a, dtype=torch.float32).to(torch.bfloat16).double().numpy()


def test_act_conv_fused_grid():
    """Activation1d as the prologue of the dilated convolution (tensor-core path, bf16): every compiled channel count,
    kernel sizes 3/7/11 x dilations 1/3/5, T around the tile boundaries, with and without the residual prologue.
    Oracle: float64 activation on the bf16 inputs, rounded to bf16 (the tile the MMA reads), float64 convolution."""
    _, FC = _fc()
    t32, taps, taps64 = _taps()
    rng = np.random.default_rng(11)
    n = 0
    for C in (8, 16, 24, 32, 48, 64):
        for (k, d) in ((3, 1), (3, 5), (7, 3), (11, 1), (11, 5)):
            for T in (1, 17, 160, 415, 1000, 1733):
                if (n % 3) and T in (17, 415):
                    n += 1
                    continue
                B = 1 + n % 2
                with_res = (n % 2) == 0
                n += 1
                assert FC.act_conv_supported(C, k, d, torch.bfloat16)
                x = torch.tensor(rng.standard_normal((B, T, C)), dtype=torch.bfloat16, device=DEV)
                res = torch.tensor(rng.standard_normal((B, T, C)), dtype=torch.bfloat16, device=DEV) if with_res else None
                xsum = torch.full((B, T, C), 9.0, dtype=torch.bfloat16, device=DEV) if with_res else None
                bias = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
                alpha = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
                beta = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
                w = torch.tensor(rng.standard_normal((k, C, C)) / np.sqrt(k * C), dtype=torch.bfloat16, device=DEV)
                y = FC.amp_act_conv_cl(x, T, alpha, beta, taps[0], taps[1], True, w, k, d, bias=bias, res=res, xsum=xsum)
                torch.cuda.synchronize()
                xs_ref, y_ref = A.amp_act_conv_cl(
                    x.double().cpu().numpy(), alpha.double().cpu().numpy(), beta.double().cpu().numpy(), True,
                    w.double().cpu().numpy(), d, bias.double().cpu().numpy(), None if res is None else res.double().cpu().numpy(),
                    taps64, taps64, round_act=_bf16_round)
                tag = f"C={C} k={k} d={d} T={T} B={B} res={with_res}"
                err = O.max_normalised_error(y.double().cpu().numpy(), y_ref)
                # the tile is bf16 (a 1-ulp flip of an activated sample moves y by <= 2^-8 |w| |a|), the output is bf16
                assert err <= TOL_BF16, (tag, err)
                if with_res:
                    assert O.max_normalised_error(xsum.double().cpu().numpy(), xs_ref) <= 4e-3, tag


def test_act_conv_fused_matches_unfused_pipeline():
    """Model shapes (stage 4 / 5, one clip): the fused kernel against the activation kernel followed by cuDNN."""
    import torch.nn.functional as F

    _, FC = _fc()
    t32, taps, _ = _taps()
    torch.manual_seed(5)
    for C, T in ((48, 110208), (24, 220416)):
        for (k, d) in ((3, 1), (7, 5), (11, 3)):
            B = 2
            x = torch.randn(B, T, C, device=DEV, dtype=torch.bfloat16)
            res = torch.randn(B, T, C, device=DEV, dtype=torch.bfloat16)
            bias = torch.randn(C, device=DEV) * 0.3
            alpha = torch.randn(C, device=DEV) * 0.5
            beta = torch.randn(C, device=DEV) * 0.5
            conv_w = (torch.randn(C, C, k, device=DEV) / (k * C) ** 0.5).to(torch.bfloat16)
            w_kcc = conv_w.permute(2, 0, 1).contiguous()
            xsum_f = torch.empty_like(x)
            y_f = FC.amp_act_conv_cl(x, T, alpha, beta, taps[0], taps[1], True, w_kcc, k, d, bias=bias, res=res, xsum=xsum_f)
            xsum_u = torch.empty_like(x)
            a_u = FC.amp_activation1d_cl(x, T, alpha, beta, taps[0], taps[1], True, bias=bias, res=res, xsum=xsum_u)
            y_u = F.conv1d(a_u.transpose(1, 2).float(), conv_w.float(), None, 1, (k * d - d) // 2, d).transpose(1, 2)
            assert torch.equal(xsum_f, xsum_u), (C, k, d)
            scale = y_u.abs().max().item()
            err = (y_f.float() - y_u).abs().max().item() / scale
            assert err <= 6e-3, (C, k, d, err)          # bf16 rounding of y (2^-9 relative) on top of identical bf16 tiles
            del x, res, y_f, a_u, y_u
            torch.cuda.empty_cache()


def test_zero_frame_restoration_in_the_tail(amp_golden, true_fp32_convs):
    """inference_e2e.py:140-201 for one clip: zero frames are dropped before the generator and come back as silence.
    The tail kernel scatters hops to their frames (frame_map); must equal the reference's host-side restoration of the
    un-mapped result bit for bit (index work), for equal and for different numbers of kept frames left / right."""
    from afa_b200 import ingest

    c = amp_golden["gen_small_1"]
    h = dict(upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=16, resblock="1",
             resblock_dilation_sizes=[[1, 3, 5]] * 3)
    gen, eng = _engine_from_sd(c["sd"], h, torch.float32)
    hop = 8
    rng = np.random.default_rng(3)
    t_mel = 30
    for zl, zr in (([0, 5, 6, 29], [1, 2, 17, 18]), ([3, 4], [9]), ([], [])):
        ml = (rng.random((80, t_mel)) * 14.5 - 12.0).astype(np.float32)
        mr = (rng.random((80, t_mel)) * 14.5 - 12.0).astype(np.float32)
        ml[:, zl] = 0.0
        mr[:, zr] = 0.0
        pcm = ingest.vocode_binaural(eng, ml, mr).cpu().numpy()
        assert pcm.shape == (t_mel * hop, 2) and pcm.dtype == np.int16
        ref = np.zeros((t_mel * hop, 2), dtype=np.int16)
        kept = [ingest.detect_zero_frames(m) for m in (ml, mr)]
        if kept[0][0].shape[1] == kept[1][0].shape[1]:      # one launch for both channels, as vocode_binaural does
            _, both = eng(torch.tensor(np.stack([kept[0][0], kept[1][0]]), device=DEV), want_pcm=True, pcm_interleave=2)
            monos = [both[0, :, 0].cpu().numpy(), both[0, :, 1].cpu().numpy()]
        else:
            monos = [eng(torch.tensor(f[None], device=DEV), want_pcm=True, pcm_interleave=1)[1][0, :, 0].cpu().numpy()
                     for f, _, _ in kept]
        for ch, (f, mask, idx) in enumerate(kept):
            ref[:, ch] = ingest.restore_silence_host(monos[ch], idx, hop, t_mel * hop)
            assert np.all(ref[np.repeat(mask, hop), ch] == 0)
        assert np.array_equal(pcm, ref), (zl, zr)


def test_act_conv_addend_and_mean_without_residual():
    """The residual add folded into the convolution's epilogue (`x = xt + x`, bigvgan.py:141): y = conv(act(x + bias)) + addend
    with one rounding, on both tensor paths; and the mean kernel with a missing residual entry."""
    from afa_b200 import _lib

    _, FC = _fc()
    t32, taps, taps64 = _taps()
    rng = np.random.default_rng(31)
    for C, k, d, T, B in ((24, 7, 3, 1000, 2), (48, 11, 5, 415, 1), (8, 3, 1, 17, 2), (32, 3, 5, 1733, 1)):
        x = torch.tensor(rng.standard_normal((B, T, C)), dtype=torch.bfloat16, device=DEV)
        r = torch.tensor(rng.standard_normal((B, T, C)), dtype=torch.bfloat16, device=DEV)
        bias = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
        alpha = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
        beta = torch.tensor(rng.standard_normal(C) * 0.5, dtype=torch.float32, device=DEV)
        w = torch.tensor(rng.standard_normal((k, C, C)) / np.sqrt(k * C), dtype=torch.bfloat16, device=DEV)
        _, y_ref = A.amp_act_conv_cl(x.double().cpu().numpy(), alpha.double().cpu().numpy(), beta.double().cpu().numpy(), True,
                                     w.double().cpu().numpy(), d, bias.double().cpu().numpy(), None, taps64, taps64,
                                     round_act=_bf16_round, addend=r.double().cpu().numpy())
        outs = []
        try:
            for path in (1, 0):
                _lib.set_tuning(3, path)
                y = FC.amp_act_conv_cl(x, T, alpha, beta, taps[0], taps[1], True, w, k, d, bias=bias, addend=r)
                torch.cuda.synchronize()
                assert O.max_normalised_error(y.double().cpu().numpy(), y_ref) <= TOL_BF16, (C, k, d, T, path)
                outs.append(y)
        finally:
            _lib.set_tuning(3, 1)
        assert torch.equal(outs[0], outs[1]), (C, k, d, T)
        m = FC.resblock_mean([outs[0], x], [None, r], bias, 0.5)
        m_ref = A.resblock_mean([outs[0].double().cpu().numpy(), x.double().cpu().numpy()], [None, r.double().cpu().numpy()],
                                bias.double().cpu().numpy(), 0.5)
        assert O.max_normalised_error(m.double().cpu().numpy(), m_ref) <= 4e-3


def test_tail_frame_map_out_of_range_is_dropped_not_written():
    """a frame index that is negative or points at / behind T_out must not become an out-of-bounds store.
    Such hops are dropped (the reference clamps every copy to the original length, inference_e2e.py:94-109); the sentinels
    around wave and pcm survive and the in-range frames are what an unmapped call produces."""
    torch.manual_seed(31)
    B, C, hop, frames = 2, 24, 8, 12
    T = hop * frames
    t_out = T                # the C ABI requires T_out >= T; frame indices >= frames are the out-of-range ones
    x = torch.randn(B, T, C, device=DEV)
    a = torch.randn(C, device=DEV) * 0.3
    b = torch.randn(C, device=DEV) * 0.3
    w = torch.randn(C, 7, device=DEV) * 0.1
    F_afa, FC = _fc()
    tu = td = F_afa.host_taps(TP.make_taps())
    fmap = torch.arange(frames, dtype=torch.int32, device=DEV).repeat(B, 1).contiguous()
    fmap[0, 3] = -1          # negative
    fmap[0, 7] = frames      # == T_out / hop: first frame behind the row
    fmap[1, 2] = 1 << 20     # far outside
    guard = 64
    wbuf = torch.full((B * t_out + 2 * guard,), 7.0, device=DEV)
    pbuf = torch.full((B * t_out + 2 * guard,), 77, dtype=torch.int16, device=DEV)
    wave = wbuf[guard : guard + B * t_out].view(B, t_out)
    pcm = pbuf[guard : guard + B * t_out].view(B, t_out, 1)
    wave.zero_()
    pcm.zero_()
    FC.tail_cl(x, T, a, b, tu, td, True, w, None, want_wave=True, want_pcm=True, pcm_interleave=1, wave=wave, pcm=pcm,
               frame_map=fmap, hop=hop, t_out=t_out)
    torch.cuda.synchronize()
    assert torch.all(wbuf[:guard] == 7.0) and torch.all(wbuf[guard + B * t_out :] == 7.0)
    assert torch.all(pbuf[:guard] == 77) and torch.all(pbuf[guard + B * t_out :] == 77)
    ref, _ = FC.tail_cl(x, T, a, b, tu, td, True, w, None, want_wave=True, want_pcm=False)
    for bi in range(B):
        for f in range(frames):
            src = [s for s in range(frames) if int(fmap[bi, s]) == f]
            got = wave[bi, f * hop : (f + 1) * hop]
            if not src:
                assert torch.all(got == 0), (bi, f)                     # silence: nobody mapped here
            elif len(src) == 1:
                assert torch.equal(got, ref[bi, src[0] * hop : (src[0] + 1) * hop]), (bi, f)


def test_generators_match_the_unmodified_reference_on_the_shipped_stage_plan(amp_golden, true_fp32_convs):
    """Whole-generator golden produced by the UNMODIFIED bigvgan.BigVGAN (tests/golden/make_golden_amp.py) with the shipped
    stage plan (six stages, rates 4,4,2,2,2,2, AMPBlock1 x {3, 7, 11}) narrowed to upsample_initial_channel = 192, T_mel = 16.
    Weights are synthesised from (name, shape) on both sides (tests/golden/synth_weights.py).  The fused [B, C, T] generator
    and the channels-last engine must reproduce the reference's fp64 output in fp32; the bf16 engine within bf16 budget."""
    import sys

    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    from synth_weights import synth_state_dict
    from afa_b200.engine import ChannelsLastVocoder
    from afa_b200.vocoder import BINAURAL_22KHZ_80BAND_256X, BigVGANGenerator

    c = amp_golden["gen_model_192"]
    h = dict(BINAURAL_22KHZ_80BAND_256X)
    h.update(upsample_initial_channel=192)
    gen = BigVGANGenerator(h)
    sd = gen.state_dict()
    assert len(sd) == int(c["meta"][2])                           # same state-dict keys as the reference generator
    new = synth_state_dict({k: tuple(v.shape) for k, v in sd.items()})
    assert abs(float(sum(np.abs(v).sum() for v in new.values())) - float(c["param_checksum"][0])) <= 1e-6 * float(c["param_checksum"][0])
    gen.load_state_dict({k: torch.tensor(new[k]) if k in new else v for k, v in sd.items()})
    gen = gen.to(DEV).eval()
    mel = torch.tensor(c["mel"], device=DEV)
    with torch.no_grad():
        y = gen(mel)
    ref = c["y_f64"]
    assert tuple(y.shape) == ref.shape == (2, 1, 16 * 256)
    e_gen = O.max_normalised_error(y.double().cpu().numpy(), ref)
    assert e_gen <= 5e-5, e_gen                                   # 109 fused activations + cuDNN fp32 convolutions in between
    eng = ChannelsLastVocoder(gen, dtype=torch.float32)
    wave, pcm = eng(mel, want_pcm=True)
    e_eng = O.max_normalised_error(wave.double().cpu().numpy(), ref)
    assert e_eng <= 5e-5, e_eng
    ref_pcm = A.pcm_stereo(c["y_f32"][:, 0, :])
    diff = np.abs(pcm.cpu().numpy()[0].astype(np.int32) - ref_pcm.astype(np.int32))
    assert diff.max() <= 1 and (diff == 0).mean() >= 0.95
    eng_b = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
    w_b, _ = eng_b(mel)
    rel_l2 = float(np.linalg.norm(w_b.double().cpu().numpy() - ref) / np.linalg.norm(ref))
    assert rel_l2 <= 5e-2, rel_l2                                 # bf16 storage between ~110 layers (measured 3.5e-2)
    gen_b = BigVGANGenerator(h)
    gen_b.load_state_dict(gen.state_dict())
    gen_b = gen_b.to(DEV).bfloat16().eval()
    with torch.no_grad():
        y_b = gen_b(mel.bfloat16())
    rel_l2 = float(np.linalg.norm(y_b.double().cpu().numpy() - ref) / np.linalg.norm(ref))
    assert rel_l2 <= 8e-2, rel_l2                                 # [B, C, T] generator, cuDNN bf16 convolutions in between (measured 5.1e-2)


def test_batched_ingest_compaction_matches_the_reference_functions(amp_golden, true_fp32_convs):
    """afa_compact_zero_frames against the fixtures produced by the reference's detect_and_exclude_zero_frames
    (bit-exact index work), and BatchedVocoder against per-clip vocode_binaural (itself pinned on the reference's
    reconstruct_audio_with_silence by test_zero_frame_restoration_in_the_tail): PCM bit for bit."""
    from afa_b200 import ingest

    for name in ("zf_some", "zf_none", "zf_all_but_one"):
        c = amp_golden[name]
        mel = torch.tensor(c["mel"], device=DEV)[None].contiguous()
        packed, fmap, n_kept = ingest.compact_zero_frames(mel)
        k = int(n_kept[0])
        assert k == c["filtered"].shape[1], name
        assert np.array_equal(packed[0, :, :k].cpu().numpy(), c["filtered"]), name
        assert np.array_equal(fmap[0, :k].cpu().numpy(), c["nonzero_indices"]), name
        assert torch.all(fmap[0, k:] == -1) and torch.all(packed[0, :, k:] == 0), name
    # edge cases: empty batch, every frame zero, T spanning several 256-frame passes with zero runs across the pass boundaries
    p, f, n = ingest.compact_zero_frames(torch.zeros(0, 80, 16, device=DEV))
    assert p.shape == (0, 80, 16) and n.numel() == 0
    p, f, n = ingest.compact_zero_frames(torch.zeros(3, 80, 700, device=DEV))
    assert torch.all(n == 0) and torch.all(f == -1)
    rng = np.random.default_rng(5)
    big = (rng.random((4, 80, 1000)) * 14.5 - 12.0).astype(np.float32)
    for r, zs in enumerate(([250, 251, 255, 256, 257, 511, 512], list(range(200, 600)), [0, 999], [])):
        big[r][:, zs] = 0.0
    p, f, n = ingest.compact_zero_frames(torch.tensor(big, device=DEV))
    for r in range(4):
        filt, mask, idx = ingest.detect_zero_frames(big[r])
        assert int(n[r]) == filt.shape[1]
        assert np.array_equal(p[r, :, : filt.shape[1]].cpu().numpy(), filt) and np.array_equal(f[r, : filt.shape[1]].cpu().numpy(), idx)

    c = amp_golden["gen_small_1"]
    h = dict(upsample_rates=[4, 2], upsample_kernel_sizes=[8, 4], upsample_initial_channel=16, resblock="1",
             resblock_dilation_sizes=[[1, 3, 5]] * 3)
    gen, eng = _engine_from_sd(c["sd"], h, torch.float32)
    t_mel, clips = 30, 3
    bv = ingest.BatchedVocoder(eng, clips, t_mel)
    mels = (rng.random((clips, 2, 80, t_mel)) * 14.5 - 12.0).astype(np.float32)
    def same_pcm(a, b, tag):
        # the batch (6 rows, cuDNN algorithms picked by timing) and the per-clip call (2 rows) may run different convolution
        # kernels: fp32 results agree to ~1e-6, so a sample next to an integer boundary may truncate differently
        d = np.abs(a.astype(np.int32) - b.astype(np.int32))
        assert d.max() <= 1 and (d == 0).mean() >= 0.95, (tag, int(d.max()), float((d == 0).mean()))      # seen: 0.967 .. 1.0 across boxes

    full = bv(torch.tensor(mels, device=DEV)).cpu().numpy()                      # graphed path, nothing dropped
    for ci in range(clips):
        same_pcm(full[ci], ingest.vocode_binaural(eng, mels[ci, 0], mels[ci, 1]).cpu().numpy(), ci)
    mels[0, 0][:, [0, 5, 6, 29]] = 0.0
    mels[0, 1][:, [1, 2, 17, 18]] = 0.0          # same count left / right
    mels[1, 1][:, [3, 4]] = 0.0                  # left full, right short
    mels[2, 0][:, :] = 0.0                       # a silent channel
    mixed = bv(torch.tensor(mels, device=DEV)).cpu().numpy()
    for ci in range(clips):
        ref_pcm = ingest.vocode_binaural(eng, mels[ci, 0], mels[ci, 1]).cpu().numpy()
        same_pcm(mixed[ci], ref_pcm, ci)
    kept_l, mask_l, _ = ingest.detect_zero_frames(mels[0, 0])
    assert np.all(mixed[0, np.repeat(mask_l, bv.hop), 0] == 0)                # restored silence exactly where the reference puts it
    assert np.all(mixed[2, :, 0] == 0)
