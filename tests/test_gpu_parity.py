"""GPU parity tests: the CUDA path (through the C ABI, via the Python mirror) against the oracle.

Tolerances (BASELINE.json north_star / SURVEY.md section 8d), E = max|y - ref| / max|ref|:
    fp32 forward and gx   : E <= 1e-5   (vs the reference's own fp32 output AND vs fp64 truth)
    bf16 I/O, fp32 math   : E <= 1e-2   (vs fp64 truth evaluated on the bf16-rounded input)
    galpha / gbeta (fp32) : E <= 1e-4   (long reductions)
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import afa_oracle as O
from oracle import torch_path as TP

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 1e-2
TOL_PGRAD = 1e-4
TOL_PGRAD_MASS = 1e-5  # |g - ref| / sum|terms|: the bound that still means something when the sum cancels


def _pgrad_ok(g, ref, mass, tol=TOL_PGRAD):
    """Parameter-gradient check: max-normalised error <= tol, or (for sums that cancel, e.g. one short
    row and one channel) error small against the mass of the reduction."""
    g = np.asarray(g, dtype=np.float64)
    if O.max_normalised_error(g, ref) <= tol:
        return True
    return bool(np.all(np.abs(g - ref) <= TOL_PGRAD_MASS * np.maximum(mass, 1e-30)))


def _mods():
    import afa_b200
    from afa_b200 import _lib, functional
    from afa_b200.activations import Snake, SnakeBeta

    return afa_b200, _lib, functional, Snake, SnakeBeta


def _make(C, kind, logscale, alpha, beta, dev):
    afa_b200, _, _, Snake, SnakeBeta = _mods()
    act = (SnakeBeta if kind == "snakebeta" else Snake)(C, alpha_logscale=logscale)
    with torch.no_grad():
        act.alpha.copy_(torch.as_tensor(alpha))
        if kind == "snakebeta":
            act.beta.copy_(torch.as_tensor(beta))
    return afa_b200.Activation1d(activation=act).to(dev)


def _run(m, x, gy=None):
    x = x.clone().requires_grad_(gy is not None)
    y = m(x)
    if gy is None:
        return y.detach(), None, None, None
    for p in m.parameters():
        p.grad = None
    y.backward(gy)
    gb = m.act.beta.grad if hasattr(m.act, "beta") else None
    return y.detach(), x.grad, m.act.alpha.grad, gb


def test_golden_vectors_fp32(golden, golden_cases):
    """Every vector the unmodified reference produced (tests/golden/make_golden.py), forward and backward."""
    dev = torch.device("cuda:0")
    for name, c in golden_cases.items():
        B, C, T, is_beta, logscale = [int(v) for v in c["meta"]]
        kind = "snakebeta" if is_beta else "snake"
        m = _make(C, kind, bool(logscale), c["alpha"], c.get("beta"), dev)
        np.testing.assert_array_equal(m.upsample.filter.cpu().numpy().reshape(-1), golden["taps_f32"])
        y, gx, ga, gb = _run(m, torch.from_numpy(c["x"]).to(dev), torch.from_numpy(c["gy"]).to(dev))
        for ref in ("y_f32", "y_f64"):
            assert O.max_normalised_error(y.detach().cpu().numpy(), c[ref]) <= TOL_F32, (name, ref)
        for ref in ("gx_f32", "gx_f64"):
            assert O.max_normalised_error(gx.cpu().numpy(), c[ref]) <= TOL_F32, (name, ref)
        assert O.max_normalised_error(ga.cpu().numpy(), c["galpha_f64"]) <= TOL_PGRAD, name
        if is_beta:
            assert O.max_normalised_error(gb.cpu().numpy(), c["gbeta_f64"]) <= TOL_PGRAD, name


EDGE_T = [1, 2, 3, 4, 5, 8, 11, 12, 31, 35, 36, 37, 40, 44, 71, 72, 73, 76, 127, 128, 129, 144, 148, 1000, 3444]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kind,logscale", [("snakebeta", True), ("snake", True), ("snakebeta", False), ("snake", False)])
def test_edge_lengths_forward_backward(dtype, kind, logscale):
    """Ragged / tiny / odd row lengths around every segment boundary (L = 20, 36 fp32; 40, 72 bf16)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1234)
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    for T in EDGE_T:
        for B, C in ((1, 1), (2, 3), (3, 24)):
            if logscale:
                alpha = (torch.randn(C, generator=g) * 0.5).numpy()
                beta = (torch.randn(C, generator=g) * 0.5).numpy()
            else:
                alpha = (torch.rand(C, generator=g) * 2 + 0.25).numpy()
                beta = (torch.rand(C, generator=g) * 2 + 0.25).numpy()
            beta_ = beta if kind == "snakebeta" else None
            x = torch.randn(B, C, T, generator=g).to(dtype)
            gy = torch.randn(B, C, T, generator=g).to(dtype)
            m = _make(C, kind, logscale, alpha, beta, dev)
            y, gx, ga, gb = _run(m, x.to(dev), gy.to(dev))
            xr, gr = x.float().numpy(), gy.float().numpy()
            y_ref = O.activation1d_forward(xr, alpha, beta_, logscale, taps, taps)
            gx_ref, ga_ref, gb_ref = O.activation1d_backward(xr, gr, alpha, beta_, logscale, taps, taps)
            tag = (T, B, C, str(dtype))
            assert y.dtype == dtype and y.shape == (B, C, T)
            assert O.max_normalised_error(y.detach().float().cpu().numpy(), y_ref) <= tol, tag
            assert O.max_normalised_error(gx.detach().float().cpu().numpy(), gx_ref) <= tol, tag
            ma, mb = O.param_grad_mass(xr, gr, alpha, beta_, logscale, taps, taps)
            assert _pgrad_ok(ga.cpu().numpy(), ga_ref, ma), tag
            if beta_ is not None:
                assert _pgrad_ok(gb.cpu().numpy(), gb_ref, mb), tag


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_unaligned_base_pointer_and_noncontiguous(dtype):
    """A view that starts 1 element into an allocation (not 16-byte aligned) and a transposed view."""
    dev = torch.device("cuda:0")
    torch.manual_seed(7)
    C, T = 5, 256
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    alpha, beta = np.full(C, 0.3, np.float32), np.full(C, -0.2, np.float32)
    m = _make(C, "snakebeta", True, alpha, beta, dev)
    buf = torch.randn(2 * C * T + 1, device=dev).to(dtype)
    x = buf[1:].view(2, C, T)
    assert x.data_ptr() % 16 != 0
    y = m(x)
    ref = O.activation1d_forward(x.float().cpu().numpy(), alpha, beta, True, taps, taps)
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    assert O.max_normalised_error(y.detach().float().cpu().numpy(), ref) <= tol
    xt = torch.randn(2, T, C, device=dev).to(dtype).transpose(1, 2)
    assert not xt.is_contiguous()
    yt = m(xt)
    assert yt.is_contiguous()
    ref = O.activation1d_forward(xt.float().cpu().numpy(), alpha, beta, True, taps, taps)
    assert O.max_normalised_error(yt.detach().float().cpu().numpy(), ref) <= tol


def test_large_argument_range_reduction():
    """|alpha*x| in the hundreds: the fast-sine path must stay inside the fp32 budget."""
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    C, T = 4, 4096
    alpha = np.array([1.5, 1.0, 0.5, 0.0], np.float32)
    beta = np.array([0.0, 0.5, -0.5, 1.0], np.float32)
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    m = _make(C, "snakebeta", True, alpha, beta, dev)
    x = torch.randn(2, C, T) * 10.0
    y = m(x.to(dev))
    ref = O.activation1d_forward(x.numpy(), alpha, beta, True, taps, taps)
    assert O.max_normalised_error(y.detach().cpu().numpy(), ref) <= TOL_F32


@pytest.mark.parametrize("shape", [(2, 512, 8192), (2, 768, 3444), (2, 24, 220416), (32, 96, 2048)])
def test_full_size_vs_torch_oracle_on_device(shape):
    """BASELINE.json sizes: compare with the torch-op oracle executed on the same device (fp32), plus the
    DC identity (constant rows pass through as c + sin^2(alpha c)/beta)."""
    dev = torch.device("cuda:0")
    B, C, T = shape
    g = torch.Generator(device="cpu").manual_seed(1234)
    alpha = torch.randn(C, generator=g) * 0.5
    beta = torch.randn(C, generator=g) * 0.5
    m = _make(C, "snakebeta", True, alpha.numpy(), beta.numpy(), dev)
    x = torch.randn(B, C, T, generator=g).to(dev)
    y = m(x)
    taps = m.upsample.filter
    with torch.no_grad():
        ref = TP.activation1d_torch(x, alpha.to(dev), beta.to(dev), True, taps, taps)
    err = (y - ref).abs().max().item() / ref.abs().max().item()
    assert err <= TOL_F32, err
    # backward at full size against autograd of the oracle
    gy = torch.randn(B, C, T, generator=g).to(dev)
    xg = x.clone().requires_grad_(True)
    m.zero_grad()
    m(xg).backward(gy)
    gx_ref, ga_ref, gb_ref = TP.activation1d_torch_grads(x, gy, alpha.to(dev), beta.to(dev), True, taps, taps)
    assert ((xg.grad - gx_ref).abs().max() / gx_ref.abs().max()).item() <= TOL_F32
    assert ((m.act.alpha.grad - ga_ref).abs().max() / ga_ref.abs().max()).item() <= TOL_PGRAD
    assert ((m.act.beta.grad - gb_ref).abs().max() / gb_ref.abs().max()).item() <= TOL_PGRAD
    # DC identity
    c = torch.linspace(-2, 2, C, device=dev).view(1, C, 1).expand(B, C, T).contiguous()
    yc = m(c)
    expect = c + torch.sin(torch.exp(alpha.to(dev)).view(1, C, 1) * c) ** 2 / (torch.exp(beta.to(dev)).view(1, C, 1) + 1e-9)
    assert (yc - expect).abs().max().item() <= 2e-6 * max(1.0, expect.abs().max().item())


def test_segmentation_invariance_bitwise():
    """The per-sample arithmetic does not depend on how rows are cut into segments: every compiled
    segment size, and the aligned (TMA) vs unaligned (scalar staging) kernels, agree bit for bit."""
    _, _lib, _, _, _ = _mods()
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    C, T = 6, 1024
    alpha, beta = np.linspace(-0.5, 0.5, C).astype(np.float32), np.linspace(0.4, -0.4, C).astype(np.float32)
    m = _make(C, "snakebeta", True, alpha, beta, dev)
    for dtype in (torch.float32, torch.bfloat16):
        big = torch.randn(2 * C * T + 8, device=dev).to(dtype)
        x_al = big[: 2 * C * T].view(2, C, T)
        x_un = big[1 : 2 * C * T + 1].view(2, C, T)
        x_un.copy_(x_al.clone())
        x_al = x_un.clone()
        outs = []
        for which_x in (x_al, x_un):
            for ch in (5, 9):
                _lib.set_tuning(0, ch, 0)
                outs.append(m(which_x).clone())
        _lib.set_tuning(0, 0, 0)
        for o in outs[1:]:
            assert torch.equal(o, outs[0])


def test_cuda_graph_capture_and_streams():
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    C, T = 24, 2048
    m = _make(C, "snakebeta", True, np.zeros(C, np.float32), np.zeros(C, np.float32), dev)
    x = torch.randn(2, C, T, device=dev)
    y_ref = m(x).clone()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        y_side = m(x)
    s.synchronize()
    assert torch.equal(y_side, y_ref)
    static_x = x.clone()
    g = torch.cuda.CUDAGraph()
    m(static_x)  # warm-up (host tap cache, attribute setup) outside capture
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        static_y = m(static_x)
    static_x.copy_(x * 0.5)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_y, m(x * 0.5))


def test_c_abi_direct_call_and_errors():
    """Straight through ctypes, no Python mirror: argument errors come back as codes + messages."""
    _, _lib, functional, _, _ = _mods()
    lib = _lib.load_library()
    dev = torch.device("cuda:0")
    B, C, T = 2, 3, 100
    x = torch.randn(B, C, T, device=dev)
    y = torch.empty_like(x)
    a = torch.zeros(C, device=dev)
    taps = functional.host_taps(TP.make_taps())
    rc = lib.afa_activation1d_fwd(x.data_ptr(), y.data_ptr(), a.data_ptr(), a.data_ptr(), taps, taps, B, C, T, 0, 1, None)
    assert rc == 0
    torch.cuda.synchronize()
    ref = O.activation1d_forward(x.cpu().numpy(), np.zeros(C), np.zeros(C), True)
    assert O.max_normalised_error(y.detach().cpu().numpy(), ref) <= TOL_F32
    assert lib.afa_activation1d_fwd(x.data_ptr(), x.data_ptr(), a.data_ptr(), a.data_ptr(), taps, taps, B, C, T, 0, 1, None) == -1
    assert b"alias" in lib.afa_last_error()
    assert lib.afa_activation1d_fwd(x.data_ptr(), y.data_ptr(), a.data_ptr(), None, taps, taps, B, C, T, 0, 1, None) == -1
    assert lib.afa_activation1d_fwd(x.data_ptr(), y.data_ptr(), a.data_ptr(), a.data_ptr(), taps, taps, B, C, T, 7, 1, None) == -2
    ws = torch.empty(8, dtype=torch.uint8, device=dev)
    rc = lib.afa_activation1d_bwd(x.data_ptr(), x.data_ptr(), y.data_ptr(), a.data_ptr(), a.data_ptr(), a.data_ptr(),
                                  a.data_ptr(), taps, taps, B, C, T, 0, 1, ws.data_ptr(), 8, None)
    assert rc == -4 and b"workspace" in lib.afa_last_error()
    # empty tensors are a no-op
    assert lib.afa_activation1d_fwd(x.data_ptr(), y.data_ptr(), a.data_ptr(), a.data_ptr(), taps, taps, 0, C, T, 0, 1, None) == 0
    assert lib.afa_activation1d_fwd(x.data_ptr(), y.data_ptr(), a.data_ptr(), a.data_ptr(), taps, taps, B, C, 0, 0, 1, None) == 0
    info = _lib.kernel_info(0, 0, 8192)
    assert info["registers"] > 0 and info["ctas_per_sm"] >= 1
    assert _lib.launch_count() >= 1


def test_module_semantics_on_device():
    afa_b200, _, _, Snake, SnakeBeta = _mods()
    dev = torch.device("cuda:0")
    C = 8
    m = afa_b200.Activation1d(activation=SnakeBeta(C, alpha_logscale=True)).to(dev)
    assert sorted(m.state_dict().keys()) == ["act.alpha", "act.beta", "downsample.lowpass.filter", "upsample.filter"]
    x = torch.randn(2, C, 64, device=dev)
    with torch.no_grad():
        y0 = m(x)
    with torch.inference_mode():
        y1 = m(x)
    assert torch.equal(y0, y1)
    with pytest.raises(ValueError):
        m(torch.randn(C, 64, device=dev))
    with pytest.raises(RuntimeError):
        m(torch.randn(2, C + 1, 64, device=dev))
    with pytest.raises(RuntimeError):
        m(torch.randn(2, C, 64))
    with pytest.raises(TypeError):
        m(x.double())
    # whole-module bf16 cast (cfg 3: generator.bfloat16()): parameters are read back in fp32 inside the op
    mb = afa_b200.Activation1d(activation=SnakeBeta(C, alpha_logscale=True)).to(dev).bfloat16()
    yb = mb(x.bfloat16())
    assert yb.dtype == torch.bfloat16
    assert O.max_normalised_error(yb.detach().float().cpu().numpy(), y0.detach().cpu().numpy()) <= TOL_BF16
    # Snake: single parameter receives the summed gradient
    ms = afa_b200.Activation1d(activation=Snake(C, alpha_logscale=False)).to(dev)
    xs = x.clone().requires_grad_(True)
    ms(xs).sum().backward()
    assert ms.act.alpha.grad is not None and xs.grad is not None


def test_ddp_wrapped_module_gets_all_gradients():
    """train_binaural_mel.py:540-543 wraps the generator in DDP with find_unused_parameters=False: the fused op
    must hand autograd a gradient for act.alpha and act.beta (single-process NCCL group here)."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    afa_b200, _, _, _, SnakeBeta = _mods()
    dev = torch.device("cuda:0")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29533", rank=0, world_size=1, device_id=dev)
        created = True
    try:
        C = 16
        m = afa_b200.Activation1d(activation=SnakeBeta(C, alpha_logscale=True)).to(dev)
        ddp = DDP(m, device_ids=[0], find_unused_parameters=False)
        x = torch.randn(4, C, 512, device=dev, requires_grad=True)
        ddp(x).square().mean().backward()
        assert m.act.alpha.grad is not None and m.act.beta.grad is not None and x.grad is not None
        assert torch.isfinite(m.act.alpha.grad).all() and m.act.alpha.grad.abs().sum() > 0
        torch.nn.utils.clip_grad_norm_(m.parameters(), 500)              # train_binaural_mel.py:788-790
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_no_out_of_bounds_writes(dtype):
    """compute-sanitizer is closed on this pool, so guard bands do its job for writes: outputs live inside a
    larger sentinel-filled allocation (16-byte aligned, 8-byte aligned and unaligned placements) and the
    sentinels must survive forward and backward."""
    _, _lib, Fn, _, _ = _mods()
    dev = torch.device("cuda:0")
    torch.manual_seed(21)
    taps = Fn.host_taps(TP.make_taps())
    guard = 64
    es = 4 if dtype == torch.float32 else 2
    for (B, C, T) in [(1, 1, 1), (2, 3, 4), (3, 5, 12), (2, 3, 37), (1, 7, 100), (2, 24, 3444), (5, 3, 2300)]:
        n = B * C * T
        a = torch.randn(C, device=dev) * 0.5
        b = torch.randn(C, device=dev) * 0.5
        x = torch.randn(B, C, T, device=dev).to(dtype)
        gy = torch.randn(B, C, T, device=dev).to(dtype)
        for shift in (0, 8 // es, 1):                       # 16-byte aligned, 8-byte aligned, element aligned
            big = torch.full((n + 2 * guard + 8,), 12345.0, device=dev, dtype=dtype)
            off = guard + shift
            y = big[off : off + n].view(B, C, T)
            Fn.activation1d_forward_raw(x, a, b, taps, taps, True, out=y)
            torch.cuda.synchronize()
            assert torch.all(big[:off] == 12345.0) and torch.all(big[off + n :] == 12345.0), (B, C, T, shift, "fwd")
            assert torch.isfinite(y.float()).all()
        gx, ga, gb = Fn.activation1d_backward_raw(x, gy, a, b, taps, taps, True)
        assert torch.isfinite(gx.float()).all() and torch.isfinite(ga).all() and torch.isfinite(gb).all()


def test_backward_is_run_to_run_deterministic():
    """Two-stage reduction, no atomics: the parameter gradients (and gx) are bit-identical across runs, which
    DDP-averaged training (train_binaural_mel.py:540-543) can rely on."""
    _, _, Fn, _, _ = _mods()
    dev = torch.device("cuda:0")
    torch.manual_seed(9)
    taps = Fn.host_taps(TP.make_taps())
    for dtype in (torch.float32, torch.bfloat16):
        for (B, C, T) in [(4, 24, 8192), (32, 96, 2048), (2, 768, 3444)]:
            x = torch.randn(B, C, T, device=dev).to(dtype)
            gy = torch.randn(B, C, T, device=dev).to(dtype)
            a = torch.randn(C, device=dev) * 0.5
            b = torch.randn(C, device=dev) * 0.5
            first = Fn.activation1d_backward_raw(x, gy, a, b, taps, taps, True)
            for _ in range(3):
                again = Fn.activation1d_backward_raw(x, gy, a, b, taps, taps, True)
                assert all(torch.equal(p, q) for p, q in zip(first, again))


# ------------------------------------------------------------------------------------------------
# tensor-core forward (csrc/afa_tc_kernels.cuh): bf16 tensors with T % 8 == 0 and 16-byte aligned rows
# ------------------------------------------------------------------------------------------------
def _tc_forward(x, alpha, beta, logscale, mode, ny=0, rlog2=-1):
    _, _lib, Fn, _, _ = _mods()
    taps = Fn.host_taps(TP.make_taps())
    _lib.set_tuning(5, mode, ny)
    _lib.set_tuning(6, rlog2, 0)
    try:
        n0 = _lib.launch_count()
        y = Fn.activation1d_forward_raw(x, alpha, beta, taps, taps, logscale)
        torch.cuda.synchronize()
        assert _lib.launch_count() == n0 + 1
    finally:
        _lib.set_tuning(5, 1, 0)
        _lib.set_tuning(6, -1, 0)
    return y


TC_EDGE = [(1, 8, 64), (1, 8, 72), (2, 3, 128), (3, 5, 1000), (1, 24, 8), (2, 24, 256), (1, 128, 264), (1, 130, 512),
           (2, 24, 2040), (5, 7, 4104), (1, 16, 16), (2, 12, 24), (1, 9, 40),
           (2, 6, 1040), (1, 10, 2072), (1, 4, 1936),      # T % 32 = 16 / 24 / 16: the row end at every position (element 10, 26, 42, 58) of a 64-value block
           # T % 8 == 4 with an even row count: rows travel as 16-byte aligned PAIRS (second row shifted 4 samples, edge chunks written by
           # the lanes); T % 32 = 4, 12, 20, 28 puts the row end at elements 2, 18, 34, 50
           (1, 8, 68), (2, 7, 100), (4, 5, 260), (3, 6, 1004), (2, 24, 2052), (1, 130, 516), (5, 8, 4100), (1, 16, 2060), (2, 4, 1972), (1, 12, 3444)]


@pytest.mark.parametrize("kind,logscale", [("snakebeta", True), ("snake", True), ("snakebeta", False), ("snake", False)])
def test_tensor_core_forward_edge_grid(kind, logscale):
    """Rows that are no multiple of the CTA's row count, rows shorter than one CTA span, T = 8 ... around the 64-sample chunk
    and 256-sample strip boundaries, every blocks-per-lane and rows-per-CTA variant, against the float64 oracle on the
    bf16-rounded input (1e-2) and against the register-walk kernel (both round the same fp32 math to bf16: <= 2 bf16 ulps)."""
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4321)
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    for (B, C, T) in TC_EDGE:
        if logscale:
            alpha = (torch.randn(C, generator=g) * 0.5)
            beta = (torch.randn(C, generator=g) * 0.5)
        else:
            alpha = (torch.rand(C, generator=g) * 2 + 0.25)
            beta = (torch.rand(C, generator=g) * 2 + 0.25)
        beta_ = beta if kind == "snakebeta" else None
        x = torch.randn(B, C, T, generator=g).to(torch.bfloat16).to(dev)
        y_ref = O.activation1d_forward(x.float().cpu().numpy(), alpha.numpy(), None if beta_ is None else beta_.numpy(), logscale, taps, taps)
        a_d, b_d = alpha.to(dev), None if beta_ is None else beta_.to(dev)
        y_walk = _tc_forward(x, a_d, b_d, logscale, 0)
        for ny, rlog2 in ((0, -1), (4, 3), (8, 4), (12, 5), (16, 6), (4, 7), (16, 3)):
            if T < 64 and ny == 0 and rlog2 == -1:
                continue                                   # the built-in choice keeps such rows on the walk kernel
            y = _tc_forward(x, a_d, b_d, logscale, 2, ny, rlog2)
            tag = (B, C, T, ny, rlog2)
            assert y.dtype == torch.bfloat16 and y.shape == (B, C, T)
            assert torch.isfinite(y.float()).all(), tag
            assert O.max_normalised_error(y.float().cpu().numpy(), y_ref) <= TOL_BF16, tag
            assert (y.float() - y_walk.float()).abs().max().item() <= 4 * 2.0 ** -8 * max(1.0, float(np.abs(y_ref).max())), tag


@pytest.mark.parametrize("shape", [(2, 24, 220416), (16, 768, 3440), (16, 384, 13776), (32, 96, 2048), (2, 512, 8192), (16, 24, 220416),
                                   (16, 768, 3444), (2, 768, 3444)])
def test_tensor_core_forward_model_sizes(shape):
    """bf16 at the sizes the bench runs (the full-size test was fp32 only): against the torch-op oracle in
    fp32 on the same device evaluated on the bf16 input, bitwise run-to-run determinism, and the DC identity."""
    _, _lib, Fn, _, _ = _mods()
    dev = torch.device("cuda:0")
    B, C, T = shape
    g = torch.Generator(device="cpu").manual_seed(1234)
    alpha = (torch.randn(C, generator=g) * 0.5).to(dev)
    beta = (torch.randn(C, generator=g) * 0.5).to(dev)
    x = torch.randn(B, C, T, generator=g).to(torch.bfloat16).to(dev)
    taps_t = TP.make_taps().to(dev)
    y = _tc_forward(x, alpha, beta, True, 2)
    with torch.no_grad():
        ref = TP.activation1d_torch(x.float(), alpha, beta, True, taps_t, taps_t)
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= TOL_BF16, err
    # error statistics, not only the maximum: the mean error must sit at bf16 output-rounding level
    assert ((y.float() - ref).abs().mean() / ref.abs().mean()).item() <= 4e-3
    for _ in range(2):
        assert torch.equal(_tc_forward(x, alpha, beta, True, 2), y)
    c = torch.linspace(-2, 2, C, device=dev).view(1, C, 1).expand(B, C, T).contiguous().to(torch.bfloat16)
    yc = _tc_forward(c, alpha, beta, True, 2).float()
    cf = c.float()
    expect = cf + torch.sin(torch.exp(alpha).view(1, C, 1) * cf) ** 2 / (torch.exp(beta).view(1, C, 1) + 1e-9)
    assert ((yc - expect).abs().max() / expect.abs().max()).item() <= TOL_BF16


def test_tensor_core_forward_large_arguments_and_degenerate_parameters():
    """|alpha * x| in the hundreds (x * 10, alpha up to e^1.5), and non-log-scale alpha in {0, -0.5} / beta -> 0
    (raw parameters are only guarded by + 1e-9)."""
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    C, T = 8, 4096
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    x = (torch.randn(2, C, T) * 10.0).to(torch.bfloat16).to(dev)
    alpha = torch.tensor([1.5, 1.0, 0.5, 0.0, -0.5, 1.2, 0.3, -1.0])
    beta = torch.tensor([0.0, 0.5, -0.5, 1.0, 0.2, -0.3, 1.5, 0.7])
    y = _tc_forward(x, alpha.to(dev), beta.to(dev), True, 2)
    ref = O.activation1d_forward(x.float().cpu().numpy(), alpha.numpy(), beta.numpy(), True, taps, taps)
    assert O.max_normalised_error(y.float().cpu().numpy(), ref) <= TOL_BF16
    x1 = torch.randn(2, C, T).to(torch.bfloat16).to(dev)
    alpha = torch.tensor([0.0, -0.5, 1.0, 2.0, 0.5, -2.0, 0.1, 3.0])
    beta = torch.tensor([1.0, 0.5, 1e-3, 2.0, -0.5, 0.25, 1e-2, 4.0])
    for b_ in (beta, None):
        y = _tc_forward(x1, alpha.to(dev), None if b_ is None else b_.to(dev), False, 2)
        a_np, b_np = alpha.numpy(), None if b_ is None else b_.numpy()
        ref = O.activation1d_forward(x1.float().cpu().numpy(), a_np, b_np, False, taps, taps)
        # per channel: beta = 1e-3 scales the sin^2 term by 1000, so normalise each channel by its own maximum
        yv, rv = y.float().cpu().numpy(), ref
        for ch in range(C):
            assert O.max_normalised_error(yv[:, ch], rv[:, ch]) <= TOL_BF16, (ch, b_ is None)


def test_tensor_core_forward_graph_capture_guard_band_and_module_dispatch():
    """CUDA-graph capture (tensor maps travel by value), no write outside y (sentinels), and the module picks the tensor-core
    kernel by itself for large bf16 tensors (launch is one kernel either way; results within 2 ulps of the walk kernel)."""
    _, _lib, Fn, _, _ = _mods()
    dev = torch.device("cuda:0")
    torch.manual_seed(8)
    taps = Fn.host_taps(TP.make_taps())
    B, C, T = 3, 20, 5000 // 8 * 8
    a = torch.randn(C, device=dev) * 0.5
    b = torch.randn(C, device=dev) * 0.5
    x = torch.randn(B, C, T, device=dev).to(torch.bfloat16)
    n = B * C * T
    big = torch.full((n + 256,), 12345.0, device=dev, dtype=torch.bfloat16)
    y = big[128 : 128 + n].view(B, C, T)
    _lib.set_tuning(5, 2, 0)
    try:
        Fn.activation1d_forward_raw(x, a, b, taps, taps, True, out=y)
        torch.cuda.synchronize()
        assert torch.all(big[:128] == 12345.0) and torch.all(big[128 + n :] == 12345.0)
        y0 = y.clone()
        sx = x.clone()
        sy = torch.empty_like(x)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            Fn.activation1d_forward_raw(sx, a, b, taps, taps, True, out=sy)
        sx.copy_(x)
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(sy, y0)
        info = _lib.kernel_info(5, 1, T)
        assert info["registers"] > 0 and info["ctas_per_sm"] >= 1 and info["threads"] == 320
    finally:
        _lib.set_tuning(5, 1, 0)


def test_tensor_core_dependent_launch_chain_is_bit_identical():
    """Programmatic dependent launch (afa_set_tuning(9, ...)): each kernel of a chain reads the tensor the kernel in front of it
    on the stream wrote (y1 = a(x), y2 = a(y1), ...), eagerly and as one CUDA graph; its set-up may overlap the predecessor's
    tail, its reads may not.  Same bits with the attribute on and off, for the [B, C, T] kernel and the channels-last one."""
    _, _lib, Fn, _, _ = _mods()
    from afa_b200 import functional_cl as FC

    dev = torch.device("cuda:0")
    torch.manual_seed(21)
    taps = Fn.host_taps(TP.make_taps())
    B, C, T = 2, 96, 8192
    a = torch.randn(C, device=dev) * 0.3
    b = torch.randn(C, device=dev) * 0.3 + 1.0
    x = (torch.randn(B, C, T, device=dev) * 0.5).to(torch.bfloat16)
    xcl = x.transpose(1, 2).contiguous()
    bias = torch.randn(C, device=dev) * 0.1

    def chain(n=6):
        bufs = [torch.empty_like(x) for _ in range(2)]
        cur = x
        for i in range(n):
            Fn.activation1d_forward_raw(cur, a, b, taps, taps, True, out=bufs[i % 2])
            cur = bufs[i % 2]
        return cur.clone()

    def chain_cl(n=6):
        bufs = [torch.empty_like(xcl) for _ in range(2)]
        cur = xcl
        for i in range(n):
            FC.amp_activation1d_cl(cur, T, a, b, taps, taps, True, bias=bias, out=bufs[i % 2])
            cur = bufs[i % 2]
        return cur.clone()

    def graphed(fn):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay()
        torch.cuda.synchronize()
        return out.clone()

    _lib.set_tuning(5, 2, 0)
    _lib.set_tuning(7, 2, 0)
    try:
        res = {}
        for pdl in (0, 1):
            _lib.set_tuning(9, pdl)
            before = _lib.launch_count()
            res[pdl] = (chain(), chain_cl(), graphed(chain), graphed(chain_cl))
            torch.cuda.synchronize()
            assert _lib.launch_count() == before + 6 * 6, pdl      # one kernel per call: both tensor-core kernels were taken
        for u, v in zip(res[0], res[1]):
            assert torch.equal(u, v)
        assert torch.equal(res[1][0], res[1][2]) and torch.equal(res[1][1], res[1][3])
        assert torch.isfinite(res[1][0].float()).all() and torch.isfinite(res[1][1].float()).all()
    finally:
        _lib.set_tuning(9, 1)
        _lib.set_tuning(5, 1, 0)
        _lib.set_tuning(7, 1, 0)


# ------------------------------------------------------------------------------------------------
# round-2 parity gaps ("What's weak" 1)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 24, 220416), (16, 768, 3444), (16, 384, 13776), (32, 96, 2048)])
def test_bf16_forward_backward_at_model_sizes(shape):
    """bf16 forward AND backward at the sizes the bench runs: T = 3444 takes the half-aligned TMA kernels, 13776 / 220416 the
    CH = 13 / 17 variants, (16, 384, 13776) and (32, 96, 2048) the tensor-core forward.  Reference: the torch-op oracle in
    fp32 on the same device, evaluated on the bf16-rounded tensors."""
    dev = torch.device("cuda:0")
    B, C, T = shape
    g = torch.Generator(device="cpu").manual_seed(99)
    alpha = (torch.randn(C, generator=g) * 0.5).to(dev)
    beta = (torch.randn(C, generator=g) * 0.5).to(dev)
    m = _make(C, "snakebeta", True, alpha.cpu().numpy(), beta.cpu().numpy(), dev)
    x = torch.randn(B, C, T, generator=g).to(torch.bfloat16).to(dev)
    gy = torch.randn(B, C, T, generator=g).to(torch.bfloat16).to(dev)
    y, gx, ga, gb = _run(m, x, gy)
    taps = m.upsample.filter
    with torch.no_grad():
        ref = TP.activation1d_torch(x.float(), alpha, beta, True, taps, taps)
    assert ((y.float() - ref).abs().max() / ref.abs().max()).item() <= TOL_BF16
    gx_ref, ga_ref, gb_ref = TP.activation1d_torch_grads(x.float(), gy.float(), alpha, beta, True, taps, taps)
    assert ((gx.float() - gx_ref).abs().max() / gx_ref.abs().max()).item() <= TOL_BF16
    # parameter gradients: fp32 reductions of bf16-rounded inputs (the kernel reads the same bf16 x / gy as the oracle)
    assert ((ga - ga_ref).abs().max() / ga_ref.abs().max()).item() <= 1e-3
    assert ((gb - gb_ref).abs().max() / gb_ref.abs().max()).item() <= 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_large_argument_backward(dtype):
    """The backward evaluates sin / cos of 2 * alpha * u, twice the forward's argument: x * 10 with alpha up to e^1.5 puts it in
    the hundreds.  Reference: float64 oracle."""
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    C, T = 4, 4096
    alpha = np.array([1.5, 1.0, 0.5, 0.0], np.float32)
    beta = np.array([0.0, 0.5, -0.5, 1.0], np.float32)
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    m = _make(C, "snakebeta", True, alpha, beta, dev)
    x = (torch.randn(2, C, T) * 10.0).to(dtype)
    gy = torch.randn(2, C, T).to(dtype)
    y, gx, ga, gb = _run(m, x.to(dev), gy.to(dev))
    xr, gr = x.float().numpy(), gy.float().numpy()
    gx_ref, ga_ref, gb_ref = O.activation1d_backward(xr, gr, alpha, beta, True, taps, taps)
    if dtype == torch.float32:
        # fast sine / cosine at |argument| ~ 1e3: absolute error ~ |argument| * 2^-23, times alpha * ib * |gy|
        assert O.max_normalised_error(gx.cpu().numpy(), gx_ref) <= 2e-4
        assert O.max_normalised_error(ga.cpu().numpy(), ga_ref) <= 2e-3
        assert O.max_normalised_error(gb.cpu().numpy(), gb_ref) <= 1e-4
    else:
        assert O.max_normalised_error(gx.float().cpu().numpy(), gx_ref) <= TOL_BF16
        assert O.max_normalised_error(ga.cpu().numpy(), ga_ref) <= 2e-3
        assert O.max_normalised_error(gb.cpu().numpy(), gb_ref) <= 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kind", ["snakebeta", "snake"])
def test_non_logscale_degenerate_parameters(kind, dtype):
    """without alpha_logscale the raw parameters are used as they are, guarded only by + 1e-9:
    alpha in {0, -0.5}, beta -> 0 (1e-3, 1e-6) and negative beta must match the reference's arithmetic, forward and backward."""
    dev = torch.device("cuda:0")
    torch.manual_seed(17)
    C, T = 6, 1000
    alpha = np.array([0.0, -0.5, 1.0, 2.0, 0.5, -2.0], np.float32)
    beta = np.array([1.0, 0.5, 1e-3, 1e-6, -0.5, 0.25], np.float32)
    if kind == "snake":                       # beta := alpha: alpha = 0 divides by 1e-9 but multiplies sin^2(0) = 0
        beta_ = None
    else:
        beta_ = beta
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    m = _make(C, kind, False, alpha, beta, dev)
    x = torch.randn(2, C, T).to(dtype)
    gy = torch.randn(2, C, T).to(dtype)
    y, gx, ga, gb = _run(m, x.to(dev), gy.to(dev))
    xr, gr = x.float().numpy(), gy.float().numpy()
    y_ref = O.activation1d_forward(xr, alpha, beta_, False, taps, taps)
    gx_ref, ga_ref, gb_ref = O.activation1d_backward(xr, gr, alpha, beta_, False, taps, taps)
    tol = TOL_F32 if dtype == torch.float32 else TOL_BF16
    yv, gxv = y.float().cpu().numpy(), gx.float().cpu().numpy()
    for ch in range(C):                        # 1 / beta spans nine orders of magnitude: normalise per channel
        assert O.max_normalised_error(yv[:, ch], y_ref[:, ch]) <= tol, (ch, "y")
        assert O.max_normalised_error(gxv[:, ch], gx_ref[:, ch]) <= tol, (ch, "gx")
    ptol = 1e-4 if dtype == torch.float32 else 2e-3
    for ch in range(C):
        scale = max(abs(float(ga_ref[ch])), 1e-30)
        if kind == "snake" and alpha[ch] == 0.0:
            continue                           # d/dalpha of sin^2(alpha u) / (alpha + 1e-9) at alpha = 0: 0 * 1e9 and 1e18 * 0 terms
        assert abs(float(ga[ch]) - float(ga_ref[ch])) <= ptol * scale + 1e-6, (ch, "galpha", float(ga[ch]), float(ga_ref[ch]))
        if beta_ is not None:
            scale = max(abs(float(gb_ref[ch])), 1e-30)
            assert abs(float(gb[ch]) - float(gb_ref[ch])) <= ptol * scale + 1e-6, (ch, "gbeta", float(gb[ch]), float(gb_ref[ch]))


def test_bf16_backward_unaligned_and_half_aligned_workspace():
    """afa_bwd_workspace_bytes plans without pointers; a bf16 tensor with T % 8 == 4 at a base that is not
    16-byte aligned selects a shorter-segment kernel than the aligned plan and needs more partial sums.  The query is an upper
    bound now, so the backward of such views runs (and matches the oracle)."""
    _, _lib, Fn, _, _ = _mods()
    dev = torch.device("cuda:0")
    torch.manual_seed(23)
    taps_h = Fn.host_taps(TP.make_taps())
    taps = TP.make_taps().reshape(-1).numpy().astype(np.float64)
    B, C, T = 2, 5, 3444
    a = (torch.randn(C) * 0.5)
    b = (torch.randn(C) * 0.5)
    for shift in (0, 4, 1):
        bufx = torch.randn(B * C * T + 8).to(torch.bfloat16).to(dev)
        bufg = torch.randn(B * C * T + 8).to(torch.bfloat16).to(dev)
        x = bufx[shift : shift + B * C * T].view(B, C, T)
        gy = bufg[shift : shift + B * C * T].view(B, C, T)
        gx, ga, gb = Fn.activation1d_backward_raw(x, gy, a.to(dev), b.to(dev), taps_h, taps_h, True)
        gx_ref, ga_ref, gb_ref = O.activation1d_backward(x.float().cpu().numpy(), gy.float().cpu().numpy(), a.numpy(), b.numpy(), True, taps, taps)
        assert O.max_normalised_error(gx.float().cpu().numpy(), gx_ref) <= TOL_BF16, shift
        assert O.max_normalised_error(ga.cpu().numpy(), ga_ref) <= 1e-3, shift
        assert O.max_normalised_error(gb.cpu().numpy(), gb_ref) <= 1e-3, shift


def test_pitched_rows_without_a_copy():
    """afa_activation1d_fwd_pitched: a time slice x[:, :, :T] of a longer buffer (rows `pitch` elements apart) gives bit for bit
    what the dense copy gives on the same (tensor-core) kernel; the module takes the view as it is in inference; layouts the
    kernel cannot take report AFA_ERR_ALIGNMENT through the C ABI and fall back to a copy in the Python shim; the buffer
    outside the slice is not touched and nothing is written outside y."""
    afa_b200, _lib, Fn, _, SnakeBeta = _mods()
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    taps = Fn.host_taps(TP.make_taps())
    lib = _lib.load_library()
    for (B, C, T, pitch) in ((2, 24, 4096, 4104), (3, 7, 1024, 2048), (1, 130, 512, 520), (2, 5, 264, 272)):
        buf = torch.randn(B, C, pitch, device=dev).to(torch.bfloat16)
        keep = buf.clone()
        x = buf[:, :, :T]
        assert not x.is_contiguous() and Fn._row_pitch(x) == pitch
        a = torch.randn(C, device=dev) * 0.5
        b = torch.randn(C, device=dev) * 0.5
        n0 = _lib.launch_count()
        y = Fn.activation1d_forward_raw(x, a, b, taps, taps, True)
        torch.cuda.synchronize()
        assert _lib.launch_count() == n0 + 1 and y.is_contiguous() and y.shape == (B, C, T)
        _lib.set_tuning(5, 2, 0)
        try:
            y_dense = Fn.activation1d_forward_raw(x.contiguous(), a, b, taps, taps, True)
        finally:
            _lib.set_tuning(5, 1, 0)
        assert torch.equal(y, y_dense), (B, C, T, pitch)
        assert torch.equal(buf, keep)
        # pitched OUTPUT as well, straight through the C ABI, inside a sentinel-filled buffer
        ybuf = torch.full((B, C, pitch), 777.0, device=dev, dtype=torch.bfloat16)
        rc = lib.afa_activation1d_fwd_pitched(x.data_ptr(), pitch, ybuf.data_ptr(), pitch, a.data_ptr(), b.data_ptr(), taps, taps,
                                              B, C, T, 1, 1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert rc == 0, lib.afa_last_error()
        assert torch.equal(ybuf[:, :, :T], y) and torch.all(ybuf[:, :, T:] == 777.0)
    # layouts outside the contract: an error code from the C ABI, a dense copy in the shim
    x32 = torch.randn(2, 8, 520, device=dev)[:, :, :512]
    a = torch.zeros(8, device=dev)
    rc = lib.afa_activation1d_fwd_pitched(x32.data_ptr(), 520, torch.empty(2, 8, 512, device=dev).data_ptr(), 512, a.data_ptr(), a.data_ptr(),
                                          taps, taps, 2, 8, 512, 0, 1, torch.cuda.current_stream().cuda_stream)
    assert rc == _lib.AFA_ERR_ALIGNMENT and b"tensor-core" in lib.afa_last_error()
    odd = torch.randn(2, 8, 523, device=dev).to(torch.bfloat16)[:, :, :512]           # pitch not a multiple of 8
    y = Fn.activation1d_forward_raw(odd, a, a, taps, taps, True)
    assert torch.equal(y, Fn.activation1d_forward_raw(odd.contiguous(), a, a, taps, taps, True))
    # module, inference: the view goes in as it is; training still saves a dense x
    m = afa_b200.Activation1d(activation=SnakeBeta(24, alpha_logscale=True)).to(dev).to(torch.bfloat16)
    buf = torch.randn(2, 24, 2056, device=dev).to(torch.bfloat16)
    with torch.no_grad():
        y1 = m(buf[:, :, :2048])
        y2 = m(buf[:, :, :2048].contiguous())
    assert (y1.float() - y2.float()).abs().max().item() <= 2 * 2.0 ** -8 * max(1.0, y2.float().abs().max().item())
