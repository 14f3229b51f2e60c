"""The oracle is pinned here: every formulation under oracle/ is checked against the golden vectors
that the unmodified reference produced (tests/golden/make_golden.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import afa_oracle as O
from oracle import torch_path as TP

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case_args(c):
    B, C, T, is_beta, logscale = [int(v) for v in c["meta"]]
    beta = c["beta"] if is_beta else None
    return B, C, T, beta, bool(logscale)


def test_taps_match_reference(golden):
    taps64 = O.default_taps()
    assert taps64.shape == (12,)
    np.testing.assert_allclose(taps64.astype(np.float32), golden["taps_f32"], rtol=0, atol=6e-8)
    np.testing.assert_array_equal(TP.make_taps().reshape(-1).numpy(), golden["taps_f32"])  # same torch ops => same bits
    # SURVEY.md section 4 decimals
    assert abs(golden["taps_f32"][5] - 0.44320979714393616) < 1e-7
    assert abs(golden["taps_f32"][0] - 0.0020289646927267313) < 1e-9
    np.testing.assert_allclose(golden["taps_f32"], golden["taps_f32"][::-1], atol=0)


def test_numpy_oracle_forward_vs_reference_f64(golden, golden_cases):
    taps = golden["taps_f32"].astype(np.float64)
    for name, c in golden_cases.items():
        B, C, T, beta, logscale = _case_args(c)
        y = O.activation1d_forward(c["x"], c["alpha"], beta, logscale, taps, taps)
        assert y.shape == (B, C, T)
        assert O.max_normalised_error(y, c["y_f64"]) < 1e-13, name
        assert O.max_normalised_error(y, c["y_f32"]) < 1e-5, name


def test_numpy_oracle_backward_vs_reference_autograd(golden, golden_cases):
    taps = golden["taps_f32"].astype(np.float64)
    for name, c in golden_cases.items():
        B, C, T, beta, logscale = _case_args(c)
        gx, ga, gb = O.activation1d_backward(c["x"], c["gy"], c["alpha"], beta, logscale, taps, taps)
        assert O.max_normalised_error(gx, c["gx_f64"]) < 1e-12, name
        assert O.max_normalised_error(ga, c["galpha_f64"]) < 1e-11, name
        if beta is not None:
            assert O.max_normalised_error(gb, c["gbeta_f64"]) < 1e-11, name
        else:
            assert gb is None


def test_torch_port_bitwise_vs_reference_f32(golden, golden_cases):
    """Same ATen op sequence, same machine => identical bits to the reference's fp32 output."""
    taps = torch.from_numpy(golden["taps_f32"]).view(1, 1, 12)
    for name, c in golden_cases.items():
        B, C, T, beta, logscale = _case_args(c)
        x = torch.from_numpy(c["x"])
        a = torch.from_numpy(c["alpha"])
        b = None if beta is None else torch.from_numpy(beta)
        y = TP.activation1d_torch(x, a, b, logscale, taps, taps)
        assert O.max_normalised_error(y.numpy(), c["y_f32"]) < 1e-6, name
        gx, ga, gb = TP.activation1d_torch_grads(x, torch.from_numpy(c["gy"]), a, b, logscale, taps, taps)
        assert O.max_normalised_error(gx.numpy(), c["gx_f32"]) < 1e-6, name
        assert O.max_normalised_error(ga.numpy(), c["galpha_f32"]) < 1e-5, name


@pytest.fixture(scope="module")
def c_oracle():
    subprocess.check_call(["make", "-s", "-C", os.path.join(REPO, "oracle")])
    lib = ctypes.CDLL(os.path.join(REPO, "oracle", "_build", "libafa_oracle.so"))
    dp = ctypes.POINTER(ctypes.c_double)
    lib.afa_oracle_fwd_f64.argtypes = [dp, dp, dp, dp, ctypes.c_int, dp, dp, ctypes.c_long, ctypes.c_long, ctypes.c_long]
    lib.afa_oracle_bwd_f64.argtypes = [dp, dp, dp, dp, dp, dp, dp, ctypes.c_int, dp, dp, ctypes.c_long, ctypes.c_long, ctypes.c_long]
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def test_c_oracle_vs_reference(golden, golden_cases, c_oracle):
    taps = np.ascontiguousarray(golden["taps_f32"].astype(np.float64))
    for name, c in golden_cases.items():
        B, C, T, beta, logscale = _case_args(c)
        x = np.ascontiguousarray(c["x"].astype(np.float64))
        gy = np.ascontiguousarray(c["gy"].astype(np.float64))
        a = np.ascontiguousarray(c["alpha"].astype(np.float64))
        b = None if beta is None else np.ascontiguousarray(beta.astype(np.float64))
        y = np.empty_like(x)
        assert c_oracle.afa_oracle_fwd_f64(_dp(x), _dp(y), _dp(a), _dp(b), int(logscale), _dp(taps), _dp(taps), B, C, T) == 0
        assert O.max_normalised_error(y, c["y_f64"]) < 1e-13, name
        gx = np.empty_like(x)
        ga = np.empty(C)
        gb = None if b is None else np.empty(C)
        assert c_oracle.afa_oracle_bwd_f64(_dp(x), _dp(gy), _dp(gx), _dp(ga), _dp(gb), _dp(a), _dp(b), int(logscale),
                                           _dp(taps), _dp(taps), B, C, T) == 0
        assert O.max_normalised_error(gx, c["gx_f64"]) < 1e-12, name
        assert O.max_normalised_error(ga, c["galpha_f64"]) < 1e-11, name
        if gb is not None:
            assert O.max_normalised_error(gb, c["gbeta_f64"]) < 1e-11, name


def test_impulse_response_and_dc_identity(golden):
    """composite response is symmetric, sums to ~1; DC passes through."""
    taps = golden["taps_f32"].astype(np.float64)
    imp = np.zeros((1, 1, 41))
    imp[0, 0, 20] = 1.0
    h = O.activation1d_forward(imp, np.zeros(1), np.ones(1), False, taps, taps).reshape(-1)
    np.testing.assert_allclose(h, golden["impulse_response_f64"], atol=1e-15)
    assert abs(h.sum() - 1.0) < 1e-6 and abs(h[20] - 0.86814) < 1e-4
    cst = np.full((1, 2, 50), 0.7)
    a, b = np.array([0.3, -0.2]), np.array([0.1, 0.4])
    y = O.activation1d_forward(cst, a, b, True, taps, taps)
    expect = 0.7 + np.sin(np.exp(a) * 0.7) ** 2 / (np.exp(b) + 1e-9)
    np.testing.assert_allclose(y, np.broadcast_to(expect[None, :, None], y.shape), rtol=0, atol=5e-7)
