"""Functional form of the fused Activation1d + its autograd.Function (torch is plumbing only).

Mirrors what autograd does for the reference's op sequence
(reference: BigVGAN/alias_free_activation/act.py:25-30; backward reached from
BigVGAN/train_binaural_mel.py:787), but in two launches: one fused forward, one fused backward
(+ a tiny deterministic finalize for the alpha/beta gradients).
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

_F12 = ctypes.c_float * 12


def _dtype_code(x: torch.Tensor) -> int:
    if x.dtype == torch.float32:
        return _lib.AFA_DTYPE_F32
    if x.dtype == torch.bfloat16:
        return _lib.AFA_DTYPE_BF16
    raise TypeError(f"fused Activation1d supports float32 and bfloat16 activations, got {x.dtype}")


def host_taps(t: torch.Tensor):
    """12 filter taps of a module buffer as a host ctypes array (one D2H copy; cache the result)."""
    v = t.detach().reshape(-1).to(dtype=torch.float32, device="cpu")
    if v.numel() != 12:
        raise ValueError(f"fused Activation1d needs 12-tap filters, got {v.numel()} taps")
    return _F12(*v.tolist())


def _check_inputs(x: torch.Tensor, alpha: torch.Tensor, beta):
    if x.dim() != 3:
        # same failure mode as `_, C, _ = x.shape` in the reference (resample.py:30)
        raise ValueError(f"not enough values to unpack: expected a [B, C, T] tensor, got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("fused Activation1d runs on CUDA tensors only (there is no CPU fallback)")
    C = x.shape[1]
    for name, p in (("alpha", alpha), ("beta", beta)):
        if p is None:
            continue
        if p.numel() != C:
            raise RuntimeError(f"The size of {name} ({p.numel()}) must match the channel dimension of x ({C})")
        if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
            raise RuntimeError(f"{name} must be a contiguous float32 CUDA tensor")
        if p.device != x.device:
            raise RuntimeError(f"{name} is on {p.device}, x is on {x.device}")


def _row_pitch(t) -> int:
    """Elements between consecutive (batch, channel) rows of a [B, C, T] tensor whose rows are dense and equally spaced
    (e.g. a time slice x[:, :, :T] of a longer buffer), or 0 when the layout is anything else."""
    if t.dim() != 3 or t.stride(2) != 1:
        return 0
    B, C, T = t.shape
    pitch = t.stride(1)
    if pitch < T or (B > 1 and t.stride(0) != C * pitch):
        return 0
    return pitch


def activation1d_forward_raw(x, alpha, beta, taps_up, taps_down, logscale: bool, out=None):
    """One call of afa_activation1d_fwd on the current stream.  x: [B,C,T] fp32/bf16; a bf16 tensor whose rows are equally
    spaced views (x[:, :, :T] of a longer buffer) goes through afa_activation1d_fwd_pitched without a copy when the kernel
    can take it (include/afa_b200.h), everything else non-contiguous is copied first."""
    _check_inputs(x, alpha, beta)
    if not x.is_contiguous():
        pitch = _row_pitch(x) if (out is None and x.dtype == torch.bfloat16) else 0
        if pitch:
            B, C, T = x.shape
            y = torch.empty((B, C, T), dtype=x.dtype, device=x.device)
            flags = (_lib.AFA_FLAG_LOGSCALE if logscale else 0) | (_lib.AFA_FLAG_SNAKE if beta is None else 0)
            lib = _lib.load_library()
            with torch.cuda.device_of(x):
                rc = lib.afa_activation1d_fwd_pitched(
                    x.data_ptr(), pitch, y.data_ptr(), T, alpha.data_ptr(), None if beta is None else beta.data_ptr(),
                    taps_up, taps_down, B, C, T, _dtype_code(x), flags, torch.cuda.current_stream(x.device).cuda_stream)
            if rc == 0:
                return y
            if rc != _lib.AFA_ERR_ALIGNMENT:
                _lib.check(rc, "afa_activation1d_fwd_pitched")
        x = x.contiguous()
    y = torch.empty_like(x) if out is None else out
    B, C, T = x.shape
    flags = (_lib.AFA_FLAG_LOGSCALE if logscale else 0) | (_lib.AFA_FLAG_SNAKE if beta is None else 0)
    lib = _lib.load_library()
    with torch.cuda.device_of(x):
        rc = lib.afa_activation1d_fwd(
            x.data_ptr(), y.data_ptr(), alpha.data_ptr(), None if beta is None else beta.data_ptr(),
            taps_up, taps_down, B, C, T, _dtype_code(x), flags,
            torch.cuda.current_stream(x.device).cuda_stream,
        )
    _lib.check(rc, "afa_activation1d_fwd")
    return y


def activation1d_backward_raw(x, gy, alpha, beta, taps_up, taps_down, logscale: bool):
    """One call of afa_activation1d_bwd: returns (gx, galpha, gbeta|None), parameter grads in fp32."""
    _check_inputs(x, alpha, beta)
    if gy.shape != x.shape or gy.dtype != x.dtype:
        raise RuntimeError(f"grad_output {tuple(gy.shape)}/{gy.dtype} does not match x {tuple(x.shape)}/{x.dtype}")
    x = x.contiguous()
    gy = gy.contiguous()
    B, C, T = x.shape
    gx = torch.empty_like(x)
    galpha = torch.empty(C, dtype=torch.float32, device=x.device)
    gbeta = None if beta is None else torch.empty(C, dtype=torch.float32, device=x.device)
    flags = (_lib.AFA_FLAG_LOGSCALE if logscale else 0) | (_lib.AFA_FLAG_SNAKE if beta is None else 0)
    lib = _lib.load_library()
    code = _dtype_code(x)
    ws_bytes = int(lib.afa_bwd_workspace_bytes(B, C, T, code))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=x.device)
    with torch.cuda.device_of(x):
        rc = lib.afa_activation1d_bwd(
            x.data_ptr(), gy.data_ptr(), gx.data_ptr(), galpha.data_ptr(),
            None if gbeta is None else gbeta.data_ptr(),
            alpha.data_ptr(), None if beta is None else beta.data_ptr(),
            taps_up, taps_down, B, C, T, code, flags, ws.data_ptr(), ws_bytes,
            torch.cuda.current_stream(x.device).cuda_stream,
        )
    _lib.check(rc, "afa_activation1d_bwd")
    return gx, galpha, gbeta


class _Activation1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, alpha, beta, taps_up, taps_down, logscale):
        a32 = alpha.detach().float().contiguous()
        b32 = None if beta is None else beta.detach().float().contiguous()
        if not any(ctx.needs_input_grad[:3]):
            # inference: nothing is saved for a backward pass, so a strided view (x[:, :, :T] of a longer buffer) can go to the
            # kernel as it is (afa_activation1d_fwd_pitched) instead of through a dense copy
            return activation1d_forward_raw(x.detach(), a32, b32, taps_up, taps_down, logscale)
        xc = x.detach().contiguous()
        y = activation1d_forward_raw(xc, a32, b32, taps_up, taps_down, logscale)
        ctx.save_for_backward(xc, a32, b32 if b32 is not None else a32)
        ctx.has_beta = beta is not None
        ctx.taps = (taps_up, taps_down)
        ctx.logscale = logscale
        ctx.param_dtypes = (alpha.dtype, None if beta is None else beta.dtype)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gy):
        x, a32, b32 = ctx.saved_tensors
        beta = b32 if ctx.has_beta else None
        gx, ga, gb = activation1d_backward_raw(x, gy.to(x.dtype), a32, beta, ctx.taps[0], ctx.taps[1], ctx.logscale)
        ga = ga.to(ctx.param_dtypes[0])
        if gb is not None:
            gb = gb.to(ctx.param_dtypes[1])
        return gx, ga, gb, None, None, None


def activation1d(x, alpha, beta, taps_up, taps_down, logscale: bool):
    """Differentiable fused down2x(snake(up2x(x))). taps_* are host arrays from `host_taps`."""
    return _Activation1dFn.apply(x, alpha, beta, taps_up, taps_down, logscale)
