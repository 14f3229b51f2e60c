"""Channels-last ([B, T, C]) AMP-block ops: thin ctypes calls into libafa_sm100.so (forward only).

What each call stands for in the reference (BigVGAN/bigvgan.py):
    amp_activation1d_cl   the Conv1d bias add + `x = xt + x` (:141) + the next Activation1d (:135,137)
    resblock_mean         `xs += resblocks[...](x)`, `x = xs / self.num_kernels` (:368-376)
    tail_cl               activation_post -> conv_post -> clamp | tanh (:379-385), optionally
                          `* MAX_WAV_VALUE`, astype("int16"), stereo interleave (inference_e2e.py:193-201)

Tensors are views over [B, T(+padding), C] memory: `x[b, t, c]` with stride(2) == 1 and
stride(1) == C; the batch stride is free (time padding).  torch is plumbing only.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .functional import _dtype_code


def _check_cl(name: str, t: torch.Tensor, B: int, T: int, C: int, dtype, device, rows=None):
    rows = T if rows is None else rows
    if t.dim() != 3 or t.shape[0] != B or t.shape[1] < rows or t.shape[2] != C:
        raise RuntimeError(f"{name}: expected a [B={B}, T>={rows}, C={C}] channels-last tensor, got {tuple(t.shape)}")
    if t.dtype != dtype or t.device != device:
        raise RuntimeError(f"{name}: dtype/device {t.dtype}/{t.device} does not match x {dtype}/{device}")
    if t.stride(2) != 1 or (t.shape[1] > 1 and t.stride(1) != C):
        raise RuntimeError(f"{name}: channels must be contiguous and rows dense (strides {t.stride()})")
    if B > 1 and t.stride(0) < rows * C:
        raise RuntimeError(f"{name}: batch stride {t.stride(0)} does not cover {rows} rows")


def _bstride(t: torch.Tensor) -> int:
    return int(t.stride(0)) if t.shape[0] > 1 else int(t.shape[1] * t.shape[2])


def _flags(logscale: bool, beta) -> int:
    return (_lib.AFA_FLAG_LOGSCALE if logscale else 0) | (_lib.AFA_FLAG_SNAKE if beta is None else 0)


def _f32(name, p, n, device):
    if p is None:
        return None
    if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or p.numel() != n or p.device != device:
        raise RuntimeError(f"{name} must be a contiguous float32 CUDA tensor of {n} elements on {device}")
    return p


def amp_activation1d_cl(x, T: int, alpha, beta, taps_up, taps_down, logscale: bool, *, bias=None, res=None,
                        xsum=None, out=None, out_tpad: int = 0):
    """y = Activation1d(x + res + bias) on channels-last data; returns y and fills `xsum` = x + res (the bias
    stays pending on the new residual stream: pass it on to the next call / to `resblock_mean`).

    x: [B, >=T, C] view (only rows < T are read).  out: [B, >=max(T, out_tpad), C] view or None (allocated
    dense, out_tpad rows).  Rows [T, out_tpad) of out are zero-filled."""
    if not x.is_cuda:
        raise RuntimeError("channels-last AMP ops run on CUDA tensors only (there is no CPU fallback)")
    B, _, C = x.shape
    dev, dt = x.device, x.dtype
    _check_cl("x", x, B, T, C, dt, dev)
    tp = max(T, out_tpad)
    if out is None:
        out = torch.empty(B, tp, C, dtype=dt, device=dev)
    _check_cl("out", out, B, T, C, dt, dev, rows=tp)
    if res is not None:
        _check_cl("res", res, B, T, C, dt, dev)
    if (xsum is None) != (res is None):
        raise RuntimeError("res and xsum come together: xsum = x + res is the new residual stream")
    if xsum is not None:
        _check_cl("xsum", xsum, B, T, C, dt, dev)
    alpha = _f32("alpha", alpha, C, dev)
    beta = _f32("beta", beta, C, dev)
    bias = _f32("bias", bias, C, dev)
    if B == 0 or T == 0:
        return out
    lib = _lib.load_library()
    with torch.cuda.device_of(x):
        rc = lib.afa_amp_activation1d_fwd_cl(
            x.data_ptr(), _bstride(x),
            None if res is None else res.data_ptr(), 0 if res is None else _bstride(res),
            None if bias is None else bias.data_ptr(),
            None if xsum is None else xsum.data_ptr(), 0 if xsum is None else _bstride(xsum),
            out.data_ptr(), _bstride(out), tp,
            alpha.data_ptr(), None if beta is None else beta.data_ptr(), taps_up, taps_down,
            B, C, T, _dtype_code(x), _flags(logscale, beta), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "afa_amp_activation1d_fwd_cl")
    return out


def act_conv_supported(channels: int, kernel_size: int, dilation: int, dtype) -> bool:
    """Is the fused activation -> convolution kernel compiled for this configuration?"""
    if dtype != torch.bfloat16:
        return False
    return bool(_lib.load_library().afa_amp_act_conv_supported(channels, kernel_size, dilation, _lib.AFA_DTYPE_BF16))


def amp_act_conv_cl(x, T: int, alpha, beta, taps_up, taps_down, logscale: bool, w_kcc, kernel_size: int, dilation: int,
                    *, bias=None, res=None, xsum=None, out=None, addend=None):
    """y = conv1d(Activation1d(x + res + bias), w, 'same' padding, dilation, NO bias) [+ addend] in one kernel;
    xsum = x + res.  `addend` ([B, T, C]) is the block's residual stream, added in fp32 before y is rounded: y is then
    the new residual stream (`x = xt + x`, bigvgan.py:141) and the next activation needs no residual prologue.

    bf16 channels-last tensors; w_kcc: bf16 [kernel_size, C, C] = conv.weight.permute(2, 0, 1).contiguous().
    Reference: `xt = a(x); xt = c(xt)` (bigvgan.py:134-138); the convolution's bias stays pending with the caller."""
    if not x.is_cuda:
        raise RuntimeError("channels-last AMP ops run on CUDA tensors only (there is no CPU fallback)")
    B, _, C = x.shape
    dev, dt = x.device, x.dtype
    if dt != torch.bfloat16:
        raise TypeError(f"the fused activation+convolution runs on bfloat16 activations, got {dt}")
    _check_cl("x", x, B, T, C, dt, dev)
    if out is None:
        out = torch.empty(B, T, C, dtype=dt, device=dev)
    _check_cl("out", out, B, T, C, dt, dev)
    if (xsum is None) != (res is None):
        raise RuntimeError("res and xsum come together: xsum = x + res is the new residual stream")
    if res is not None:
        _check_cl("res", res, B, T, C, dt, dev)
        _check_cl("xsum", xsum, B, T, C, dt, dev)
    if addend is not None:
        _check_cl("addend", addend, B, T, C, dt, dev)
    if w_kcc.dtype != torch.bfloat16 or tuple(w_kcc.shape) != (kernel_size, C, C) or not w_kcc.is_contiguous() or w_kcc.device != dev:
        raise RuntimeError(f"w_kcc must be a contiguous bfloat16 [{kernel_size}, {C}, {C}] tensor on {dev}")
    alpha = _f32("alpha", alpha, C, dev)
    beta = _f32("beta", beta, C, dev)
    bias = _f32("bias", bias, C, dev)
    if B == 0 or T == 0:
        return out
    lib = _lib.load_library()
    with torch.cuda.device_of(x):
        rc = lib.afa_amp_act_conv_fwd_cl(
            x.data_ptr(), _bstride(x),
            None if res is None else res.data_ptr(), 0 if res is None else _bstride(res),
            None if bias is None else bias.data_ptr(),
            None if xsum is None else xsum.data_ptr(), 0 if xsum is None else _bstride(xsum),
            None if addend is None else addend.data_ptr(), 0 if addend is None else _bstride(addend),
            out.data_ptr(), _bstride(out),
            alpha.data_ptr(), None if beta is None else beta.data_ptr(), taps_up, taps_down,
            w_kcc.data_ptr(), kernel_size, dilation, B, C, T, _dtype_code(x), _flags(logscale, beta),
            torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "afa_amp_act_conv_fwd_cl")
    return out


def resblock_mean(xts, xress, bias_sum=None, scale=None, out=None):
    """out = scale * (sum_j (xts[j] + xress[j]) + bias_sum[c]); all tensors dense [B, T, C] (or any dense [..., C])."""
    K = len(xts)
    if K != len(xress) or K < 1:
        raise RuntimeError("resblock_mean needs as many residual streams as convolution outputs")
    ref = xts[0]
    if not ref.is_cuda:
        raise RuntimeError("channels-last AMP ops run on CUDA tensors only (there is no CPU fallback)")
    for t in list(xts) + [r for r in xress if r is not None]:
        if t.shape != ref.shape or t.dtype != ref.dtype or t.device != ref.device or not t.is_contiguous():
            raise RuntimeError("resblock_mean: all tensors must be dense and agree in shape, dtype and device")
    C = ref.shape[-1]
    rows = ref.numel() // C
    bias_sum = _f32("bias_sum", bias_sum, C, ref.device)
    if out is None:
        out = torch.empty_like(ref)
    arr_t = (ctypes.c_void_p * K)(*[t.data_ptr() for t in xts])
    arr_r = (ctypes.c_void_p * K)(*[None if t is None else t.data_ptr() for t in xress])
    lib = _lib.load_library()
    with torch.cuda.device_of(ref):
        rc = lib.afa_resblock_mean(arr_t, arr_r, K, None if bias_sum is None else bias_sum.data_ptr(),
                                   float(1.0 / K if scale is None else scale), out.data_ptr(), rows, C,
                                   _dtype_code(ref), torch.cuda.current_stream(ref.device).cuda_stream)
    _lib.check(rc, "afa_resblock_mean")
    return out


def tail_cl(x, T: int, alpha, beta, taps_up, taps_down, logscale: bool, w_post, bias_post=None, use_tanh=False,
            want_wave=True, want_pcm=False, pcm_interleave: int = 2, pcm_scale: float = 32767.0, wave=None, pcm=None,
            frame_map=None, hop: int = 0, t_out: int = 0):
    """activation_post -> conv_post (C -> 1, k = 7) -> clamp | tanh on channels-last x [B, >=T, C].

    Returns (wave float32 [B, T_out] or None, pcm int16 [B // il, T_out, il] or None).  With `frame_map` (int32
    [B, T // hop] on the device) whole frames are scattered into rows of `t_out` samples that start as silence
    (zero-frame restoration, inference_e2e.py:80-111); otherwise T_out = T."""
    if not x.is_cuda:
        raise RuntimeError("channels-last AMP ops run on CUDA tensors only (there is no CPU fallback)")
    B, _, C = x.shape
    dev = x.device
    _check_cl("x", x, B, T, C, x.dtype, dev)
    alpha = _f32("alpha", alpha, C, dev)
    beta = _f32("beta", beta, C, dev)
    w_post = _f32("w_post", w_post, C * 7, dev)
    bias_post = _f32("bias_post", bias_post, 1, dev)
    if frame_map is not None:
        if frame_map.dtype != torch.int32 or not frame_map.is_contiguous() or frame_map.device != dev or hop < 1 \
                or T % hop or tuple(frame_map.shape) != (B, T // hop):
            raise RuntimeError(f"frame_map must be a contiguous int32 [{B}, T // hop] tensor on {dev}")
        t_out = int(t_out)
        make = torch.zeros                                   # frames nobody writes are the silence
    else:
        t_out, hop, make = T, 0, torch.empty
    if want_wave and wave is None:
        wave = make(B, t_out, dtype=torch.float32, device=dev)
    if want_pcm and pcm is None:
        if B % pcm_interleave:
            raise RuntimeError(f"batch {B} is not a multiple of pcm_interleave {pcm_interleave}")
        pcm = make(B // pcm_interleave, t_out, pcm_interleave, dtype=torch.int16, device=dev)
    lib = _lib.load_library()
    with torch.cuda.device_of(x):
        rc = lib.afa_tail_fwd_cl(
            x.data_ptr(), _bstride(x), alpha.data_ptr(), None if beta is None else beta.data_ptr(), taps_up, taps_down,
            w_post.data_ptr(), None if bias_post is None else bias_post.data_ptr(), 1 if use_tanh else 0,
            None if wave is None else wave.data_ptr(), None if pcm is None else pcm.data_ptr(), pcm_interleave,
            float(pcm_scale), None if frame_map is None else frame_map.data_ptr(), hop, t_out,
            B, C, T, _dtype_code(x), _flags(logscale, beta), torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "afa_tail_fwd_cl")
    return wave, pcm
