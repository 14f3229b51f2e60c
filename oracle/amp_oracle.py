"""CPU oracle for the AMP-block rows -- TEST INFRASTRUCTURE ONLY.

Plain numpy (float64 unless handed another dtype) restatement of the reference generator around the
anti-aliased activation, op by op:

    Conv1d / ConvTranspose1d          torch.nn.functional.conv1d / conv_transpose1d (third-party PyTorch,
                                      reference pin torch>=1.13.0, requirements.txt:2); call sites
                                      BigVGAN/bigvgan.py:56-88 (AMPBlock convs), :290-310 (ups), :284, :349
    AMPBlock1.forward                 BigVGAN/bigvgan.py:132-141
    AMPBlock2.forward                 BigVGAN/bigvgan.py:233-236
    BigVGAN.forward                   BigVGAN/bigvgan.py:361-387
    int16 stereo PCM                  BigVGAN/inference_e2e.py:174-201 (MAX_WAV_VALUE = 32767, meldataset.py:20)

and of the DECOMPOSITION the channels-last kernels implement (bias / residual / mean folded into the
activation, tail fused), so tests can check both "oracle == reference" (tests/golden/amp_golden.npz,
produced by tests/golden/make_golden_amp.py from the unmodified reference) and "CUDA == oracle".

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

from . import afa_oracle as O

MAX_WAV_VALUE = 32767.0  # meldataset.py:20


# --------------------------------------------------------------------------------------
# dense convolutions (torch semantics: cross-correlation, zero padding)
# --------------------------------------------------------------------------------------
def conv1d(x: np.ndarray, w: np.ndarray, b=None, padding: int = 0, dilation: int = 1) -> np.ndarray:
    """F.conv1d(x [B,Ci,T], w [Co,Ci,k], b, stride=1, padding, dilation)."""
    B, Ci, T = x.shape
    Co, Ci2, k = w.shape
    assert Ci == Ci2
    xp = np.zeros((B, Ci, T + 2 * padding), dtype=x.dtype)
    xp[:, :, padding:padding + T] = x
    To = T + 2 * padding - dilation * (k - 1)
    y = np.zeros((B, Co, To), dtype=x.dtype)
    for j in range(k):
        y += np.einsum("oc,bct->bot", w[:, :, j], xp[:, :, j * dilation:j * dilation + To])
    if b is not None:
        y += b[None, :, None]
    return y


def conv_transpose1d(x: np.ndarray, w: np.ndarray, b=None, stride: int = 1, padding: int = 0) -> np.ndarray:
    """F.conv_transpose1d(x [B,Ci,T], w [Ci,Co,k], b, stride, padding)."""
    B, Ci, T = x.shape
    Ci2, Co, k = w.shape
    assert Ci == Ci2
    full = np.zeros((B, Co, (T - 1) * stride + k), dtype=x.dtype)
    for j in range(k):
        full[:, :, j:j + (T - 1) * stride + 1:stride] += np.einsum("co,bct->bot", w[:, :, j], x)
    y = full[:, :, padding:full.shape[2] - padding]
    if b is not None:
        y = y + b[None, :, None]
    return y


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    return (kernel_size * dilation - dilation) // 2  # utils.py:79-80


# --------------------------------------------------------------------------------------
# reference structure, op by op
# --------------------------------------------------------------------------------------
def _act(sd, prefix, x, logscale=True):
    alpha = sd[prefix + "act.alpha"]
    beta = sd.get(prefix + "act.beta")
    return O.activation1d_forward(x, alpha, beta, logscale,
                                  sd[prefix + "upsample.filter"].reshape(-1), sd[prefix + "downsample.lowpass.filter"].reshape(-1))


def ampblock1_forward(sd, prefix, x, kernel_size, dilations, logscale=True):
    """bigvgan.py:132-141."""
    for n, d in enumerate(dilations):
        xt = _act(sd, f"{prefix}activations.{2 * n}.", x, logscale)
        xt = conv1d(xt, sd[f"{prefix}convs1.{n}.weight"], sd[f"{prefix}convs1.{n}.bias"], get_padding(kernel_size, d), d)
        xt = _act(sd, f"{prefix}activations.{2 * n + 1}.", xt, logscale)
        xt = conv1d(xt, sd[f"{prefix}convs2.{n}.weight"], sd[f"{prefix}convs2.{n}.bias"], get_padding(kernel_size, 1), 1)
        x = xt + x
    return x


def ampblock2_forward(sd, prefix, x, kernel_size, dilations, logscale=True):
    """bigvgan.py:233-236."""
    for n, d in enumerate(dilations):
        xt = _act(sd, f"{prefix}activations.{n}.", x, logscale)
        xt = conv1d(xt, sd[f"{prefix}convs.{n}.weight"], sd[f"{prefix}convs.{n}.bias"], get_padding(kernel_size, d), d)
        x = xt + x
    return x


def tail_forward(sd, x, use_tanh: bool, logscale=True):
    """bigvgan.py:379-385: activation_post -> conv_post -> tanh | clamp."""
    y = _act(sd, "activation_post.", x, logscale)
    y = conv1d(y, sd["conv_post.weight"], sd.get("conv_post.bias"), 3, 1)
    return np.tanh(y) if use_tanh else np.clip(y, -1.0, 1.0)


def pcm_stereo(wave_lr: np.ndarray) -> np.ndarray:
    """inference_e2e.py:189-201: [2, T] float (left, right) -> int16 [T, 2]; astype truncates toward zero."""
    return (wave_lr * MAX_WAV_VALUE).astype("int16").T


def generator_forward(sd, mel, h, logscale=True):
    """bigvgan.py:361-387 for a state dict without weight norm; h: dict with the generator hyper-parameters."""
    x = conv1d(mel, sd["conv_pre.weight"], sd["conv_pre.bias"], 3)
    nk = len(h["resblock_kernel_sizes"])
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"], u, (k - u) // 2)
        xs = None
        for j in range(nk):
            fn = ampblock1_forward if h["resblock"] == "1" else ampblock2_forward
            y = fn(sd, f"resblocks.{i * nk + j}.", x, h["resblock_kernel_sizes"][j], h["resblock_dilation_sizes"][j], logscale)
            xs = y if xs is None else xs + y
        x = xs / nk
    return tail_forward(sd, x, h.get("use_tanh_at_final", True), logscale)


# --------------------------------------------------------------------------------------
# the decomposition the channels-last kernels implement (layout [B, T, C])
# --------------------------------------------------------------------------------------
def amp_activation1d_cl(x_btc, alpha, beta=None, logscale=True, bias=None, res=None, taps_up=None, taps_down=None):
    """afa_amp_activation1d_fwd_cl: returns (xsum = x + res, y = Activation1d(x + res + bias[c])) in [B, T, C];
    the bias stays pending on the residual stream (the caller carries it to the next call / the mean)."""
    xs = np.asarray(x_btc, dtype=np.float64)
    if res is not None:
        xs = xs + np.asarray(res, dtype=np.float64)
    xb = xs if bias is None else xs + np.asarray(bias, dtype=np.float64)[None, None, :]
    y = O.activation1d_forward(np.ascontiguousarray(xb.transpose(0, 2, 1)), alpha, beta, logscale, taps_up, taps_down)
    return xs, np.ascontiguousarray(y.transpose(0, 2, 1))


def amp_act_conv_cl(x_btc, alpha, beta, logscale, w_kcc, dilation, bias=None, res=None, taps_up=None, taps_down=None,
                    round_act=None, addend=None):
    """afa_amp_act_conv_fwd_cl: (xsum = x + res, y = conv1d(Activation1d(x + res + bias), w, 'same', dilation) + addend)
    in [B, T, C]; w_kcc: [k, C_out, C_in].  `round_act` (a function) models the bf16 rounding of the activated tile."""
    xs, a = amp_activation1d_cl(x_btc, alpha, beta, logscale, bias, res, taps_up, taps_down)
    if round_act is not None:
        a = round_act(a)
    w = np.ascontiguousarray(np.asarray(w_kcc, dtype=np.float64).transpose(1, 2, 0))      # [C_out, C_in, k]
    k = w.shape[2]
    y = conv1d(np.ascontiguousarray(a.transpose(0, 2, 1)), w, None, get_padding(k, dilation), dilation)
    y = np.ascontiguousarray(y.transpose(0, 2, 1))
    if addend is not None:
        y = y + np.asarray(addend, dtype=np.float64)
    return xs, y


def resblock_mean(xts, xress, bias_sum=None, scale=None):
    """afa_resblock_mean: scale * (sum_j (xt_j + xres_j) + bias_sum[c]); scale defaults to 1/len."""
    acc = np.zeros_like(np.asarray(xts[0], dtype=np.float64))
    for a, r in zip(xts, xress):
        acc = acc + np.asarray(a, dtype=np.float64)
        if r is not None:                                   # None: xt_j already contains its residual stream
            acc = acc + np.asarray(r, dtype=np.float64)
    if bias_sum is not None:
        acc = acc + np.asarray(bias_sum, dtype=np.float64)
    return acc * (1.0 / len(xts) if scale is None else scale)


def tail_cl(x_btc, alpha, beta, logscale, w_post, bias_post=None, use_tanh=False, taps_up=None, taps_down=None):
    """afa_tail_fwd_cl: wave float64 [B, T] from channels-last x; w_post [C, 7]."""
    x = np.ascontiguousarray(np.asarray(x_btc, dtype=np.float64).transpose(0, 2, 1))
    y = O.activation1d_forward(x, alpha, beta, logscale, taps_up, taps_down)
    w = np.asarray(w_post, dtype=np.float64)[None, :, :]
    out = conv1d(y, w, None if bias_post is None else np.asarray(bias_post, dtype=np.float64).reshape(1), 3, 1)[:, 0, :]
    return np.tanh(out) if use_tanh else np.clip(out, -1.0, 1.0)


def pcm_interleave(wave_bt: np.ndarray, il: int = 2, scale: float = MAX_WAV_VALUE) -> np.ndarray:
    """[B, T] float -> int16 [B // il, T, il] (truncation toward zero, like astype('int16'))."""
    B, T = wave_bt.shape
    v = (np.asarray(wave_bt, dtype=np.float32) * np.float32(scale)).astype("int16")
    return np.ascontiguousarray(v.reshape(B // il, il, T).transpose(0, 2, 1))


def ampblock_decomposed(sd, prefix, x_bct, kernel_size, dilations, amp1: bool, logscale=True, up_bias=None):
    """The engine's schedule for one AMPBlock (afa_b200/engine.py::_resblock) with numpy convolutions WITHOUT bias:
    every bias and residual add goes through `amp_activation1d_cl`.  Returns (xt, pending bias, residual, pending bias)."""
    to_cl = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
    to_ncw = lambda a: np.ascontiguousarray(a.transpose(0, 2, 1))
    r, r_pend = to_cl(np.asarray(x_bct, dtype=np.float64)), up_bias
    t, t_bias = None, None

    def act(idx, x_cl, bias=None, res=None):
        p = f"{prefix}activations.{idx}."
        return amp_activation1d_cl(x_cl, sd[p + "act.alpha"], sd.get(p + "act.beta"), logscale, bias, res,
                                   sd[p + "upsample.filter"].reshape(-1), sd[p + "downsample.lowpass.filter"].reshape(-1))

    def add(*bs):
        bs = [np.asarray(b, dtype=np.float64) for b in bs if b is not None]
        return None if not bs else sum(bs)

    for n, d in enumerate(dilations):
        a_idx = 2 * n if amp1 else n
        c1 = f"{prefix}convs1.{n}." if amp1 else f"{prefix}convs.{n}."
        if n == 0:
            _, a = act(a_idx, r, bias=r_pend)
        else:
            r_pend = add(t_bias, r_pend)                     # biases the new residual stream still lacks
            r, a = act(a_idx, t, bias=r_pend, res=r)
        t = to_cl(conv1d(to_ncw(a), np.asarray(sd[c1 + "weight"], dtype=np.float64), None, get_padding(kernel_size, d), d))
        t_bias = sd[c1 + "bias"]
        if amp1:
            _, a = act(a_idx + 1, t, bias=t_bias)
            c2 = f"{prefix}convs2.{n}."
            t = to_cl(conv1d(to_ncw(a), np.asarray(sd[c2 + "weight"], dtype=np.float64), None, get_padding(kernel_size, 1), 1))
            t_bias = sd[c2 + "bias"]
    return t, t_bias, r, r_pend
