"""Channels-last inference engine for the BigVGAN generator.

Same arithmetic as `BigVGAN.forward` (reference: BigVGAN/bigvgan.py:361-387, AMPBlock1 :132-141,
AMPBlock2 :233-236), re-laid-out for B200:

* activations stay CHANNELS-LAST ([B, T, C]) from conv_pre to the waveform, so cuDNN's NHWC tensor-core
  kernels run without the NCHW<->NHWC conversion kernels that wrap every convolution of the
  [B, C, T] path, and without a separate bias kernel per convolution;
* every Conv1d runs as a (k x 1) 2-D convolution along H over the view [B, C, H = T/d, W = d] of the
  same memory (t = d*h + w): a DILATED convolution becomes a dense one (polyphase split, zero-copy;
  the activation kernel zero-fills the rows that round T up to a multiple of d);
* the bias of every convolution, the residual add `x = xt + x`, the resblock mean and the whole tail
  (activation_post -> conv_post -> clamp -> int16 PCM) are folded into the hand-written kernels
  (`afa_amp_activation1d_fwd_cl`, `afa_resblock_mean`, `afa_tail_fwd_cl`).

Inference only (no autograd).  The dense convolutions remain library code (cuDNN through torch).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import functional_cl as FC
from .modules import Activation1d
from .vocoder import AMPBlock1, AMPBlock2, BigVGANGenerator


def _cl_weight(w: torch.Tensor, dtype) -> torch.Tensor:
    """[Co, Ci, k] (Conv1d) or [Ci, Co, k] (ConvTranspose1d) -> [.., .., k, 1] in channels-last memory."""
    return w.detach().to(dtype).unsqueeze(3).contiguous(memory_format=torch.channels_last)


class _Act:
    """Parameters of one fused activation as the C ABI wants them (fp32 alpha/beta, host taps)."""

    def __init__(self, m: Activation1d):
        self.alpha = m.act.alpha.detach().float().contiguous()
        beta = getattr(m.act, "beta", None)
        self.beta = None if beta is None else beta.detach().float().contiguous()
        self.logscale = bool(m.act.alpha_logscale)
        self.taps_up, self.taps_down = m._host_taps()


class _Conv:
    def __init__(self, conv: torch.nn.Conv1d, dtype, fuse: bool = False):
        self.w = _cl_weight(conv.weight, dtype)
        self.bias = None if conv.bias is None else conv.bias.detach().float().contiguous()
        self.k = conv.kernel_size[0]
        self.d = conv.dilation[0]
        if conv.stride[0] != 1 or conv.padding[0] != (self.k * self.d - self.d) // 2 or conv.groups != 1:
            raise NotImplementedError("engine convolutions are stride-1, 'same'-padded, dense")
        # narrow stages: the activation in front of this convolution runs as its prologue (one fused kernel)
        self.fused = (fuse and conv.in_channels == conv.out_channels
                      and FC.act_conv_supported(conv.in_channels, self.k, self.d, dtype))
        self.w_kcc = conv.weight.detach().to(dtype).permute(2, 0, 1).contiguous() if self.fused else None

    def tpad(self, T: int) -> int:
        return (T + self.d - 1) // self.d * self.d

    def __call__(self, x_mem: torch.Tensor) -> torch.Tensor:
        """x_mem: dense [B, Tp, C] with Tp % d == 0 and zero rows beyond T.  Returns [B, Tp, Co] (rows >= T: garbage)."""
        B, Tp, C = x_mem.shape
        x4 = x_mem.view(B, Tp // self.d, self.d, C).permute(0, 3, 1, 2)           # logical [B, C, H, W]
        y4 = F.conv2d(x4, self.w, None, 1, (self.k // 2, 0), 1)
        return y4.permute(0, 2, 3, 1).reshape(B, Tp, self.w.shape[0])


class ChannelsLastVocoder:
    """mel [B, num_mels, T_mel] -> (wave float32 [B, 1, T], pcm int16 [B/2, T, 2] | None)."""

    def __init__(self, gen: BigVGANGenerator, dtype=torch.bfloat16, parallel_resblocks: bool = True,
                 fuse_narrow_convs: bool = True):
        self.dtype = dtype
        fz = bool(fuse_narrow_convs)
        self.h = gen.h
        # the num_kernels resblocks of a stage are independent until their mean: run them on side streams so that
        # one block's small launches fill the tails of another's (captured as parallel branches of the CUDA graph)
        self.parallel_resblocks = parallel_resblocks
        self._side_streams = None
        self.device = next(gen.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("ChannelsLastVocoder runs on CUDA only (there is no CPU fallback)")
        self.num_kernels = gen.num_kernels
        self.pre = _Conv(gen.conv_pre, dtype)
        self.pre_bias = None if gen.conv_pre.bias is None else gen.conv_pre.bias.detach().to(dtype)
        self.stages: List[Dict] = []
        for i in range(gen.num_upsamples):
            (up,) = list(gen.ups[i])
            blocks = []
            for j in range(gen.num_kernels):
                rb = gen.resblocks[i * gen.num_kernels + j]
                if isinstance(rb, AMPBlock1):
                    its = [dict(a1=_Act(a1), c1=_Conv(c1, dtype, fz), a2=_Act(a2), c2=_Conv(c2, dtype, fz))
                           for c1, c2, a1, a2 in zip(rb.convs1, rb.convs2, rb.activations[::2], rb.activations[1::2])]
                elif isinstance(rb, AMPBlock2):
                    its = [dict(a1=_Act(a), c1=_Conv(c, dtype, fz), a2=None, c2=None) for c, a in zip(rb.convs, rb.activations)]
                else:
                    raise NotImplementedError(type(rb).__name__)
                blocks.append(its)
            up_bias = None if up.bias is None else up.bias.detach().float().contiguous()
            self.stages.append(dict(
                # upsamplers run along W ([B, C, 1, T]): cuDNN's strided-dgrad kernels are 2-5x faster in this orientation
                # than along H (measured: 3.0 -> 0.5 ms per 4-clip pass), the opposite of the stride-1 convolutions
                w=up.weight.detach().to(dtype).unsqueeze(2).contiguous(memory_format=torch.channels_last),
                bias=up_bias, k=up.kernel_size[0], u=up.stride[0], p=up.padding[0],
                cout=up.out_channels, blocks=blocks))
        self.post_act = _Act(gen.activation_post)
        self.w_post = gen.conv_post.weight.detach().float().reshape(-1, 7).contiguous()       # [1, C, 7] -> [C, 7]
        if gen.conv_post.kernel_size[0] != 7 or gen.conv_post.out_channels != 1:
            raise NotImplementedError("the tail kernel implements conv_post = Conv1d(C, 1, 7, padding=3)")
        self.b_post = None if gen.conv_post.bias is None else gen.conv_post.bias.detach().float().contiguous()
        self.use_tanh = bool(gen.use_tanh_at_final)
        self._bias_cache: Dict = {}

    # -- small fp32 bias sums, made once
    def _sum(self, *bs: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        bs = [b for b in bs if b is not None]
        if not bs:
            return None
        if len(bs) == 1:
            return bs[0]
        key = tuple(b.data_ptr() for b in bs)
        if key not in self._bias_cache:
            self._bias_cache[key] = torch.stack(bs).sum(0).contiguous()
        return self._bias_cache[key]

    def _act(self, a: _Act, x, T, **kw):
        return FC.amp_activation1d_cl(x, T, a.alpha, a.beta, a.taps_up, a.taps_down, a.logscale, **kw)

    def _act_conv(self, a: _Act, conv: _Conv, x, T, addend=None, **kw):
        """conv(Activation1d(x + res + bias)) without the convolution's bias: one fused kernel on the narrow stages,
        else the activation kernel (zero-filling the polyphase padding rows) followed by cuDNN."""
        # measured on the model's shapes (tools/actconv_sweep.py): at 24 channels the fused kernel is 1.7-2.3x faster than
        # activation + cuDNN with or without the residual prologue; at 48 channels only the plain variant wins
        if conv.fused and (addend is not None or kw.get("res") is None or x.shape[2] <= 32):
            return FC.amp_act_conv_cl(x, T, a.alpha, a.beta, a.taps_up, a.taps_down, a.logscale, conv.w_kcc, conv.k,
                                      conv.d, addend=addend, **kw)
        assert addend is None
        return conv(self._act(a, x, T, out_tpad=conv.tpad(T), **kw))

    def _resblock(self, its, x, T, up_bias):
        """One AMPBlock on the raw upsampler output x (its bias `up_bias` still pending).
        Returns (xt, pending bias of xt, residual stream or None, bias still missing from the residual stream):
        the block's output is xt + residual + both pending biases (residual None: xt already contains it)."""
        r, r_pending = x, up_bias
        t, t_bias = None, None
        last = len(its) - 1
        for n, it in enumerate(its):
            # narrow stages: the LAST convolution of an iteration adds the residual stream in its epilogue (fp32, one
            # rounding), so `x = xt + x` costs nothing and the next activation has no residual prologue
            final_conv = it["c2"] if it["c2"] is not None else it["c1"]
            fold_add = final_conv.fused and t is None
            if t is None:                                                    # the block's input: r (+ pending bias)
                kw = dict(bias=r_pending)
                src = r
            else:                                                            # x = xt + x, then a1(x)
                r_new = torch.empty(x.shape[0], T, x.shape[2], dtype=x.dtype, device=x.device)
                r_pending = self._sum(t_bias, r_pending)                     # biases the new residual stream still lacks
                kw = dict(bias=r_pending, res=r, xsum=r_new)
                src = t
                r = r_new
            if it["c2"] is None:
                t = self._act_conv(it["a1"], it["c1"], src, T, addend=r if fold_add else None, **kw)
                t_bias = it["c1"].bias
            else:
                t = self._act_conv(it["a1"], it["c1"], src, T, **kw)
                t = self._act_conv(it["a2"], it["c2"], t, T, bias=it["c1"].bias, addend=r if fold_add else None)
                t_bias = it["c2"].bias
            if fold_add:
                # t = conv + r is the new residual stream (lacking t_bias and whatever r lacked); restart the chain
                r_pending = self._sum(t_bias, r_pending)
                if n == last:
                    if t.shape[1] != T:
                        t = t[:, :T].contiguous()
                    return t, None, None, r_pending
                r, t, t_bias = t, None, None
        if t.shape[1] != T:
            t = t[:, :T].contiguous()                                        # only if the LAST convolution is dilated (AMPBlock2)
        return t, t_bias, r, r_pending

    def _resblocks(self, st, x, T):
        if not self.parallel_resblocks or len(st["blocks"]) < 2:
            return [self._resblock(its, x, T, st["bias"]) for its in st["blocks"]]
        dev = x.device
        cur = torch.cuda.current_stream(dev)
        if self._side_streams is None:
            self._side_streams = [torch.cuda.Stream(dev) for _ in range(len(st["blocks"]) - 1)]
        fork = torch.cuda.Event()
        fork.record(cur)
        outs, joins = [None] * len(st["blocks"]), []
        for j, its in enumerate(st["blocks"][1:], start=1):
            side = self._side_streams[j - 1]
            side.wait_event(fork)
            with torch.cuda.stream(side):
                outs[j] = self._resblock(its, x, T, st["bias"])
                for t in (outs[j][0], outs[j][2]):
                    if t is not None:
                        t.record_stream(cur)                # consumed by the mean kernel on the main stream
                ev = torch.cuda.Event()
                ev.record(side)
                joins.append(ev)
        outs[0] = self._resblock(st["blocks"][0], x, T, st["bias"])
        for ev in joins:
            cur.wait_event(ev)
        return outs

    @torch.no_grad()
    def __call__(self, mel: torch.Tensor, want_pcm: bool = False, pcm_interleave: int = 2, frame_map=None, t_out: int = 0):
        """`frame_map` (int32 [B, T_mel], device) + `t_out`: scatter every mel frame's hop of samples to frame
        frame_map[b, i] of an output of t_out samples that starts as silence (zero-frame restoration)."""
        B = mel.shape[0]
        x = mel.to(self.dtype).transpose(1, 2).contiguous()                  # [B, T_mel, num_mels]
        x = self.pre(x)
        if self.pre_bias is not None:
            x = x + self.pre_bias
        T = x.shape[1]
        for st in self.stages:
            x4 = x.view(B, 1, T, x.shape[2]).permute(0, 3, 1, 2)             # logical [B, C, 1, T]
            y4 = F.conv_transpose2d(x4, st["w"], None, (1, st["u"]), (0, st["p"]))
            T = y4.shape[3]
            x = y4.permute(0, 2, 3, 1).reshape(B, T, st["cout"])
            outs = self._resblocks(st, x, T)
            bias_sum = self._sum(*[b for o in outs for b in (o[1], o[3])])
            x = FC.resblock_mean([o[0] for o in outs], [o[2] for o in outs], bias_sum, 1.0 / self.num_kernels)
        a = self.post_act
        wave, pcm = FC.tail_cl(x, T, a.alpha, a.beta, a.taps_up, a.taps_down, a.logscale, self.w_post, self.b_post,
                               use_tanh=self.use_tanh, want_wave=True, want_pcm=want_pcm, pcm_interleave=pcm_interleave,
                               frame_map=frame_map, hop=T // mel.shape[2] if frame_map is not None else 0, t_out=t_out)
        return wave.view(B, 1, -1), pcm


class GraphedEngine:
    """CUDA-graph replay of the channels-last engine for a fixed mel shape: the ~330 launches of a pass
    (cuDNN convolutions + the fused AMP kernels) are captured once and replayed without host work."""

    def __init__(self, engine: ChannelsLastVocoder, batch: int, t_mel: int, want_pcm: bool = True, pcm_interleave: int = 2):
        self.engine = engine
        dev = engine.device
        self.static_in = torch.zeros(batch, engine.h["num_mels"], t_mel, device=dev, dtype=torch.float32)
        # cuDNN picks its convolution algorithms by timing them (benchmark mode) during the warm-up passes; the capture
        # then replays those choices.  The global switch is restored afterwards.
        old_benchmark = torch.backends.cudnn.benchmark
        torch.backends.cudnn.benchmark = True
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):                                   # warm-up: cuDNN algorithm picks, bias sums
                    engine(self.static_in, want_pcm=want_pcm, pcm_interleave=pcm_interleave)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_wave, self.static_pcm = engine(self.static_in, want_pcm=want_pcm, pcm_interleave=pcm_interleave)
        finally:
            torch.backends.cudnn.benchmark = old_benchmark

    def __call__(self, mel: torch.Tensor):
        self.static_in.copy_(mel, non_blocking=True)
        self.graph.replay()
        return self.static_wave, self.static_pcm
