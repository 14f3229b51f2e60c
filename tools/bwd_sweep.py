"""GPU: backward kernel timing per segment length (tuning aid). usage: python tools/bwd_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: F401  (sys.path)
import torch
from afa_b200 import Activation1d, _lib, functional as Fn
from afa_b200.activations import SnakeBeta

dev = torch.device("cuda:0")
shapes = [(16, 768, 3444), (16, 384, 13776), (16, 24, 220416), (2, 96, 55104), (2, 24, 220416), (2, 384, 13776), (2, 768, 3444), (32, 96, 2048), (32, 24, 8192)]
for dtype in (torch.float32, torch.bfloat16):
    for ch in (5, 9, 13, 17):
        _lib.set_tuning(1, ch, 0)
        for (b, c, t) in shapes:
            m = Activation1d(activation=SnakeBeta(c, alpha_logscale=True)).to(dev)
            n = b * c * t
            es = 4 if dtype == torch.float32 else 2
            nbuf = max(2, min(8, int(1.2e9 // (n * es * 3))))
            xs = [torch.randn(b, c, t, device=dev).to(dtype) for _ in range(nbuf)]
            gs = [torch.randn(b, c, t, device=dev).to(dtype) for _ in range(nbuf)]
            tu, td = m._host_taps()
            a_, b_ = m.act.alpha.detach(), m.act.beta.detach()
            Fn.activation1d_backward_raw(xs[0], gs[0], a_, b_, tu, td, True)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(2 * nbuf):
                    Fn.activation1d_backward_raw(xs[i % nbuf], gs[i % nbuf], a_, b_, tu, td, True)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (6 * nbuf)
            print(f"{str(dtype):15s} ch={ch:2d} B={b:3d} C={c:4d} T={t:7d} {us:8.1f} us {n * 3 * es / us / 1e3:8.1f} GB/s {n / us / 1e3:7.1f} Gelem/s", flush=True)
            del xs, gs, g
            torch.cuda.empty_cache()
_lib.set_tuning(1, 0, 0)
print(_lib.kernel_info(1, 0, 8192), _lib.kernel_info(1, 1, 8192))
