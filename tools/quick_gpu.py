"""Quick manual GPU probe (not a pytest file): parity + first timings. Usage: python tools/quick_gpu.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402  (sets sys.path)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from afa_b200 import Activation1d, _lib  # noqa: E402
from afa_b200.activations import SnakeBeta  # noqa: E402
from oracle import torch_path as TP  # noqa: E402


def bench_shape(B, C, T, dtype, which="fwd", iters=20, nbuf=8):
    dev = torch.device("cuda:0")
    act = SnakeBeta(C, alpha_logscale=True)
    with torch.no_grad():
        act.alpha.normal_(0, 0.5)
        act.beta.normal_(0, 0.5)
    m = Activation1d(activation=act).to(dev)
    xs = [torch.randn(B, C, T, device=dev).to(dtype) for _ in range(nbuf)]
    ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
    from afa_b200 import functional as Fn
    tu, td = m._host_taps()
    a, b = m.act.alpha.detach(), m.act.beta.detach()
    if which == "fwd":
        fn = lambda i: Fn.activation1d_forward_raw(xs[i % nbuf], a, b, tu, td, True, out=ys[i % nbuf])
        bpe = 2 * xs[0].element_size()
    else:
        fn = lambda i: Fn.activation1d_backward_raw(xs[i % nbuf], ys[(i + 1) % nbuf], a, b, tu, td, True)
        for y in ys:
            y.normal_()
        bpe = 3 * xs[0].element_size()
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    n = B * C * T
    return us, n * bpe / us / 1e3  # GB/s


if __name__ == "__main__":
    ge.smoke()
    print(torch.cuda.get_device_name(0))
    for dt in (torch.float32, torch.bfloat16):
        for which in ("fwd", "bwd"):
            for ch in (5, 9):
                _lib.set_tuning(0 if which == "fwd" else 1, ch, 0)
                for (B, C, T) in [(2, 512, 8192), (2, 768, 3444), (2, 24, 220416), (16, 768, 3444), (16, 96, 55104), (16, 24, 220416)]:
                    if dt == torch.bfloat16 and T % 8:
                        tag = "(unaligned)"
                    else:
                        tag = ""
                    nbuf = max(3, int(600e6 // (B * C * T * 4)))
                    us, gbs = bench_shape(B, C, T, dt, which, nbuf=min(nbuf, 16))
                    print(f"{str(dt):16s} {which} ch={ch} B={B:3d} C={C:4d} T={T:7d} {us:9.1f} us {gbs:8.1f} GB/s {tag}", flush=True)
            _lib.set_tuning(0 if which == "fwd" else 1, 0, 0)
    print(_lib.kernel_info(0, 0, 8192), _lib.kernel_info(1, 0, 8192), _lib.kernel_info(0, 1, 8192), _lib.kernel_info(1, 1, 8192))
