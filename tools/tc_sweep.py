"""GPU: bf16 forward, tensor-core kernel (per blocks-per-lane NY) against the register-walk kernel, per model shape.
L2-cold (rotating buffers > L2), CUDA-graph replay.  usage: python tools/tc_sweep.py [quick]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: F401  (sys.path)
import torch
from afa_b200 import Activation1d, _lib, functional as Fn
from afa_b200.activations import SnakeBeta

dev = torch.device("cuda:0")
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
shapes = []
for B in (2, 16):
    shapes += [(B, 768, 3440), (B, 384, 13776), (B, 192, 27552), (B, 96, 55104), (B, 48, 110208), (B, 24, 220416)]
shapes += [(32, 768, 128), (32, 384, 512), (32, 192, 1024), (32, 96, 2048), (32, 48, 4096), (32, 24, 8192), (2, 512, 8192)]
dtype = torch.bfloat16


def time_cfg(b, c, t):
    m = Activation1d(activation=SnakeBeta(c, alpha_logscale=True)).to(dev)
    n = b * c * t
    nbuf = max(2, min(24, int(0.6e9 // (n * 2 * 2))))
    xs = [torch.randn(b, c, t, device=dev).to(dtype) for _ in range(nbuf)]
    ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
    tu, td = m._host_taps()
    a_, b_ = m.act.alpha.detach(), m.act.beta.detach()
    Fn.activation1d_forward_raw(xs[0], a_, b_, tu, td, True, out=ys[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(2 * nbuf):
            Fn.activation1d_forward_raw(xs[i % nbuf], a_, b_, tu, td, True, out=ys[i % nbuf])
    # sustained figures: ~0.2 s of replays before the timed ~0.2 s (the board reaches its power cap within ~0.1 s, and the
    # clocks it then settles on are what a long step sees)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    reps = max(3, int(200.0 / max(e0.elapsed_time(e1), 1e-3)))
    for _ in range(reps):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * 2 * nbuf)


for (b, c, t) in shapes:
    n = b * c * t
    line = f"bf16 B={b:3d} C={c:4d} T={t:7d}"
    _lib.set_tuning(5, 0, 0)
    us = time_cfg(b, c, t)
    line += f" | walk {us:7.1f} us {n * 4 / us / 1e3:6.0f} GB/s"
    for ny in ((0,) if quick else (4, 8, 16, 32, 0)):
        _lib.set_tuning(5, 2, ny)
        us = time_cfg(b, c, t)
        line += f" | tc ny{ny}: {us:7.1f} {n * 4 / us / 1e3:6.0f}"
    print(line, flush=True)
_lib.set_tuning(5, 1, 0)
