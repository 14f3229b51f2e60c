"""GPU A/B of library builds on the big fp32/bf16 forward shapes with explicit segment lengths.
usage: python tools/ab_big.py lib1.so lib2.so ..."""
import os, subprocess, sys
code = r'''
import os, sys
sys.path.insert(0, os.getcwd()); import __graft_entry__
import torch
from afa_b200 import Activation1d, _lib, functional as Fn
from afa_b200.activations import SnakeBeta
dev = torch.device("cuda:0")
for dtype in (torch.float32, torch.bfloat16):
  for (b, c, t) in [(16,384,13776),(16,96,55104),(16,24,220416),(2,96,55104)]:
    line = f"{str(dtype):15s} B={b:3d} C={c:4d} T={t:7d}"
    for ch in (9, 13, 17):
        _lib.set_tuning(0, ch, 0)
        m = Activation1d(activation=SnakeBeta(c, alpha_logscale=True)).to(dev)
        n = b*c*t; es = 4 if dtype == torch.float32 else 2
        nbuf = max(2, min(16, int(1.0e9 // (n*es*2))))
        xs = [torch.randn(b,c,t,device=dev).to(dtype) for _ in range(nbuf)]; ys = [torch.empty_like(xs[0]) for _ in range(nbuf)]
        tu, td = m._host_taps(); a_, b_ = m.act.alpha.detach(), m.act.beta.detach()
        Fn.activation1d_forward_raw(xs[0], a_, b_, tu, td, True, out=ys[0]); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(2*nbuf): Fn.activation1d_forward_raw(xs[i%nbuf], a_, b_, tu, td, True, out=ys[i%nbuf])
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): g.replay()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1)*1e3/(10*nbuf)
        line += f" | ch{ch}: {us:7.1f} us {n*2*es/us/1e3:6.0f}"
        del xs, ys, g; torch.cuda.empty_cache()
    print(line, flush=True)
print(_lib.kernel_info(0, 0, 1 << 20))
'''
for lib in sys.argv[1:]:
    print("==", lib, flush=True)
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, AFA_LIBRARY=lib))
