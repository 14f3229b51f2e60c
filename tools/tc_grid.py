"""Bring-up: the pytest grid of the fused kernel, printing the first failing case (tcgen05 path vs mma.sync path)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import numpy as np
import torch
from afa_b200 import _lib, functional as F_afa, functional_cl as FC
from afa_b200.modules import kaiser_sinc_filter1d

dev = torch.device("cuda:0")
dt = torch.bfloat16
h = F_afa.host_taps(kaiser_sinc_filter1d(0.25, 0.3, 12))
rng = np.random.default_rng(11)
n = 0
for C in (8, 16, 24, 32, 48, 64):
    for (k, d) in ((3, 1), (3, 5), (7, 3), (11, 1), (11, 5)):
        for T in (1, 17, 160, 415, 1000, 1733):
            B = 1 + n % 2
            with_res = (n % 2) == 0
            n += 1
            x = torch.tensor(rng.standard_normal((B, T, C)), dtype=dt, device=dev)
            kw = dict(res=torch.randn(B, T, C, device=dev, dtype=dt), xsum=torch.empty(B, T, C, device=dev, dtype=dt)) if with_res else {}
            alpha, beta, bias = (torch.randn(C, device=dev) * 0.5 for _ in range(3))
            w = (torch.randn(k, C, C, device=dev) / (k * C) ** 0.5).to(dt)
            tag = f"C={C} k={k} d={d} T={T} B={B} res={with_res}"
            try:
                _lib.set_tuning(3, int(os.environ.get('TC_PATH', '1')))
                _lib.set_tuning(4, int(os.environ.get('TC_XS', '1')))
                y = FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, w, k, d, bias=bias, **kw)
                torch.cuda.synchronize()
                _lib.set_tuning(3, 0)
                y0 = FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, w, k, d, bias=bias, **kw)
                torch.cuda.synchronize()
                diff = float((y.float() - y0.float()).abs().max())
                if diff != 0.0:
                    print("DIFF", tag, diff, flush=True)
            except Exception as e:
                print("ERR", tag, repr(e)[:160], flush=True)
                sys.exit(1)
print("grid done", n)
