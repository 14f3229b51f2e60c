"""Bring-up probe for the tcgen05 fused kernel: one small call, error reported synchronously."""
import os, sys
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200 import _lib, functional as F_afa, functional_cl as FC
from afa_b200.modules import kaiser_sinc_filter1d

C, T, B, k, d = (int(v) for v in sys.argv[1:6])
RES = len(sys.argv) > 6
dev = torch.device("cuda:0")
dt = torch.bfloat16
h = F_afa.host_taps(kaiser_sinc_filter1d(0.25, 0.3, 12))
torch.manual_seed(0)
x = torch.randn(B, T, C, device=dev, dtype=dt)
alpha, beta, bias = (torch.randn(C, device=dev) * 0.5 for _ in range(3))
w = (torch.randn(k, C, C, device=dev) / (k * C) ** 0.5).to(dt)
try:
    kw = dict(res=torch.randn(B, T, C, device=dev, dtype=dt), xsum=torch.empty(B, T, C, device=dev, dtype=dt)) if RES else {}
    y = FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, w, k, d, bias=bias, **kw)
    torch.cuda.synchronize()
    _lib.set_tuning(3, 0)
    y0 = FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, w, k, d, bias=bias, **kw)
    torch.cuda.synchronize()
    print("OK max|tc - mma| =", float((y.float() - y0.float()).abs().max()), "max|y| =", float(y0.float().abs().max()))
except Exception as e:
    print("ERR", repr(e)[:200])
