"""Two launches each of the fused activation -> convolution kernel (plain and residual variant) for an ncu capture.
usage: python tools/actconv_ncu_case.py C T B k d"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200 import functional as F_afa, functional_cl as FC
from afa_b200.modules import kaiser_sinc_filter1d

C, T, B, k, d = (int(v) for v in sys.argv[1:6])
dev = torch.device("cuda:0")
dt = torch.bfloat16
h = F_afa.host_taps(kaiser_sinc_filter1d(0.25, 0.3, 12))
x = torch.randn(B, T, C, device=dev, dtype=dt)
r = torch.randn(B, T, C, device=dev, dtype=dt)
s = torch.empty_like(x)
alpha, beta, bias = (torch.randn(C, device=dev) * 0.5 for _ in range(3))
w = (torch.randn(k, C, C, device=dev) / (k * C) ** 0.5).to(dt)
for _ in range(2):
    FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, w, k, d, bias=bias)
    FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, w, k, d, bias=bias, res=r, xsum=s)
torch.cuda.synchronize()
