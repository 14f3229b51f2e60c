"""One launch each of the channels-last kernels on a model shape, for an ncu capture (no timing here).
usage: python tools/cl_ncu_case.py C T B L [fp32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200 import _lib, functional as F_afa, functional_cl as FC
from afa_b200.modules import kaiser_sinc_filter1d

C, T, B, L = (int(v) for v in sys.argv[1:5])
dtype = torch.float32 if "fp32" in sys.argv else torch.bfloat16
dev = torch.device("cuda:0")
h = F_afa.host_taps(kaiser_sinc_filter1d(0.25, 0.3, 12))
x = torch.randn(B, T, C, device=dev, dtype=dtype)
r = torch.randn(B, T, C, device=dev, dtype=dtype)
y = torch.empty_like(x); s = torch.empty_like(x)
alpha = torch.randn(C, device=dev) * 0.5
beta = torch.randn(C, device=dev) * 0.5
bias = torch.randn(C, device=dev) * 0.3
_lib.set_tuning(2, L // 12)
for _ in range(2):
    FC.amp_activation1d_cl(x, T, alpha, beta, h, h, True, bias=bias, out=y)
    FC.amp_activation1d_cl(x, T, alpha, beta, h, h, True, bias=bias, res=r, xsum=s, out=y)
torch.cuda.synchronize()
