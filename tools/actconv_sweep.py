"""GPU: fused activation->convolution kernel against (activation kernel + cuDNN polyphase convolution), per model
shape of the narrow stages, bf16."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
import torch.nn as nn
from afa_b200 import _lib, functional as F_afa, functional_cl as FC
from afa_b200.engine import _Conv
from afa_b200.modules import kaiser_sinc_filter1d

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
B = int(os.environ.get("SWEEP_B", "8"))
h = F_afa.host_taps(kaiser_sinc_filter1d(0.25, 0.3, 12))
dt = torch.bfloat16
_lib.set_tuning(3, int(os.environ.get('TC_PATH', '1')))
_lib.set_tuning(4, int(os.environ.get('TC_XS', '1')))


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for C, T in ((96 if os.environ.get("SWEEP_96") else 48, 55104 if os.environ.get("SWEEP_96") else 110208), (48, 110208), (24, 220416)):
    if C > 64:
        continue
    x = torch.randn(B, T, C, device=dev, dtype=dt)
    r = torch.randn(B, T, C, device=dev, dtype=dt)
    s = torch.empty_like(x)
    alpha, beta, bias = (torch.randn(C, device=dev) * 0.5 for _ in range(3))
    for (k, d) in ((3, 1), (3, 3), (3, 5), (7, 1), (7, 3), (7, 5), (11, 1), (11, 3), (11, 5)):
        conv = nn.Conv1d(C, C, k, 1, dilation=d, padding=(k * d - d) // 2).to(dev)
        cu = _Conv(conv, dt, fuse=False)
        cf = _Conv(conv, dt, fuse=True)
        tp = cu.tpad(T)
        t_act = timeit(lambda: FC.amp_activation1d_cl(x, T, alpha, beta, h, h, True, bias=bias, out_tpad=tp))
        a = FC.amp_activation1d_cl(x, T, alpha, beta, h, h, True, bias=bias, out_tpad=tp)
        t_conv = timeit(lambda: cu(a))
        t_f = timeit(lambda: FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, cf.w_kcc, k, d, bias=bias))
        t_fr = timeit(lambda: FC.amp_act_conv_cl(x, T, alpha, beta, h, h, True, cf.w_kcc, k, d, bias=bias, res=r, xsum=s))
        t_actr = timeit(lambda: FC.amp_activation1d_cl(x, T, alpha, beta, h, h, True, bias=bias, res=r, xsum=s, out_tpad=tp))
        print(f"C={C} k={k:2d} d={d}: act {t_act:6.1f} + cudnn {t_conv:6.1f} = {t_act + t_conv:6.1f} us | fused {t_f:6.1f} us || "
              f"res: act {t_actr:6.1f} + cudnn = {t_actr + t_conv:6.1f} | fused {t_fr:6.1f}", flush=True)
