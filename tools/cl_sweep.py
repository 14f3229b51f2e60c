"""GPU: per-shape time of the channels-last activation kernels (plain+bias, residual variant) next to the [B, C, T]
kernel, bf16 (and fp32 with --fp32); L2-cold (rotating buffers) and L2-warm; optional segment-length sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200 import _lib, functional as F_afa, functional_cl as FC
from afa_b200.modules import kaiser_sinc_filter1d

dev = torch.device("cuda:0")
dtype = torch.float32 if "--fp32" in sys.argv else torch.bfloat16
B = int(os.environ.get("SWEEP_B", "8"))
Ls = [int(v) for v in os.environ.get("SWEEP_L", "0").split(",")]
h = F_afa.host_taps(kaiser_sinc_filter1d(0.25, 0.3, 12))
taps = (h, h)
stages = [(768, 3444), (384, 13776), (192, 27552), (96, 55104), (48, 110208), (24, 220416)]
esz = 2 if dtype == torch.bfloat16 else 4


def timeit(fn, nsets, n=12):
    for i in range(3):
        fn(i % nsets)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i % nsets)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us


for C, T in stages:
    nsets = 4
    xs = [torch.randn(B, T, C, device=dev, dtype=dtype) for _ in range(nsets)]
    rs = [torch.randn(B, T, C, device=dev, dtype=dtype) for _ in range(nsets)]
    ys = [torch.empty(B, T, C, device=dev, dtype=dtype) for _ in range(nsets)]
    ss = [torch.empty(B, T, C, device=dev, dtype=dtype) for _ in range(nsets)]
    xn = [x.view(B, C, T) for x in xs]          # same bytes, read as [B, C, T] by the NCW kernel
    yn = [y.view(B, C, T) for y in ys]
    alpha = torch.randn(C, device=dev) * 0.5
    beta = torch.randn(C, device=dev) * 0.5
    bias = torch.randn(C, device=dev) * 0.3
    n_el = B * C * T
    t_ncw = timeit(lambda i: F_afa.activation1d_forward_raw(xn[i], alpha, beta, h, h, True, out=yn[i]), nsets)
    t_ncw_w = timeit(lambda i: F_afa.activation1d_forward_raw(xn[0], alpha, beta, h, h, True, out=yn[0]), 1)
    line = f"C={C:4d} T={T:6d} B={B}: NCW {t_ncw:7.1f} us ({n_el * 2 * esz / t_ncw / 1e3:6.0f} GB/s) warm {t_ncw_w:7.1f}"
    for L in Ls:
        _lib.set_tuning(2, L // 12)
        t_p = timeit(lambda i: FC.amp_activation1d_cl(xs[i], T, alpha, beta, h, h, True, bias=bias, out=ys[i]), nsets)
        t_pw = timeit(lambda i: FC.amp_activation1d_cl(xs[0], T, alpha, beta, h, h, True, bias=bias, out=ys[0]), 1)
        t_r = timeit(lambda i: FC.amp_activation1d_cl(xs[i], T, alpha, beta, h, h, True, bias=bias, res=rs[i], xsum=ss[i], out=ys[i]), nsets)
        t_rw = timeit(lambda i: FC.amp_activation1d_cl(xs[0], T, alpha, beta, h, h, True, bias=bias, res=rs[0], xsum=ss[0], out=ys[0]), 1)
        line += f" | L={L or 'auto'}: CL {t_p:7.1f} us ({n_el * 2 * esz / t_p / 1e3:6.0f} GB/s) warm {t_pw:7.1f}; RES {t_r:7.1f} us ({n_el * 4 * esz / t_r / 1e3:6.0f} GB/s) warm {t_rw:7.1f}"
    _lib.set_tuning(2, 0)
    print(line, flush=True)
    del xs, rs, ys, ss
    torch.cuda.empty_cache()
