"""GPU probe: per-layer device time of the whole generator (which convolution is the slow fallback?),
and what cuDNN does with channels-last / bias-free calls for the AMP convolution shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
import torch.nn as nn
import torch.nn.functional as F
from afa_b200.vocoder import BigVGANGenerator

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
B = int(os.environ.get("PROBE_B", "8"))


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


torch.manual_seed(0)
gen = BigVGANGenerator().to(dev).bfloat16().eval()
mel = (torch.rand(B, 80, 861, device=dev) * 14.5 - 12).bfloat16()
recs = {}
names = {m: n for n, m in gen.named_modules()}


def pre(m, inp):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    m._ev0 = ev


def post(m, inp, out):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    recs.setdefault(names[m], []).append((m._ev0, ev, tuple(inp[0].shape), tuple(out.shape)))


from afa_b200.modules import Activation1d
for m in gen.modules():
    if isinstance(m, (nn.Conv1d, nn.ConvTranspose1d, Activation1d)):
        m.register_forward_pre_hook(pre)
        m.register_forward_hook(post)
with torch.no_grad():
    for _ in range(3):
        recs.clear()
        e_all0 = torch.cuda.Event(enable_timing=True); e_all1 = torch.cuda.Event(enable_timing=True)
        e_all0.record()
        gen(mel)
        e_all1.record()
        torch.cuda.synchronize()
total = e_all0.elapsed_time(e_all1)
rows = []
for n, lst in recs.items():
    e0, e1, si, so = lst[-1]
    rows.append((e0.elapsed_time(e1), n, si, so))
rows.sort(reverse=True)
print(f"eager pass B={B}: {total:.2f} ms; hooked layers sum {sum(r[0] for r in rows):.2f} ms")
for ms, n, si, so in rows[:25]:
    print(f"{ms:8.3f} ms  {n:34s} {si} -> {so}")
by_kind = {}
for ms, n, si, so in rows:
    k = "act" if "activation" in n else ("ups" if n.startswith("ups") else ("conv_pre/post" if n.startswith("conv_p") else "resblock conv"))
    by_kind[k] = by_kind.get(k, 0.0) + ms
print("by kind:", {k: round(v, 2) for k, v in by_kind.items()})

# ---- channels-last / bias-free experiments on the AMP convolution shapes
print("\nconv experiments (bf16): ms per call")
stages = [(768, 3444), (384, 13776), (192, 27552), (96, 55104), (48, 110208), (24, 220416)]
for C, T in stages:
    for k, d in ((3, 1), (7, 3), (11, 5)):
        x = torch.randn(B, C, T, device=dev, dtype=torch.bfloat16)
        w = torch.randn(C, C, k, device=dev, dtype=torch.bfloat16) * 0.01
        b = torch.randn(C, device=dev, dtype=torch.bfloat16)
        p = (k * d - d) // 2
        with torch.no_grad():
            t_bias = timeit(lambda: F.conv1d(x, w, b, 1, p, d))
            t_nobias = timeit(lambda: F.conv1d(x, w, None, 1, p, d))
            x4 = x.unsqueeze(2).contiguous(memory_format=torch.channels_last)
            w4 = w.unsqueeze(2).contiguous(memory_format=torch.channels_last)
            t_cl = timeit(lambda: F.conv2d(x4, w4, None, 1, (0, p), (1, d)))
            y4 = F.conv2d(x4, w4, None, 1, (0, p), (1, d))
            y = F.conv1d(x, w, None, 1, p, d)
            err = (y4.squeeze(2).float() - y.float()).abs().max().item()
            cl_out = y4.is_contiguous(memory_format=torch.channels_last)
        flops = 2.0 * B * C * C * k * T
        print(f"C={C:4d} T={T:6d} k={k:2d} d={d}: bias {t_bias:7.3f}  nobias {t_nobias:7.3f}  channels_last {t_cl:7.3f} "
              f"(out CL {cl_out}, maxdiff {err:.2e})  -> {flops / t_cl / 1e9:7.1f} TF/s CL, {flops / t_nobias / 1e9:7.1f} TF/s NCW")
        del x, w, x4, w4, y4, y
# conv_post and the last upsampler in isolation
x = torch.randn(B, 24, 220416, device=dev, dtype=torch.bfloat16)
w = torch.randn(1, 24, 7, device=dev, dtype=torch.bfloat16) * 0.01
with torch.no_grad():
    print("conv_post 24->1 k7 bf16:", round(timeit(lambda: F.conv1d(x, w, None, 1, 3)), 3), "ms;  fp32:",
          round(timeit(lambda: F.conv1d(x.float(), w.float(), None, 1, 3)), 3), "ms (incl. casts)")
for i, (cin, t, k, u) in enumerate([(1536, 861, 8, 4), (768, 3444, 8, 4), (384, 13776, 4, 2), (192, 27552, 4, 2), (96, 55104, 4, 2), (48, 110208, 4, 2)]):
    x = torch.randn(B, cin, t, device=dev, dtype=torch.bfloat16)
    w = torch.randn(cin, cin // 2, k, device=dev, dtype=torch.bfloat16) * 0.01
    with torch.no_grad():
        print(f"ups[{i}] {cin}->{cin // 2} k{k} s{u} T={t}:", round(timeit(lambda: F.conv_transpose1d(x, w, None, u, (k - u) // 2)), 3), "ms")
