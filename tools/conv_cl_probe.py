"""GPU probe: cuDNN on channels-last [B, T, C] activations.
 (1) dilated Conv1d as a NON-dilated (k x 1) 2-D convolution over the polyphase view [B, T/d, d, C]
     (t = d*h + w; the k taps step along h) -- zero-copy when T % d == 0;
 (2) ConvTranspose1d upsamplers and conv_pre / conv_post in channels-last;
 numerics are checked against F.conv1d on the NCW tensor."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

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


stages = [(768, 3444), (384, 13776), (192, 27552), (96, 55104), (48, 110208), (24, 220416)]
tot = {"ncw_bias": 0.0, "cl": 0.0, "poly": 0.0}
for C, T in stages:
    Tp = (T + 14) // 15 * 15
    for k in (3, 7, 11):
        for d, n_calls in ((1, 4), (3, 1), (5, 1)):       # convs2 (d=1) x3 + convs1 d=1,3,5
            xbtc = torch.zeros(B, Tp, C, device=dev, dtype=torch.bfloat16)
            xbtc[:, :T] = torch.randn(B, T, C, device=dev, dtype=torch.bfloat16)
            w = torch.randn(C, C, k, device=dev, dtype=torch.bfloat16) * 0.02
            b = torch.randn(C, device=dev, dtype=torch.bfloat16)
            p = (k * d - d) // 2
            x_ncw = xbtc[:, :T].permute(0, 2, 1).contiguous()
            # plain channels-last: logical [B, C, 1, T] over memory [B, T, C]
            x4 = xbtc[:, :T].permute(0, 2, 1).unsqueeze(2)
            w4 = w.unsqueeze(2).contiguous(memory_format=torch.channels_last)
            # polyphase: logical [B, C, Tp/d, d] over the same memory, kernel (k, 1), no dilation
            x5 = xbtc.view(B, Tp // d, d, C).permute(0, 3, 1, 2)
            w5 = w.unsqueeze(3).contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                t0 = timeit(lambda: F.conv1d(x_ncw, w, b, 1, p, d))
                t1 = timeit(lambda: F.conv2d(x4, w4, None, 1, (0, p), (1, d)))
                t2 = timeit(lambda: F.conv2d(x5, w5, None, 1, (k // 2, 0), 1))
                ref = F.conv1d(x_ncw, w, None, 1, p, d).float()
                y4 = F.conv2d(x4, w4, None, 1, (0, p), (1, d))
                y5 = F.conv2d(x5, w5, None, 1, (k // 2, 0), 1)
                e4 = (y4.squeeze(2).float() - ref).abs().max().item()
                y5f = y5.permute(0, 2, 3, 1).reshape(B, Tp, C)[:, :T].permute(0, 2, 1).float()
                e5 = (y5f - ref).abs().max().item()
                cl5 = y5.permute(0, 2, 3, 1).is_contiguous()
            fl = 2.0 * B * C * C * k * T
            tot["ncw_bias"] += n_calls * t0
            tot["cl"] += n_calls * t1
            tot["poly"] += n_calls * min(t1, t2)
            print(f"C={C:4d} T={T:6d} k={k:2d} d={d}: NCW+bias {t0:6.3f}  CL {t1:6.3f} ({fl / t1 / 1e9:6.0f} TF/s)  polyphase {t2:6.3f} ({fl / t2 / 1e9:6.0f} TF/s)"
                  f"  err CL {e4:.1e} poly {e5:.1e} (ref max {ref.abs().max().item():.2f}) out-contig {cl5}", flush=True)
            del xbtc, x_ncw, x4, x5, y4, y5, ref, y5f
    torch.cuda.empty_cache()
print("sum over one resblock-triple per stage x calls (ms):", {k: round(v, 2) for k, v in tot.items()})

# upsamplers + pre/post in channels-last
for i, (cin, t, k, u) in enumerate([(1536, 861, 8, 4), (768, 3444, 8, 4), (384, 13776, 4, 2), (192, 27552, 4, 2), (96, 55104, 4, 2), (48, 110208, 4, 2)]):
    x = torch.randn(B, cin, t, device=dev, dtype=torch.bfloat16)
    w = torch.randn(cin, cin // 2, k, device=dev, dtype=torch.bfloat16) * 0.02
    x4 = x.unsqueeze(2).contiguous(memory_format=torch.channels_last)
    w4 = w.unsqueeze(2).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        t0 = timeit(lambda: F.conv_transpose1d(x, w, None, u, (k - u) // 2))
        t1 = timeit(lambda: F.conv_transpose2d(x4, w4, None, (1, u), (0, (k - u) // 2)))
        y = F.conv_transpose2d(x4, w4, None, (1, u), (0, (k - u) // 2))
        err = (y.squeeze(2).float() - F.conv_transpose1d(x, w, None, u, (k - u) // 2).float()).abs().max().item()
    print(f"ups[{i}] {cin}->{cin // 2} k{k} s{u} T={t}: NCW {t0:.3f} ms  CL {t1:.3f} ms  out CL {y.is_contiguous(memory_format=torch.channels_last)} err {err:.1e}")
x = torch.randn(B, 80, 861, device=dev, dtype=torch.bfloat16)
w = torch.randn(1536, 80, 7, device=dev, dtype=torch.bfloat16) * 0.02
x4 = x.unsqueeze(2).contiguous(memory_format=torch.channels_last)
w4 = w.unsqueeze(2).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    print("conv_pre NCW", round(timeit(lambda: F.conv1d(x, w, None, 1, 3)), 3), "CL", round(timeit(lambda: F.conv2d(x4, w4, None, 1, (0, 3))), 3))
