"""GPU: whole-generator time, [B, C, T] harness (graph) vs channels-last engine (graph), bf16, + kernel breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200.vocoder import BigVGANGenerator, GraphedVocoder
from afa_b200.engine import ChannelsLastVocoder, GraphedEngine

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234)
gen = BigVGANGenerator().to(dev)
with torch.no_grad():
    for n, p in gen.named_parameters():
        if n.endswith("alpha") or n.endswith("beta"):
            p.normal_(0, 0.5)
gen = gen.bfloat16().eval()


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for clips in (int(c) for c in os.environ.get("PROBE_CLIPS", "1,4").split(",")):
    B = 2 * clips
    mel = torch.rand(B, 80, 861, device=dev) * 14.5 - 12
    eng0 = ChannelsLastVocoder(gen, dtype=torch.bfloat16, fuse_narrow_convs=False)
    ge0 = GraphedEngine(eng0, B, 861, want_pcm=True)
    t_e0 = timeit(lambda: ge0(mel))
    w_e0 = ge0(mel)[0].clone()
    del ge0
    print(f"clips={clips}: unfused narrow convs {t_e0:.2f} ms ({clips * 10 / t_e0 * 1e3:.0f} audio-s/s)", flush=True)
    eng1 = ChannelsLastVocoder(gen, dtype=torch.bfloat16, parallel_resblocks=False)
    ge1 = GraphedEngine(eng1, B, 861, want_pcm=True)
    t_e1 = timeit(lambda: ge1(mel))
    w_e1 = ge1(mel)[0].clone()
    del ge1
    eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
    ge = GraphedEngine(eng, B, 861, want_pcm=True)
    t_e = timeit(lambda: ge(mel))
    w_e = ge(mel)[0].clone()
    print(f"clips={clips}: serial resblocks {t_e1:.2f} ms, parallel {t_e:.2f} ms, identical {bool(torch.equal(w_e, w_e1))}; "
          f"fused vs unfused max|diff| {float((w_e - w_e0).abs().max()):.3e} (max |w| {float(w_e0.abs().max()):.3e})", flush=True)
    if os.environ.get("PROBE_NCW", "1") == "1":
        gv = GraphedVocoder(gen, B, 861, dtype=torch.bfloat16, device=dev)
        t_v = timeit(lambda: gv(mel.bfloat16()))
        w_v = gv(mel.bfloat16()).float()
        print(f"clips={clips}: NCW harness {t_v:.2f} ms ({clips * 10 / t_v * 1e3:.0f} audio-s/s)  CL engine {t_e:.2f} ms ({clips * 10 / t_e * 1e3:.0f} audio-s/s)"
              f"  max|diff| {float((w_e - w_v).abs().max()):.3e} (max |w| {float(w_v.abs().max()):.3e})", flush=True)
        del gv
    else:
        print(f"clips={clips}: CL engine {t_e:.2f} ms ({clips * 10 / t_e * 1e3:.0f} audio-s/s)", flush=True)
    if clips == 4:
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            eng(mel, want_pcm=True)
            torch.cuda.synchronize()
        rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:14]
        tot = sum(r.device_time_total for r in prof.key_averages())
        for r in rows:
            print(f"{r.device_time_total / 1e3:9.2f} ms {100 * r.device_time_total / tot:5.1f}% x{r.count:4d} {r.key[:100]}")
    del ge, eng
    torch.cuda.empty_cache()
