"""GPU: where does the whole-generator time go? (cuDNN autotune on/off, batch size, share of the fused activation)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200.vocoder import BigVGANGenerator, GraphedVocoder

dev = torch.device("cuda:0")
for bench_flag in (False, True):
    torch.backends.cudnn.benchmark = bench_flag
    for clips in (1, 4, 8):
        torch.manual_seed(0)
        gen = BigVGANGenerator().to(dev).bfloat16().eval()
        B = 2 * clips
        t0 = time.time()
        gv = GraphedVocoder(gen, B, 861, dtype=torch.bfloat16, device=dev)
        mel = torch.rand(B, 80, 861, device=dev) * 14.5 - 12
        for _ in range(2):
            gv(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            gv(mel)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"cudnn.benchmark={bench_flag} clips/batch={clips} {ms:8.2f} ms/batch {ms / clips:7.2f} ms/clip {clips * 10 / ms * 1e3:8.1f} audio-s/s (setup {time.time() - t0:.1f}s)", flush=True)
        del gv, gen
        torch.cuda.empty_cache()
# share of the activation: profile one eager pass
torch.backends.cudnn.benchmark = True
gen = BigVGANGenerator().to(dev).bfloat16().eval()
mel = (torch.rand(8, 80, 861, device=dev) * 14.5 - 12).bfloat16()
with torch.no_grad():
    for _ in range(3):
        gen(mel)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        gen(mel)
        torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:12]
tot = sum(r.device_time_total for r in prof.key_averages())
for r in rows:
    print(f"{r.device_time_total / 1e3:9.2f} ms {100 * r.device_time_total / tot:5.1f}% x{r.count:4d} {r.key[:90]}")
