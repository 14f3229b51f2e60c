"""GPU: channels-last engine (CUDA graph, bf16, 4 clips) with the channels-last tensor-core activation off / on (same box):
whole-pass time, audio-s/s, output difference, kernel breakdown of the 'on' pass.  usage: python tools/engine_tc_cl_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200 import _lib
from afa_b200.vocoder import BigVGANGenerator
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


def timeit(fn, n=10):
    for _ in range(3):
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
    res = {}
    for mode in (0, 1, 0, 1):
        _lib.set_tuning(7, mode, 0)
        eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
        ge = GraphedEngine(eng, B, 861, want_pcm=True)
        t = timeit(lambda: ge(mel))
        w = ge(mel)[0].clone()
        print(f"clips={clips} tc_cl={mode}: {t:.3f} ms ({clips * 10 / t * 1e3:.0f} audio-s/s)", flush=True)
        res[mode] = w
        if mode == 1 and clips == 4:
            with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
                eng(mel, want_pcm=True)
                torch.cuda.synchronize()
            rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)[:12]
            tot = sum(r.device_time_total for r in prof.key_averages())
            for r in rows:
                print(f"{r.device_time_total / 1e3:9.2f} ms {100 * r.device_time_total / tot:5.1f}% x{r.count:4d} {r.key[:100]}")
        del ge, eng
        torch.cuda.empty_cache()
    d = (res[0].float() - res[1].float()).abs().max().item()
    print(f"clips={clips}: max |wave(on) - wave(off)| = {d:.3e} (max |wave| {res[0].float().abs().max().item():.3e})", flush=True)
_lib.set_tuning(7, 1, 0)
