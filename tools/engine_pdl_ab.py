"""GPU: channels-last engine (CUDA graph, bf16, 4 clips) and the 109-call bf16 activation step with programmatic dependent launch
of the tensor-core kernels off / on (afa_set_tuning(9, 0 | 1); same box).  usage: python tools/engine_pdl_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200 import _lib
from afa_b200.vocoder import BigVGANGenerator
from afa_b200.engine import ChannelsLastVocoder, GraphedEngine
import bench

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


clips, B = 4, 8
mel = torch.rand(B, 80, 861, device=dev) * 14.5 - 12
waves = {}
for mode in (0, 1, 0, 1):
    _lib.set_tuning(9, mode)
    eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
    ge = GraphedEngine(eng, B, 861, want_pcm=True)
    t = timeit(lambda: ge(mel))
    waves[mode] = ge(mel)[0].clone()
    print(f"engine clips={clips} pdl={mode}: {t:.3f} ms ({clips * 10 / t * 1e3:.0f} audio-s/s)", flush=True)
    del ge, eng
    torch.cuda.empty_cache()
print("engine output identical with and without dependent launch:", bool(torch.equal(waves[0], waves[1])))
for mode in (0, 1, 0, 1):
    _lib.set_tuning(9, mode)
    wl = bench.Workload(dev, 8, bench.T_MEL_10S, torch.bfloat16)
    wl.step_device(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        wl.step_device()
    t = timeit(g.replay)
    print(f"109-call bf16 step (8 clips) pdl={mode}: {t:.3f} ms", flush=True)
    del g, wl
    torch.cuda.empty_cache()
_lib.set_tuning(9, 1)
