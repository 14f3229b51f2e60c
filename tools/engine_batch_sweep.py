"""GPU: CUDA-graph replay time of the bf16 channels-last engine per clips-per-batch (sustained: ~1 s of replays per point).
usage: python tools/engine_batch_sweep.py [clips ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
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
for par in (True, False):
    eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16, parallel_resblocks=par)
    for clips in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
        ge = GraphedEngine(eng, 2 * clips, 861, want_pcm=True, pcm_interleave=2)
        mel = torch.rand(2 * clips, 80, 861, device=dev) * 14.5 - 12
        for _ in range(3):
            ge(mel)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(5, int(200 / clips))
        e0.record()
        for _ in range(reps):
            ge(mel)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(f"parallel_resblocks={par} clips/batch {clips}: {ms:8.3f} ms per pass, {ms / clips:6.3f} ms per clip, {clips * 10.0 / (ms * 1e-3):7.1f} audio-s/s", flush=True)
        del ge
        torch.cuda.empty_cache()
