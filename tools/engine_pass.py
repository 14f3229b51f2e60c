"""Eager passes of the channels-last engine (bf16, `clips` binaural clips); the LAST pass is bracketed by
cudaProfilerStart/Stop, so `ncu --profile-from-start off` lists exactly one warm pass (cuDNN's autotuning of the
first pass stays outside).  This is the command the launch list in profiles/ is taken from.
usage: python tools/engine_pass.py [clips]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200.vocoder import BigVGANGenerator
from afa_b200.engine import ChannelsLastVocoder

clips = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
torch.manual_seed(1234)
gen = BigVGANGenerator().to(dev)
with torch.no_grad():
    for n, p in gen.named_parameters():
        if n.endswith("alpha") or n.endswith("beta"):
            p.normal_(0, 0.5)
gen = gen.bfloat16().eval()
eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16, parallel_resblocks=False)
mel = torch.rand(2 * clips, 80, 861, device=dev) * 14.5 - 12
for _ in range(int(os.environ.get("PASSES", "3")) - 1):
    wave, pcm = eng(mel, want_pcm=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
wave, pcm = eng(mel, want_pcm=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", tuple(wave.shape), tuple(pcm.shape))
