"""GPU: two engine graphs (4 clips each) replayed back to back on one stream vs concurrently on two streams (same box):
do the tensor-bound convolutions of one batch overlap the activation kernels of the other?  usage: python tools/engine_two_streams.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa
import torch
from afa_b200.vocoder import BigVGANGenerator
from afa_b200.engine import ChannelsLastVocoder, GraphedEngine

dev = torch.device("cuda:0")
torch.manual_seed(1234)
gen = BigVGANGenerator().to(dev)
with torch.no_grad():
    for n, p in gen.named_parameters():
        if n.endswith("alpha") or n.endswith("beta"):
            p.normal_(0, 0.5)
gen = gen.bfloat16().eval()
clips = int(os.environ.get("PROBE_CLIPS", "4"))
B = 2 * clips
mels = [torch.rand(B, 80, 861, device=dev) * 14.5 - 12 for _ in range(2)]
eng = ChannelsLastVocoder(gen, dtype=torch.bfloat16)
ges = [GraphedEngine(eng, B, 861, want_pcm=True) for _ in range(2)]
streams = [torch.cuda.Stream(dev) for _ in range(2)]
ref = [ges[i](mels[i])[0].clone() for i in range(2)]


def sequential():
    for i in range(2):
        ges[i](mels[i])


def concurrent():
    cur = torch.cuda.current_stream(dev)
    for i in range(2):
        streams[i].wait_stream(cur)
        with torch.cuda.stream(streams[i]):
            ges[i](mels[i])
    for i in range(2):
        cur.wait_stream(streams[i])


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


for rep in range(2):
    ts, tc = timeit(sequential), timeit(concurrent)
    print(f"2 x {clips} clips: one stream {ts:.3f} ms ({2 * clips * 10 / ts * 1e3:.0f} audio-s/s), two streams {tc:.3f} ms ({2 * clips * 10 / tc * 1e3:.0f} audio-s/s)", flush=True)
concurrent(); torch.cuda.synchronize()
print("outputs identical:", all(bool(torch.equal(ges[i](mels[i])[0], ref[i])) for i in range(2)))
