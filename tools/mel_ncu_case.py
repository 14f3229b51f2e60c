"""One workload for `ncu --set full -k regex:afa_logmel`: the single-scale log mel of BASELINE config 5's batch
([32, 8192], n_fft 1024, hop 256, 80 mels) and the 2048-point scale of the multi-scale loss, three launches each."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")]
from afa_b200 import mel as P  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
y = (0.3 * torch.randn(32, 8192, device=dev)).clamp(-1, 1)
plan = P.MelPlan(1024, torch.hann_window(1024, dtype=torch.float64), P.slaney_mel_filterbank(22050, 1024, 80), dev)
plan2 = P.MelPlan(2048, torch.hann_window(2048, dtype=torch.float64), P.slaney_mel_filterbank(22050, 2048, 320), dev)
for _ in range(3):
    P.logmel(y, plan, 256, 384)
    P.logmel(y, plan2, 512, 1024, mag_eps=0.0, log_scale=0.4342944819)
torch.cuda.synchronize()
print("ok")
