"""One workload for `ncu --set full -k regex:afa_logmel`: the single-scale log mel of BASELINE config 5's batch
([32, 8192], n_fft 1024, hop 256, 80 mels), forward and backward, and the same for 16 ten-second clips."""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")]
from afa_b200 import mel as P  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
plan = P.MelPlan(1024, torch.hann_window(1024, dtype=torch.float64), P.slaney_mel_filterbank(22050, 1024, 80), dev)
cfg = (plan, 256, 384, P.AFA_MEL_PAD_REFLECT, 1e-9, 1e-5, 1.0, False)
for shape in ((32, 8192), (16, 220416)):
    y = (0.3 * torch.randn(*shape, device=dev)).clamp(-1, 1)
    for _ in range(2):
        out = P.logmel_forward_raw(y, *cfg)
        P.logmel_backward_raw(y, torch.ones_like(out), *cfg)
torch.cuda.synchronize()
print("ok")
