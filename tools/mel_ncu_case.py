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
if len(sys.argv) > 1 and sys.argv[1] == "small":      # the 32-point scale of the loss: 32 frames in flight per CTA
    plan32 = P.MelPlan(32, torch.hann_window(32, dtype=torch.float64), P.slaney_mel_filterbank(22050, 32, 5), dev)
    cfg32 = (plan32, 8, 16, P.AFA_MEL_PAD_REFLECT, 0.0, 1e-5, 0.4342944819, False)
    y = (0.3 * torch.randn(32, 8192, device=dev)).clamp(-1, 1)
    for _ in range(2):
        out = P.logmel_forward_raw(y, *cfg32)
        P.logmel_backward_raw(y, torch.ones_like(out), *cfg32)
    torch.cuda.synchronize()
    print("ok small")
    sys.exit(0)
for shape in ((32, 8192), (16, 220416)):
    y = (0.3 * torch.randn(*shape, device=dev)).clamp(-1, 1)
    for _ in range(2):
        out = P.logmel_forward_raw(y, *cfg)
        P.logmel_backward_raw(y, torch.ones_like(out), *cfg)
torch.cuda.synchronize()
print("ok")
