"""The multi-scale mel loss step of BASELINE config 5 (batch 32, segment 8192): forward + backward to d loss / d estimate,
`ours` (fused kernels) or `torch` (the reference's op chain), for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/mel_loss_step.py ours"""
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")]
from afa_b200 import mel as P  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ours"
dev = "cuda:0"
torch.manual_seed(0)
y = (0.3 * torch.randn(32, 1, 8192, device=dev)).clamp(-1, 1)
yh = (y + 0.05 * torch.randn_like(y)).clamp(-1, 1)
msl = P.MultiScaleMelSpectrogramLoss(22050)
wins, nms = msl.window_lengths, msl.n_mels
bases = [torch.from_numpy(P.slaney_mel_filterbank(22050, w, nm)).to(dev) for w, nm in zip(wins, nms)]
hann = {w: torch.hann_window(w, device=dev) for w in wins}
log10 = torch.log(torch.tensor(10.0))


def logmels(wav, basis, n):       # loss.py:110-167 + :195-197
    stft = torch.stft(wav.reshape(-1, wav.shape[-1]), n_fft=n, hop_length=n // 4, window=hann[n], return_complex=True, center=True)
    mels = (torch.abs(stft).transpose(1, 2) @ basis.T).transpose(1, 2)
    return torch.log(mels.clamp(min=1e-5)) / log10


def step():
    x = yh.detach().requires_grad_(True)
    if which == "ours":
        loss = msl(x, y)
    else:
        loss = sum(torch.nn.functional.l1_loss(logmels(x, b, w), logmels(y, b, w)) for b, w in zip(bases, wins))
    loss.backward()
    return loss, x.grad


for _ in range(2):
    loss, g = step()
torch.cuda.synchronize()
print(which, float(loss), float(g.abs().max()))
