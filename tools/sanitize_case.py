"""Small fused-Activation1d workload for compute-sanitizer (memcheck / racecheck): every kernel variant
(fp32 / bf16; 16-byte aligned, half-aligned, unaligned; forward + backward; tiles spanning rows; several
tiles per persistent warp), checked against the torch-op oracle.
usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: F401
import torch

from afa_b200 import Activation1d
from afa_b200.activations import SnakeBeta
from oracle import torch_path as TP

dev = torch.device("cuda:0")
torch.manual_seed(0)
big = os.environ.get("AFA_SANITIZE_BIG", "1") == "1"
cases = [(torch.float32, 2, 3, 100), (torch.float32, 2, 5, 37), (torch.bfloat16, 2, 3, 200), (torch.bfloat16, 3, 5, 36),
         (torch.bfloat16, 1, 3, 77), (torch.float32, 4, 24, 1000)]
if big:
    cases += [(torch.float32, 8, 96, 5508), (torch.bfloat16, 8, 96, 5508)]   # > resident warps: several tiles per warp
worst = 0.0
for dtype, B, C, T in cases:
    act = SnakeBeta(C, alpha_logscale=True)
    with torch.no_grad():
        act.alpha.normal_(0, 0.5)
        act.beta.normal_(0, 0.5)
    m = Activation1d(activation=act).to(dev)
    x = torch.randn(B, C, T, device=dev).to(dtype).requires_grad_(True)
    gy = torch.randn(B, C, T, device=dev).to(dtype)
    y = m(x)
    y.backward(gy)
    torch.cuda.synchronize()
    taps = m.upsample.filter
    with torch.no_grad():
        ref = TP.activation1d_torch(x.detach().float(), act.alpha.to(dev), act.beta.to(dev), True, taps, taps)
    err = ((y.detach().float() - ref).abs().max() / ref.abs().max()).item()
    worst = max(worst, err / (1e-5 if dtype == torch.float32 else 1e-2))
    print(dtype, B, C, T, f"E={err:.2e}", flush=True)
assert worst <= 1.0, worst
print("sanitize_case OK")
