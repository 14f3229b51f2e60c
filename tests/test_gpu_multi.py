"""Two-rank NCCL tests (-m gpu; skipped on boxes with fewer than two GPUs): the fused Activation1d under
DistributedDataParallel, as train_binaural_mel.py:540-543 / :787-791 runs it."""
import os
import socket
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(REPO, "diffbinaural-binaural-audio-generation_b200")
for p in (REPO, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_model(C, dev):
    """conv -> Activation1d(SnakeBeta) -> conv -> Activation1d(Snake): filters, both activation kinds, weights before and after."""
    from afa_b200 import Activation1d
    from afa_b200.activations import Snake, SnakeBeta

    torch.manual_seed(77)
    m = torch.nn.Sequential(
        torch.nn.Conv1d(4, C, 3, padding=1),
        Activation1d(activation=SnakeBeta(C, alpha_logscale=True)),
        torch.nn.Conv1d(C, C, 3, padding=1),
        Activation1d(activation=Snake(C, alpha_logscale=True)),
    )
    with torch.no_grad():
        m[1].act.alpha.normal_(0, 0.3)
        m[1].act.beta.normal_(0, 0.3)
        m[3].act.alpha.normal_(0, 0.3)
    return m.to(dev)


def _inputs(world, B, T):
    g = torch.Generator().manual_seed(5)
    return [torch.randn(B, 4, T, generator=g) for _ in range(world)]


def _ddp_worker(rank, world, port, out_dir, C, B, T):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        m = _make_model(C, dev)
        ddp = DDP(m, device_ids=[rank], find_unused_parameters=False)          # train_binaural_mel.py:540-543 (broadcast_buffers default)
        x = _inputs(world, B, T)[rank].to(dev)
        for step in range(2):                                                   # two steps: the buffer re-broadcast of step 2 included
            ddp.zero_grad(set_to_none=True)
            ddp(x).square().mean().backward()
        torch.cuda.synchronize()
        torch.save({n: p.grad.detach().cpu() for n, p in m.named_parameters()}, os.path.join(out_dir, f"g{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_ddp_two_ranks_nccl_gradients_equal_the_single_process_average(tmp_path):
    """DDP's all-reduced gradients on both ranks == the average of the per-shard gradients computed in one process on one GPU
    (same fused kernels, deterministic two-stage parameter-gradient reductions): bit-identical across ranks, and equal to the
    single-process average to fp32 summation-order accuracy -- for convolution weights and for alpha / beta of both
    activation kinds."""
    import torch.multiprocessing as mp

    world, C, B, T = 2, 16, 3, 1000
    mp.spawn(_ddp_worker, args=(world, _free_port(), str(tmp_path), C, B, T), nprocs=world, join=True)
    g = [torch.load(os.path.join(str(tmp_path), f"g{r}.pt")) for r in range(world)]
    for n in g[0]:
        assert torch.equal(g[0][n], g[1][n]), n                                  # one all-reduce result on both ranks

    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    m = _make_model(C, dev)
    acc = {n: torch.zeros_like(p) for n, p in m.named_parameters()}
    for x in _inputs(world, B, T):
        m.zero_grad(set_to_none=True)
        m(x.to(dev)).square().mean().backward()
        for n, p in m.named_parameters():
            acc[n] += p.grad / world
    for n in acc:
        ref = acc[n].cpu()
        err = (g[0][n] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        assert err <= 2e-5, (n, err)
    assert {"1.act.alpha", "1.act.beta", "3.act.alpha"} <= set(acc)
