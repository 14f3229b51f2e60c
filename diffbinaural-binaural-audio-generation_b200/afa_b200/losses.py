"""Sync-free drop-ins for the GAN loss helpers of the reference's training step.

Reference: BigVGAN/loss.py:213-257, called at BigVGAN/train_binaural_mel.py:726-776.  `discriminator_loss` there calls
`.item()` on two scalars per discriminator output (loss.py:239-240): with MPD (5 periods) + CQT-D (3 scales) that is 32 host
synchronisations in the middle of every training step, for lists the trainer never reads (train_binaural_mel.py:726, :730 only
bind them).  The functions here compute the same values with the same torch ops and return the per-discriminator terms as
detached 0-dim device tensors -- the `List[torch.Tensor]` the reference's own annotation promises -- so nothing waits for the
GPU unless the caller formats or `float()`s an entry.  No kernels: these are a handful of reductions over tiny tensors.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def feature_loss(fmap_r: List[List[torch.Tensor]], fmap_g: List[List[torch.Tensor]]) -> torch.Tensor:
    """loss.py:213-224: 2 * sum over discriminators and layers of mean |real - generated| feature maps."""
    losses = []
    for dr, dg in zip(fmap_r, fmap_g):
        for rl, gl in zip(dr, dg):
            losses.append(torch.mean(torch.abs(rl - gl)))
    return sum(losses) * 2


def discriminator_loss(disc_real_outputs: List[torch.Tensor], disc_generated_outputs: List[torch.Tensor]
                       ) -> Tuple[torch.Tensor, List[torch.Tensor], List[torch.Tensor]]:
    """loss.py:227-243 without its `.item()` calls: (sum_d mean((1 - D(y))^2) + mean(D(g)^2), [real terms], [generated terms])."""
    losses, r_losses, g_losses = [], [], []
    for dr, dg in zip(disc_real_outputs, disc_generated_outputs):
        r_loss = torch.mean((1 - dr) ** 2)
        g_loss = torch.mean(dg ** 2)
        losses.append(r_loss + g_loss)
        r_losses.append(r_loss.detach())
        g_losses.append(g_loss.detach())
    return sum(losses), r_losses, g_losses


def generator_loss(disc_outputs: List[torch.Tensor]) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """loss.py:246-257: (sum_d mean((1 - D(g))^2), [terms])."""
    gen_losses = [torch.mean((1 - dg) ** 2) for dg in disc_outputs]
    return sum(gen_losses), gen_losses
