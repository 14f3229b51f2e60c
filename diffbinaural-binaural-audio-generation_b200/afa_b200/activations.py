"""Parameter holders for Snake / SnakeBeta (reference: BigVGAN/activations.py:9-126).

A user of the reference passes the reference's own `activations.Snake/SnakeBeta` instances into
Activation1d (bigvgan.py:108-124); the fused kernel only reads their `alpha`, `beta`,
`alpha_logscale`.  These stand-ins carry the same constructor, parameter names and initial values so
that harnesses on a box without the reference tree (tests, bench.py) build identical state dicts.
The periodic activation itself is evaluated inside the fused kernel; there is no eager path here.
"""
from __future__ import annotations

import torch
from torch import nn
from torch.nn import Parameter


class _SnakeBase(nn.Module):
    _has_beta = False

    def __init__(self, in_features, alpha=1.0, alpha_trainable=True, alpha_logscale=False):
        super().__init__()
        self.in_features = in_features
        self.alpha_logscale = alpha_logscale
        init = torch.zeros if alpha_logscale else torch.ones                   # activations.py:42-45, 101-106
        self.alpha = Parameter(init(in_features) * alpha)
        self.alpha.requires_grad = alpha_trainable
        if self._has_beta:
            self.beta = Parameter(init(in_features) * alpha)
            self.beta.requires_grad = alpha_trainable
        self.no_div_by_zero = 0.000000001

    def forward(self, x):
        raise RuntimeError(
            f"{type(self).__name__} here is a parameter holder for the fused Activation1d kernel; "
            "wrap it in afa_b200.Activation1d (no eager path is shipped)."
        )


class Snake(_SnakeBase):
    """x + sin^2(alpha x) / alpha"""


class SnakeBeta(_SnakeBase):
    """x + sin^2(alpha x) / beta"""
    _has_beta = True
