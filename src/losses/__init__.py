"""Mirror of the reference's src/losses/__init__.py:11-38 for the one entry the CreamFL path instantiates:
`losses.create('softmax')` (ClientTrainer.py:280,284 - every unimodal client's criterion).  The other seven factory
names are the metric-learning losses of the code base CreamFL was forked from; no CreamFL command line reaches them
(dead code) and they are not built."""
from __future__ import annotations

import torch
import torch.nn as nn

from creamfl_b200 import ops


class CrossEntropyLoss(nn.Module):
    """nn.CrossEntropyLoss() with its default mean reduction on the fused CUDA kernel (creamfl_ce_fwd writes the loss
    and the logit gradient in one pass, creamfl_b200/csrc/loss_ops.cu)."""

    def forward(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        return ops.cross_entropy(logits, labels)


__factory = {'softmax': CrossEntropyLoss}
_not_built = ('triplet', 'histogram', 'gaussian', 'batchall', 'neighbour', 'distance_match', 'neighard')


def names():
    return sorted(__factory.keys())


def create(name, *args, **kwargs):
    """src/losses/__init__.py:27-38."""
    if name in _not_built:
        raise KeyError(f'loss {name!r} is dead code in the reference (never selected by src/main.py) and is not built')
    if name not in __factory:
        raise KeyError("Unknown loss:", name)
    return __factory[name](*args, **kwargs)
