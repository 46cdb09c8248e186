"""Mirror of the reference's criterion factory (src/criterions/__init__.py:4-8) on the CUDA loss kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class LossDict(dict):
    """The reference returns 11 `.item()` floats per step (probemb.py:244-255) - 11 host synchronisations.  This
    dict materialises them from one device tensor on first access, so a step that never looks pays nothing."""

    def __init__(self, parts: torch.Tensor, shift: torch.Tensor, scale: torch.Tensor):
        super().__init__()
        self._src = (parts.detach(), shift.detach(), scale.detach())

    def _fill(self):
        if self._src is None:
            return
        parts, shift, scale = self._src
        self._src = None
        loss, pos, neg = (float(x) for x in parts.cpu())
        d = {'i2t_loss': pos + neg, 't2i_loss': pos + neg, 'i2t_pos_loss': pos, 'i2t_neg_loss': neg,
             't2i_pos_loss': pos, 't2i_neg_loss': neg, 'uniform_loss': 0, 'vib_loss': 0, 'shift': float(shift),
             'negative_scale': float(scale), 'loss': loss}
        super().update(d)

    def __getitem__(self, k):
        self._fill()
        return super().__getitem__(k)

    def __iter__(self):
        self._fill()
        return super().__iter__()

    def __len__(self):
        self._fill()
        return super().__len__()

    def keys(self):
        self._fill()
        return super().keys()

    def items(self):
        self._fill()
        return super().items()

    def values(self):
        self._fill()
        return super().values()

    def __repr__(self):
        self._fill()
        return super().__repr__()


class MCSoftContrastiveLoss(nn.Module):
    """src/criterions/probemb.py:89-256 for the configuration the reference runs (one embedding per item,
    uniform_lambda = vib_beta = 0, reduction 'sum'): loss = i2t + t2i soft-contrastive NLL over all N^2 pairs with
    learnable `shift` / `negative_scale`."""

    def __init__(self, config, reduction='sum'):
        super().__init__()
        if reduction not in {'mean', 'sum', None}:
            raise ValueError('unknown reduction {}'.format(reduction))
        get = config.get if hasattr(config, 'get') else (lambda k, d=None: getattr(config, k, d))
        if get('uniform_lambda', 0) != 0 or get('vib_beta', 0) != 0:
            raise NotImplementedError('uniform / VIB terms are disabled in the reference configuration (coco.yaml)')
        if reduction != 'sum':
            raise NotImplementedError("the reference instantiates the criterion with reduction='sum' only")
        self.reduction = reduction
        dev = 'cuda:0' if torch.cuda.is_available() else 'cpu'
        self.shift = nn.Parameter(float(get('init_shift')) * torch.ones(1, device=dev))
        self.negative_scale = nn.Parameter(float(get('init_negative_scale')) * torch.ones(1, device=dev))
        self.num_samples = get('num_samples', 1)

    def match_prob(self, image_features, caption_features, image_logsigma=None, caption_logsigma=None, **kw):
        """probemb.py:210-219 (+ batchwise_cdist :7-45): matching probability sigma(-a d + b) with the reference's
        sigma(x) = e^x / (e^x + e^-x), averaged over the K x K embedding pairs of each row.  Inputs [N, K, D] (2-D
        inputs mean K = 1); the row counts must be equal or one of them 1 (a query broadcast against the gallery, the
        way MatchingProbModule calls it, eval_coco.py:66-69).  Evaluation-only and not on the hot path (both yaml
        configurations evaluate with 'matmul'): a handful of device tensor ops, no kernel of its own.  Written as
        sigmoid(2x), which equals the reference's quotient wherever that does not overflow to inf / inf."""
        a, b = image_features, caption_features
        if a.dim() != 3 or b.dim() != 3:
            a, b = a.unsqueeze(1), b.unsqueeze(1)
        na, nb = a.size(0), b.size(0)
        if not (na == nb or na == 1 or nb == 1):
            raise RuntimeError(f'samples1 ({a.size()}) and samples2 ({b.size()}) dimensionalities '
                               'are non-broadcastable.')
        dist = torch.sqrt(((a.unsqueeze(1) - b.unsqueeze(2)) ** 2).sum(-1) + 1e-6).reshape(max(na, nb), -1)
        logits = -self.negative_scale * dist.to(self.negative_scale.device).float() + self.shift
        return torch.sigmoid(2.0 * logits).mean(dim=1)

    def forward(self, image_features, caption_features, image_logsigma=None, caption_logsigma=None, **kwargs):
        loss, parts = ops.pcme_loss(image_features, caption_features, self.shift, self.negative_scale)
        return loss, LossDict(parts, self.shift, self.negative_scale)


def get_criterion(criterion_name, config):
    if criterion_name == 'pcme':
        return MCSoftContrastiveLoss(config)
    raise ValueError(f'Invalid criterion name: {criterion_name}')
