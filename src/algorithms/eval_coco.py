"""Mirror of the reference's src/algorithms/eval_coco.py (COCOEvaluator.evaluate / evaluate_recall / evaluate_n_fold,
:273-448) on the CUDA rank kernel: rank of the best positive = number of gallery items scoring strictly higher
(equivalent to the reference's sort + search; the 7 identical embedding copies of
eval_coco.py:135,175 scale every similarity by 49 and cannot change a rank, so one copy is used)."""
from __future__ import annotations

import numpy as np
import torch

from creamfl_b200 import ops


def recall_at_k(ranks, k):
    """eval_coco.py:22-29."""
    return 100.0 * len(np.where(ranks < k)[0]) / len(ranks)


class COCOEvaluator:
    def __init__(self, eval_method='matmul', verbose=False, eval_device='cuda', n_crossfolds=5):
        if eval_method not in ('matmul', 'matching_prob'):
            raise ValueError(f'unknown eval_method {eval_method!r} (matmul | matching_prob, eval_coco.py:78)')
        self.eval_method, self.verbose, self.eval_device, self.n_crossfolds = eval_method, verbose, eval_device, n_crossfolds
        self.model = self.criterion = self.logger = None

    def set_model(self, model):
        self.model = model

    def set_criterion(self, criterion):
        self.criterion = criterion

    def set_logger(self, logger):
        self.logger = logger

    @torch.no_grad()
    def extract_features(self, dataloader):
        """eval_coco.py:118-223: image features de-duplicated by image id, one caption feature per annotation."""
        self.model.eval()
        dev = next(self.model.parameters()).device
        img_feats, img_ids, cap_feats, cap_labels, seen = [], [], [], [], set()
        for images, targets, captions, lens, ann_ids, image_ids, _ in dataloader:
            out = self.model(images.to(dev, non_blocking=True), targets.to(dev), captions, lens)
            fi, ft = out['image_features'], out['caption_features']
            for j, iid in enumerate(image_ids):
                if iid not in seen:
                    seen.add(iid)
                    img_feats.append(fi[j])
                    img_ids.append(iid)
            cap_feats.append(ft)
            cap_labels.extend(image_ids)
        return torch.stack(img_feats), torch.cat(cap_feats), torch.tensor(img_ids), torch.tensor(cap_labels)

    @torch.no_grad()
    def evaluate_recall(self, q_features, g_features, q_labels, g_labels):
        """eval_coco.py:273-334."""
        if len(q_features) != len(q_labels):
            raise RuntimeError('length mismatch {}, {}'.format(q_features.shape, q_labels.shape))
        if len(g_features) != len(g_labels):
            raise RuntimeError('length mismatch {}, {}'.format(g_features.shape, g_labels.shape))
        dev = q_features.device
        if self.eval_method == 'matching_prob':
            ranks = self._ranks_by_matching_prob(q_features, g_features, q_labels.to(dev), g_labels.to(dev))
        else:
            ranks = ops.recall_ranks(q_features.float(), g_features.float(), q_labels.to(dev), g_labels.to(dev))
        best = ranks.cpu().numpy().astype(np.float64)
        scores = {'recall_1': recall_at_k(best, 1), 'recall_5': recall_at_k(best, 5), 'recall_10': recall_at_k(best, 10)}
        scores['rsum'] = scores['recall_1'] + scores['recall_5'] + scores['recall_10']
        scores['medr'] = float(np.floor(np.median(best)) + 1)
        scores['meanr'] = float(np.mean(best) + 1)
        return scores

    def _ranks_by_matching_prob(self, q, g, q_labels, g_labels, chunk=64):
        """eval_method 'matching_prob' (MatchingProbModule, eval_coco.py:54-72; neither yaml uses it): similarities are
        criterion.match_prob(query, gallery); rank of the best positive = gallery items with a strictly higher
        probability, the same rule as the rank kernel.  Plain device tensor ops, `chunk` queries at a time."""
        if self.criterion is None:
            raise RuntimeError("eval_method 'matching_prob' needs set_criterion()")
        ranks = torch.empty(len(q), dtype=torch.int32, device=q.device)
        gk = g if g.dim() == 3 else g.unsqueeze(1)
        for s in range(0, len(q), chunk):
            qs = q[s:s + chunk]
            qk = qs if qs.dim() == 3 else qs.unsqueeze(1)
            sims = torch.stack([self.criterion.match_prob(one.unsqueeze(0), gk, None, None) for one in qk])
            pos = q_labels[s:s + chunk, None] == g_labels[None, :]
            best = sims.masked_fill(~pos, float('-inf')).max(dim=1, keepdim=True).values
            r = (sims > best).sum(dim=1).to(torch.int32)
            ranks[s:s + chunk] = torch.where(pos.any(dim=1), r, torch.full_like(r, len(g)))
        return ranks

    def evaluate_n_fold(self, extracted, n_crossfolds, n_images_per_crossfold, n_captions_per_crossfold):
        """eval_coco.py:336-390: COCO-1K protocol - average over folds of 1000 images / 5000 captions."""
        img, cap, il, cl = extracted
        acc = {'i2t': {}, 't2i': {}}
        for f in range(n_crossfolds):
            i0, c0 = f * n_images_per_crossfold, f * n_captions_per_crossfold
            fi, fl = img[i0:i0 + n_images_per_crossfold], il[i0:i0 + n_images_per_crossfold]
            fc, fcl = cap[c0:c0 + n_captions_per_crossfold], cl[c0:c0 + n_captions_per_crossfold]
            for name, s in (('i2t', self.evaluate_recall(fi, fc, fl, fcl)), ('t2i', self.evaluate_recall(fc, fi, fcl, fl))):
                for k, v in s.items():
                    acc[name].setdefault(k, []).append(v)
        out = {d: {k: float(np.mean(v)) for k, v in s.items()} for d, s in acc.items()}
        out['rsum'] = out['i2t']['rsum'] + out['t2i']['rsum']
        return out

    @torch.no_grad()
    def evaluate(self, dataloader, n_crossfolds=None, n_images_per_crossfold=1000, n_captions_per_crossfold=5000,
                 eval_batch_size=1024, key=None):
        """eval_coco.py:392-448."""
        extracted = self.extract_features(dataloader)
        img, cap, il, cl = extracted
        scores = {}
        if n_crossfolds and n_crossfolds > 0:
            scores['n_fold'] = self.evaluate_n_fold(extracted, n_crossfolds, n_images_per_crossfold,
                                                    n_captions_per_crossfold)
        scores['i2t'] = self.evaluate_recall(img, cap, il, cl)
        scores['t2i'] = self.evaluate_recall(cap, img, cl, il)
        scores['rsum'] = scores['i2t']['rsum'] + scores['t2i']['rsum']
        return scores
