"""CPU oracle for the CreamFL hot path - TEST INFRASTRUCTURE ONLY.

This module restates, in plain torch-CPU / numpy, the arithmetic of the reference (FLAIR-THU/CreamFL) for every
operation that creamfl_b200 implements in CUDA.  It exists so that tests can check the CUDA path; nothing in the
product (creamfl_b200/, src/) may import it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs use it.

Parity pinning: the reference ships no tests; its only golden artefacts are index fixtures, of which
data_partition/client_noniid_flicker30k.pkl and coco_subset_idx_file are reproducible without a dataset -
shard_partition() / public_subset_indices() below and the product's reproduce them bit for bit
(tests/test_cpu_partition.py).  Everything else is pinned
against outputs of the reference's own modules executed in the build container: tests/golden/make_golden.py imports
/root/reference (with import shims for packages absent from the image) and writes tests/golden/*.npz;
tests/test_oracle_golden.py checks every function below against those files.  Functions whose reference code is
inline in a trainer method are pinned by driving that very method (MMClientTrainer.train_epoch,
ClientTrainer.tra, MMFL.distill) with stub models/loaders and recording loss and gradients.

Every function cites the reference lines it follows.  Default precision is float64 so that tolerances in the CUDA
parity tests measure the CUDA path, not the oracle.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------- helpers
def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    """reference src/utils/tensor_utils.py:25-27 (F.normalize, eps 1e-12)."""
    return x / x.norm(p=2, dim=-1, keepdim=True).clamp_min(1e-12)


# ------------------------------------------------------------------------------------------------- PCME loss
def pcme_pair_distance(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """reference src/criterions/probemb.py:7-45 with K = 1 embedding per item: d[i,j] = sqrt(|a_i-b_j|^2+eps)."""
    diff = a[:, None, :] - b[None, :, :]
    return torch.sqrt((diff ** 2).sum(-1) + eps)


def pcme_direction_loss(anchors, candidates, shift, negative_scale) -> Dict[str, torch.Tensor]:
    """reference src/criterions/probemb.py:150-208 (full_sampling + soft_contrastive_nll with one sample).

    For a single sample per item, soft_contrastive_nll(l, m) = -(l*m - logaddexp(l, -l)) + log(1).
    """
    if len(anchors) != len(candidates):
        raise RuntimeError('# anchors ({}) != # candidates ({})'.format(anchors.shape, candidates.shape))
    n = len(anchors)
    dist = pcme_pair_distance(anchors, candidates)
    logits = -negative_scale * dist + shift
    matched = 2.0 * torch.eye(n, dtype=logits.dtype, device=logits.device) - 1.0
    nll = -(logits * matched - torch.logaddexp(logits, -logits))
    eye = torch.eye(n, dtype=torch.bool, device=logits.device)
    pos = nll[eye].sum()
    neg = nll[~eye].sum()
    return {'loss': pos + neg, 'pos_loss': pos, 'neg_loss': neg}


def pcme_loss(img, txt, shift, negative_scale) -> Tuple[torch.Tensor, Dict[str, float]]:
    """reference src/criterions/probemb.py:221-256 with uniform_lambda = vib_beta = 0 (coco.yaml)."""
    i2t = pcme_direction_loss(img, txt, shift, negative_scale)
    t2i = pcme_direction_loss(txt, img, shift, negative_scale)
    loss = i2t['loss'] + t2i['loss']
    info = {
        'i2t_loss': float(i2t['loss']), 't2i_loss': float(t2i['loss']),
        'i2t_pos_loss': float(i2t['pos_loss']), 'i2t_neg_loss': float(i2t['neg_loss']),
        't2i_pos_loss': float(t2i['pos_loss']), 't2i_neg_loss': float(t2i['neg_loss']),
        'uniform_loss': 0, 'vib_loss': 0, 'shift': float(shift), 'negative_scale': float(negative_scale),
        'loss': float(loss),
    }
    return loss, info


def pcme_match_prob(a: torch.Tensor, b: torch.Tensor, shift, negative_scale, eps: float = 1e-6) -> torch.Tensor:
    """reference src/criterions/probemb.py:210-219 over batchwise_cdist (:7-45): a, b are [N, K, D] (2-D: K = 1) with
    equal row counts or one of them 1; distance [N, K*K] rounded to fp32 as the reference does (:215), probability
    e^l / (e^l + e^-l) of l = -scale * d + shift, mean over the K*K pairs."""
    if a.dim() != 3 or b.dim() != 3:
        a, b = a.unsqueeze(1), b.unsqueeze(1)
    if not (a.size(0) == b.size(0) or a.size(0) == 1 or b.size(0) == 1):
        raise RuntimeError('non-broadcastable')
    n = max(a.size(0), b.size(0))
    dist = torch.sqrt(((a.unsqueeze(1) - b.unsqueeze(2)) ** 2).sum(-1) + eps).view(n, -1).float()
    logits = -negative_scale * dist + shift
    return (torch.exp(logits) / (torch.exp(logits) + torch.exp(-logits))).mean(dim=1)


# ------------------------------------------------------------------------------------------------- contrast
def inter_infonce(q: torch.Tensor, bank: torch.Tensor, labels: torch.Tensor, tau: float = 0.5) -> torch.Tensor:
    """reference MMClientTrainer.py:194-200 / ClientTrainer.py:388,398-401:
    CrossEntropyLoss()(matmul(q, bank.T) / 0.5, labels)."""
    logits = torch.matmul(q, bank.T) / tau
    return torch.nn.functional.cross_entropy(logits, labels)


def moon_intra(z: torch.Tensor, z_old: torch.Tensor, target: torch.Tensor, tau: float = 0.5,
               denom: Optional[int] = None) -> torch.Tensor:
    """reference MMClientTrainer.py:172-191 (one modality block) / ClientTrainer.py:404-414:
    logits = [<z,target>, <z,z_old>] / 0.5, label 0, mean over `denom` rows (2B when both modalities stack)."""
    pos = (z * target).sum(-1, keepdim=True)
    neg = (z * z_old).sum(-1, keepdim=True)
    logits = torch.cat((pos, neg), dim=1) / tau
    labels = torch.zeros(z.shape[0], dtype=torch.long, device=z.device)
    ce = torch.nn.functional.cross_entropy(logits, labels, reduction='sum')
    return ce / (z.shape[0] if denom is None else denom)


def mm_client_contrast_loss(out_img, out_txt, old_img, old_txt, g_img, g_txt, d_idx, interintra_weight=0.5,
                            loss_scale=False) -> Dict[str, torch.Tensor]:
    """reference MMClientTrainer.py:169-206 (both flags): intra over the stacked [2B,2] logits + inter both ways."""
    b = out_img.shape[0]
    idx = torch.as_tensor(d_idx, dtype=torch.long, device=out_img.device)
    intra = moon_intra(out_img, old_img, g_img[idx], denom=2 * b) + moon_intra(out_txt, old_txt, g_txt[idx],
                                                                                denom=2 * b)
    inter = inter_infonce(out_img, g_txt, idx) + inter_infonce(out_txt, g_img, idx)
    if not loss_scale:
        loss = (intra + inter) * interintra_weight
    else:
        loss = (intra + inter / (inter / intra).detach()) * interintra_weight
    return {'loss': loss, 'intra': intra, 'inter': inter}


def unimodal_contrast_loss(feat, old_feat, g_same, g_other, d_idx, interintra_weight=0.5, loss_scale=False):
    """reference ClientTrainer.py:383-419: inter against the opposite modality bank, MOON against own modality."""
    idx = torch.as_tensor(d_idx, dtype=torch.long)
    inter = inter_infonce(feat, g_other, idx)
    moon = moon_intra(feat, old_feat, g_same[idx])
    if not loss_scale:
        loss = (moon + inter) * interintra_weight
    else:
        loss = (moon + inter / (inter / moon).detach()) * interintra_weight
    return {'loss': loss, 'intra': moon, 'inter': inter}


# ------------------------------------------------------------------------------------------------- con_w
def conw_scores(vec: torch.Tensor, global_other: torch.Tensor, chunk: int = 4096) -> torch.Tensor:
    """reference MMFL.py:304-307: diag(logits - log(sum(exp(logits), dim=1))) with logits = vec @ global.T.
    Row-chunked so that N = 50000 fits in memory; the per-row arithmetic is unchanged (no max subtraction,
    exactly like the reference)."""
    n = vec.shape[0]
    out = torch.empty(n, dtype=vec.dtype)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        logits = torch.matmul(vec[s:e], global_other.T)
        log_den = torch.log(torch.sum(torch.exp(logits), dim=1))
        diag = logits[torch.arange(e - s), torch.arange(s, e)]
        out[s:e] = diag - log_den
    return out


def conw_aggregate(vecs: Sequence[torch.Tensor], global_other: torch.Tensor,
                   chunk: int = 4096) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference MMFL.py:298-314 (image branch; the text branch :317-331 is the same with roles swapped), with the
    hard-coded 50000 replaced by len(vec).  Returns (aggregated [N,D], weights [C,N])."""
    scores = torch.stack([conw_scores(v, global_other, chunk) for v in vecs], dim=0)
    w = torch.softmax(scores, dim=0)
    agg = torch.zeros_like(vecs[0])
    for c, v in enumerate(vecs):
        agg = agg + v * w[c].reshape(-1, 1)
    return agg, w


def distill_mse(out: torch.Tensor, agg: torch.Tensor, d_idx) -> torch.Tensor:
    """reference MMFL.py:296,355-365: nn.MSELoss()(out, agg[d_idx, :])."""
    idx = torch.as_tensor(d_idx, dtype=torch.long)
    return torch.nn.functional.mse_loss(out, agg[idx, :].to(out.dtype))


# ------------------------------------------------------------------------------------------------- Recall@K
def recall_ranks_sorted(q: torch.Tensor, g: torch.Tensor, q_labels, g_labels) -> np.ndarray:
    """reference eval_coco.py:40-50 + 311-317: sort gallery by descending similarity, take the position of every
    positive, keep the best one."""
    q_labels = np.asarray(q_labels)
    g_labels = np.asarray(g_labels)
    sims = q.double().mm(g.double().t())
    _, pred_ranks = (-sims).sort()
    best = np.zeros(len(q_labels))
    for qi in range(len(q_labels)):
        pos = np.where(g_labels == q_labels[qi])[0]
        ranks = [torch.where(pred_ranks[qi] == p)[0][0].item() for p in pos]
        best[qi] = min(ranks)
    return best


def recall_ranks_count(q: torch.Tensor, g: torch.Tensor, q_labels, g_labels) -> np.ndarray:
    """Count form of the same rank: #{j : sim(q,j) > max_pos sim(q,pos)} (equal to the sorted form when no
    similarity ties with the best positive exist)."""
    q_labels = torch.as_tensor(np.asarray(q_labels))
    g_labels = torch.as_tensor(np.asarray(g_labels))
    sims = q.double().mm(g.double().t())
    pos_mask = q_labels[:, None] == g_labels[None, :]
    best_pos = torch.where(pos_mask, sims, torch.full_like(sims, -float('inf'))).max(dim=1).values
    return (sims > best_pos[:, None]).sum(dim=1).numpy().astype(np.float64)


def recall_scores(best_ranks: np.ndarray) -> Dict[str, float]:
    """reference eval_coco.py:22-29,319-332."""
    def at(k):
        return 100.0 * len(np.where(best_ranks < k)[0]) / len(best_ranks)
    r1, r5, r10 = at(1), at(5), at(10)
    return {'recall_1': r1, 'recall_5': r5, 'recall_10': r10, 'rsum': r1 + r5 + r10,
            'medr': float(np.floor(np.median(best_ranks)) + 1), 'meanr': float(np.mean(best_ranks) + 1)}


# ------------------------------------------------------------------------------------------------- partition
def hetero_partition(dataset: str, num_samples: int, num_nets: int, alpha: float, y_train: np.ndarray,
                     seed: Optional[int] = None) -> Dict[int, List[int]]:
    """reference src/datasets/load_FL_datasets.py:96-120 (the 'hetero' Dirichlet branch, without the pickle
    cache).  Uses numpy's legacy global RNG exactly like the reference; pass `seed` to seed it first."""
    if seed is not None:
        np.random.seed(seed)
    min_size = 0
    k_classes = int(max(y_train)) + 1
    threshold = 10 if dataset == "cifar100" else (3000 if dataset == "AG_NEWS" else 500)
    idx_batch: List[List[int]] = []
    while min_size < threshold:
        idx_batch = [[] for _ in range(num_nets)]
        for k in range(k_classes):
            idx_k = np.where(y_train == k)[0]
            np.random.shuffle(idx_k)
            proportions = np.random.dirichlet(np.repeat(alpha, num_nets))
            proportions = np.array(
                [p * (len(idx_j) < num_samples / num_nets) for p, idx_j in zip(proportions, idx_batch)])
            proportions = proportions / proportions.sum()
            proportions = (np.cumsum(proportions) * len(idx_k)).astype(int)[:-1]
            idx_batch = [idx_j + idx.tolist() for idx_j, idx in zip(idx_batch, np.split(idx_k, proportions))]
            min_size = min([len(idx_j) for idx_j in idx_batch])
    out = {}
    for j in range(num_nets):
        np.random.shuffle(idx_batch[j])
        out[j] = idx_batch[j]
    return out


def public_subset_indices(subset_num: int = 50000, n_total: int = 566435, seed: int = 2021) -> List[int]:
    """reference src/utils/load_datasets.py:148-157 (`random.shuffle(full_idx); idx = full_idx[0:50000]; idx.sort()`),
    seeded; seed 2021 gives the coco_subset_idx_file the reference ships."""
    import random
    keep = random.getstate()
    random.seed(seed)
    full_idx = [i for i in range(n_total)]
    random.shuffle(full_idx)
    random.setstate(keep)                      # the oracle leaves the caller's global RNG as it found it
    idx = full_idx[0:subset_num]
    idx.sort()
    return idx


def partition_digest(part: Dict[int, Sequence[int]]) -> str:
    h = hashlib.sha256()
    for j in sorted(part):
        h.update(np.asarray(part[j], dtype=np.int64).tobytes())
    return h.hexdigest()


def shard_partition(n_items: int, num_users: int = 15, num_shards: int = 150,
                    seed: Optional[int] = None) -> Dict[int, np.ndarray]:
    """reference src/datasets/flickr30k.py:79-102 (non_iid without the pickle cache): random shards of
    n_items // num_shards consecutive indices, num_shards // num_users shards per client drawn without
    replacement from numpy's legacy global RNG, leftovers appended to the last client.  Iteration over the drawn
    shard set follows CPython set order, exactly like the reference."""
    if seed is not None:
        np.random.seed(seed)
    num_imgs = int(n_items / num_shards)
    idx_shard = [i for i in range(num_shards)]
    dict_users = {i: np.array([], dtype=int) for i in range(num_users)}
    idxs = np.arange(num_shards * num_imgs)
    img_idx = [i for i in range(n_items)]
    i = 0
    for i in range(num_users):
        rand_set = set(np.random.choice(idx_shard, int(num_shards / num_users), replace=False))
        idx_shard = list(set(idx_shard) - rand_set)
        for rand in rand_set:
            dict_users[i] = np.concatenate((dict_users[i], idxs[rand * num_imgs:(rand + 1) * num_imgs]), axis=0)
            img_idx = list(set(img_idx) - set(idxs[rand * num_imgs:(rand + 1) * num_imgs]))
    dict_users[i] = np.concatenate([dict_users[i], img_idx])
    return dict_users


# ------------------------------------------------------------------------------------------------- optimizer
def clip_grad_norm(grads: Sequence[torch.Tensor], max_norm: float) -> float:
    """torch.nn.utils.clip_grad_norm_ (retrieval_trainer.py:211-214): in-place, returns the total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
    coef = min(1.0, max_norm / (float(total) + 1e-6))
    for g in grads:
        g.mul_(coef)
    return float(total)


def adamp_step(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor], exp_avgs: Sequence[torch.Tensor],
               exp_avg_sqs: Sequence[torch.Tensor], step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
               weight_decay: float = 0.0, delta: float = 0.1, wd_ratio: float = 0.1) -> List[int]:
    """AdamP.step restated from the published algorithm (Heo et al., ICLR 2021, Algorithm 1/2 and the adamp==0.3.0
    package the reference pins in requirements.txt:1 and calls at src/algorithms/optimizers.py:24-28; nesterov
    False).  PARITY UNPINNED: the package source is not in the reference tree nor in this image.
    `step` is the 1-based step count.  Returns, per tensor, which projection fired (0 none, 1 channel, 2 layer)."""
    import math
    beta1, beta2 = betas
    fired = []
    for p, grad, exp_avg, exp_avg_sq in zip(params, grads, exp_avgs, exp_avg_sqs):
        bias_correction1 = 1 - beta1 ** step
        bias_correction2 = 1 - beta2 ** step
        exp_avg.mul_(beta1).add_(grad, alpha=1 - beta1)
        exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
        denom = (exp_avg_sq.sqrt() / math.sqrt(bias_correction2)).add_(eps)
        step_size = lr / bias_correction1
        perturb = exp_avg / denom
        ratio, which = 1.0, 0
        if p.dim() > 1:
            views = [lambda x: x.reshape(x.shape[0], -1), lambda x: x.reshape(1, -1)]
            expand = [-1] + [1] * (p.dim() - 1)
            for vi, view in enumerate(views):
                gv, pv = view(grad), view(p)
                cosine = (gv * pv).sum(1).abs() / (gv.norm(dim=1) + eps) / (pv.norm(dim=1) + eps)
                if cosine.max() < delta / math.sqrt(pv.shape[1]):
                    p_n = p / (pv.norm(dim=1).reshape(expand) + eps)
                    perturb = perturb - p_n * view(p_n * perturb).sum(1).reshape(expand)
                    ratio, which = wd_ratio, vi + 1
                    break
        if weight_decay > 0:
            p.mul_(1 - lr * weight_decay * ratio)
        p.add_(perturb, alpha=-step_size)
        fired.append(which)
    return fired


class AdamPRestated(torch.optim.Optimizer):
    """`adamp.AdamP` as the reference constructs it (src/algorithms/optimizers.py:24-28: lr, betas (0.9, 0.999),
    eps 1e-8, weight_decay 0, delta 0.1, wd_ratio 0.1, nesterov False) - a torch.optim.Optimizer around `adamp_step`,
    in the parameters' own dtype and device, one tensor at a time like the package.  Used by bench.py's baseline legs
    (host CPU and eager-torch-on-GPU) so that they pay the reference optimizer's cost; PARITY UNPINNED like
    `adamp_step` itself."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, delta=0.1, wd_ratio=0.1):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, delta=delta,
                                      wd_ratio=wd_ratio))

    @torch.no_grad()
    def step(self, closure=None):
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st['step'], st['exp_avg'], st['exp_avg_sq'] = 0, torch.zeros_like(p), torch.zeros_like(p)
                st['step'] += 1
                adamp_step([p], [p.grad], [st['exp_avg']], [st['exp_avg_sq']], st['step'], group['lr'], group['betas'],
                           group['eps'], group['weight_decay'], group['delta'], group['wd_ratio'])


def sgd_momentum_step(params, grads, bufs, step: int, lr: float, momentum: float = 0.9, weight_decay: float = 0.0):
    """torch.optim.SGD semantics (ClientTrainer.py:287-288): g += wd*p; buf = g (first step) or mom*buf + g; p -= lr*buf."""
    for p, g, buf in zip(params, grads, bufs):
        d = g + weight_decay * p
        if step == 1:
            buf.copy_(d)
        else:
            buf.mul_(momentum).add_(d)
        p.add_(buf, alpha=-lr)
