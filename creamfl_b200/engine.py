"""Step drivers of the CreamFL hot path: what the reference's trainers do per batch, on the creamfl_b200 kernels.

  ServerEngine.train_step    - retrieval_trainer.TrainerEngine.train body (retrieval_trainer.py:192-214)
  ServerEngine.extract       - global public representations (MMFL.py:194-221)
  ServerEngine.distill_step  - MMFL.distill body (MMFL.py:346-391)
  MMClient.private_step      - MMClientTrainer.train_epoch, private pass (MMClientTrainer.py:118-143)
  MMClient.contrast_step     - inter + intra contrast pass (MMClientTrainer.py:154-222)
  MMClient.generate          - MMClientTrainer.generate_logits (MMClientTrainer.py:326-359)
  aggregate                  - MMFL.distill.aggregation, `con_w` (MMFL.py:298-335)
  exchange_*                 - the one real exchange of the path when clients are sharded one per GPU: all-gather of
                               the public-set representations into the server ensemble

Everything numerical happens in libcreamfl_b200.so; this file sequences calls and owns buffers.  Apex O2 loss
scaling has no counterpart (bf16 needs none).
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops
from .clients import ClientPCME
from .criterions import get_criterion
from .optim import FusedOptimizer
from .towers import PCME

PCME_CRITERION_CFG = {'init_shift': 15, 'init_negative_scale': 15, 'num_samples': 7}    # coco.yaml:41-47


def default_device(index: Optional[int] = None) -> torch.device:
    """The CUDA device the engines run on (current device unless an index is given).  There is no CPU path: without a
    CUDA device this raises instead of handing back a CPU device."""
    if not torch.cuda.is_available():
        raise RuntimeError('creamfl_b200 engines need a CUDA device (the hot path has no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device() if index is None else index)


def _features(output: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    return output['image_features'], output['caption_features']


class GraphedStep:
    """One step function captured as a CUDA graph (the client steps issue ~1500 short kernels and are otherwise
    bound by launch overhead).  Inputs are copied into static buffers, the graph is replayed, outputs are static."""

    def __init__(self, fn, example: Dict[str, torch.Tensor], warmup: int = 3):
        self.static = {k: v.clone() for k, v in example.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(**self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launches()
        with torch.cuda.graph(self.graph, capture_error_mode='thread_local'):
            self.out = fn(**self.static)
        self.launches = ops.launches() - l0          # kernels of ours inside one replay

    def __call__(self, **inputs):
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        ops._launches += self.launches
        out = self.out
        if torch.is_tensor(out) and out.numel() <= 16:
            return out.clone()                       # scalars (losses) must survive the next replay
        return out


class ServerEngine:
    def __init__(self, embed_dim: int = 256, cnn_type: str = 'resnet101', lr: float = 2e-4, grad_clip: float = 2.0,
                 kd_weight: float = 0.3, device: Optional[torch.device] = None, data_parallel: bool = False,
                 use_graphs: bool = False):
        self.device = device or default_device()
        self._graphs = {}
        self.model = PCME(None, {'embed_dim': embed_dim, 'cnn_type': cnn_type, 'not_bert': False}).to(self.device)
        self.criterion = get_criterion('pcme', PCME_CRITERION_CFG).to(self.device)
        self.model.store()
        params = [p for p in self.model.parameters() if p.requires_grad] + list(self.criterion.parameters())
        self.optimizer = FusedOptimizer(params, lr=lr, max_norm=grad_clip, mode='adamp',
                                        no_clip=list(self.criterion.parameters())).attach_stores(self.model)
        self.kd_weight = kd_weight
        self.data_parallel = data_parallel and dist.is_initialized() and dist.get_world_size() > 1
        # data-parallel: forward+backward is one captured graph, the NCCL all-reduce of the flat gradient buffer and the
        # 4-launch optimizer step run eagerly after it (capturing the collective inside the graph deadlocked on 2 GPUs)
        self.use_graphs = use_graphs

    def _graphed(self, name, fn, **tensors):
        key = (name, tuple((k, tuple(t.shape)) for k, t in tensors.items()))
        g = self._graphs.get(key)
        if g is None:
            self.optimizer.prepare()
            g = self._graphs[key] = GraphedStep(fn, tensors)
        return g(**tensors)

    def _sync_grads(self) -> None:
        """Replicated server, batches sharded over ranks: average the flat gradient buffer (one NCCL call)."""
        if self.data_parallel:
            average_gradients(self.model.store().grad)
            for p in self.criterion.parameters():
                average_gradients(p.grad)

    def train_step(self, images, tokens) -> torch.Tensor:
        if self.use_graphs and isinstance(tokens, dict):
            fn = self._train_fwd_bwd if self.data_parallel else self._train_step
            loss = self._graphed('train', lambda images, ids, mask: fn(
                images, {'input_ids': ids, 'attention_mask': mask}), images=images, ids=tokens['input_ids'],
                mask=tokens['attention_mask'])
            if self.data_parallel:
                self._sync_grads()
                self.optimizer.step()
            return loss
        return self._train_step(images, tokens)

    def _train_fwd_bwd(self, images, tokens) -> torch.Tensor:
        self.model.train()
        output = self.model(images, None, tokens, None)
        loss, _ = self.criterion(**output)
        self.optimizer.zero_grad()
        loss.backward()
        return loss.detach()

    def _train_step(self, images, tokens) -> torch.Tensor:
        loss = self._train_fwd_bwd(images, tokens)
        self._sync_grads()
        self.optimizer.step()
        return loss

    def extract(self, images, tokens) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.use_graphs and isinstance(tokens, dict):
            return self._graphed('extract', lambda images, ids, mask: self._extract(
                images, {'input_ids': ids, 'attention_mask': mask}), images=images, ids=tokens['input_ids'],
                mask=tokens['attention_mask'])
        return self._extract(images, tokens)

    @torch.no_grad()
    def _extract(self, images, tokens) -> Tuple[torch.Tensor, torch.Tensor]:
        self.model.eval()
        return _features(self.model(images, None, tokens, None))

    def distill_step(self, images, tokens, d_idx, agg_img, agg_txt, img_terms: int = 1, txt_terms: int = 1):
        """`img_terms` / `txt_terms`: how many times the reference adds the image / text MSE - once per client type
        that carries the modality (MMFL.py:361-378; 2 each when image, text and multimodal clients all exist)."""
        if self.use_graphs and isinstance(tokens, dict) and agg_img is not None and agg_txt is not None:
            # aggregated targets are passed as graph inputs (they are re-created every round)
            fn = self._distill_fwd_bwd if self.data_parallel else self._distill_step
            loss = self._graphed(('distill', img_terms, txt_terms),
                                 lambda images, ids, mask, d_idx, agg_img, agg_txt: fn(
                                     images, {'input_ids': ids, 'attention_mask': mask}, d_idx, agg_img, agg_txt,
                                     img_terms, txt_terms),
                                 images=images, ids=tokens['input_ids'], mask=tokens['attention_mask'], d_idx=d_idx,
                                 agg_img=agg_img, agg_txt=agg_txt)
            if self.data_parallel:
                self._sync_grads()
                self.optimizer.step()
            return loss
        return self._distill_step(images, tokens, d_idx, agg_img, agg_txt, img_terms, txt_terms)

    def _distill_step(self, images, tokens, d_idx, agg_img, agg_txt, img_terms: int = 1, txt_terms: int = 1):
        loss = self._distill_fwd_bwd(images, tokens, d_idx, agg_img, agg_txt, img_terms, txt_terms)
        self._sync_grads()
        self.optimizer.step()
        return loss

    def _distill_fwd_bwd(self, images, tokens, d_idx, agg_img, agg_txt, img_terms: int = 1, txt_terms: int = 1):
        self.model.train()
        out_img, out_txt = _features(self.model(images, None, tokens, None))
        loss = 0
        if agg_img is not None and img_terms:
            loss = loss + (self.kd_weight * img_terms) * ops.mse_gather_loss(out_img, agg_img, d_idx)
        if agg_txt is not None and txt_terms:
            loss = loss + (self.kd_weight * txt_terms) * ops.mse_gather_loss(out_txt, agg_txt, d_idx)
        self.optimizer.zero_grad()
        loss.backward()
        return loss.detach()


class MMClient:
    def __init__(self, embed_dim: int = 256, lr: float = 2e-4, grad_clip: float = 2.0, interintra_weight: float = 0.5,
                 vocab_size: int = 11755, device: Optional[torch.device] = None, use_graphs: bool = False):
        self.device = device or default_device()
        self.use_graphs = use_graphs
        self._graphs = {}
        self.model = ClientPCME(vocab_size, embed_dim).to(self.device)
        self.criterion = get_criterion('pcme', PCME_CRITERION_CFG).to(self.device)
        self.model.store()
        params = [p for p in self.model.parameters() if p.requires_grad] + list(self.criterion.parameters())
        self.optimizer = FusedOptimizer(params, lr=lr, max_norm=grad_clip, mode='adamp',
                                        no_clip=list(self.criterion.parameters())).attach_stores(self.model)
        self.w = interintra_weight
        self.old_model = None

    def begin_round(self) -> None:
        """old_model = deepcopy(model) (MMClientTrainer.py:92-93); after the first round the copy is refreshed in
        place so that its addresses - and any captured graph - stay valid."""
        if self.old_model is None:
            self.old_model = copy.deepcopy(self.model).eval()
            self.old_model.store()
        else:
            self.old_model.copy_weights_from(self.model)
        self.model.train()

    def _graphed(self, name, lengths, fn, **tensors):
        key = (name, tuple(int(v) for v in lengths), tuple((k, tuple(t.shape)) for k, t in tensors.items()))
        g = self._graphs.get(key)
        if g is None:
            self.optimizer.prepare()
            g = self._graphs[key] = GraphedStep(fn, tensors)
        return g(**tensors)

    def private_step(self, images, captions, lengths) -> torch.Tensor:
        if self.use_graphs:
            return self._graphed('private', lengths, lambda images, captions: self._private_step(images, captions, lengths),
                                 images=images, captions=captions)
        return self._private_step(images, captions, lengths)

    def _private_step(self, images, captions, lengths) -> torch.Tensor:
        self.model.train()
        output = self.model(images, captions, None, lengths)
        loss, _ = self.criterion(**output)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def contrast_step(self, images, captions, lengths, d_idx, g_img, g_txt, g_img16, g_txt16, intra: bool = True,
                      inter: bool = True, loss_scale: bool = False) -> torch.Tensor:
        """g_img / g_txt: fp32 [N_pub, D] server features; g_*16 their bf16 copies (made once per round)."""
        if self.use_graphs:
            # the banks are persistent buffers refreshed in place by the caller: captured by address
            key_banks = (g_img.data_ptr(), g_txt.data_ptr(), g_img16.data_ptr(), g_txt16.data_ptr(), intra, inter,
                         loss_scale)
            return self._graphed(('contrast',) + key_banks, lengths,
                                 lambda images, captions, d_idx: self._contrast_step(
                                     images, captions, lengths, d_idx, g_img, g_txt, g_img16, g_txt16, intra, inter,
                                     loss_scale),
                                 images=images, captions=captions, d_idx=d_idx)
        return self._contrast_step(images, captions, lengths, d_idx, g_img, g_txt, g_img16, g_txt16, intra, inter,
                                   loss_scale)

    def _contrast_step(self, images, captions, lengths, d_idx, g_img, g_txt, g_img16, g_txt16, intra: bool = True,
                       inter: bool = True, loss_scale: bool = False) -> torch.Tensor:
        self.model.train()
        self.optimizer.zero_grad()
        out_img, out_txt = _features(self.model(images, captions, None, lengths))
        b = images.shape[0]
        loss_intra = loss_inter = None
        if intra:
            with torch.no_grad():
                old_img, old_txt = _features(self.old_model(images, captions, None, lengths))
            loss_intra = ops.moon_intra_loss(out_img, old_img, g_img, d_idx, 2.0, 2 * b) + \
                ops.moon_intra_loss(out_txt, old_txt, g_txt, d_idx, 2.0, 2 * b)
        if inter:
            loss_inter = ops.infonce_loss(out_img, g_txt16, d_idx, 2.0) + ops.infonce_loss(out_txt, g_img16, d_idx, 2.0)
        if intra and inter:
            if not loss_scale:
                loss = (loss_intra + loss_inter) * self.w
            else:
                loss = (loss_intra + loss_inter / (loss_inter / loss_intra).detach()) * self.w
        else:
            loss = loss_intra if intra else loss_inter                # MMClientTrainer.py:264,308: no weight
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def generate(self, images, captions, lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        if self.use_graphs:
            return self._graphed('generate', lengths, lambda images, captions: self._generate(images, captions, lengths),
                                 images=images, captions=captions)
        return self._generate(images, captions, lengths)

    @torch.no_grad()
    def _generate(self, images, captions, lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        self.model.eval()
        return _features(self.model(images, captions, None, lengths))


def aggregate(vecs: Sequence[torch.Tensor], global_other: torch.Tensor) -> torch.Tensor:
    """con_w aggregation of one modality on one device (MMFL.py:298-314)."""
    return ops.conw_aggregate(list(vecs), global_other)


def gather_client_rows(score: torch.Tensor, vec: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The exchange step: every rank contributes its client's contrastive scores [N_pub] and representations
    [N_pub, D]; every rank receives them stacked in rank order, [C, N_pub] and [C, N_pub, D] (C = world size).
    Pure torch.distributed (NCCL on GPUs, gloo in the CPU tests); with one process it just adds the client axis."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return score.unsqueeze(0), vec.unsqueeze(0)
    world = dist.get_world_size()
    scores = torch.empty((world,) + tuple(score.shape), dtype=score.dtype, device=score.device)
    vecs = torch.empty((world,) + tuple(vec.shape), dtype=vec.dtype, device=vec.device)
    dist.all_gather_into_tensor(scores.view(-1), score.contiguous().view(-1))
    dist.all_gather_into_tensor(vecs.view(-1), vec.contiguous().view(-1))
    return scores, vecs


def average_gradients(flat_grad: torch.Tensor) -> None:
    """Replicated server with the public batches sharded over ranks: mean of the flat gradient buffer."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        if flat_grad.is_cuda and dist.get_backend() == 'nccl':
            dist.all_reduce(flat_grad, op=dist.ReduceOp.AVG)         # one pass: NCCL divides inside the collective
        else:                                                         # gloo (CPU tests) has no AVG
            dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
            flat_grad.mul_(1.0 / dist.get_world_size())


def exchange_and_aggregate(own_vec: torch.Tensor, global_other16: torch.Tensor) -> torch.Tensor:
    """Clients sharded one per rank: every rank scores its own client's [N_pub, D] representations against the
    replicated server features (tcgen05, 1.28 TFLOP each), then the ranks all-gather scores (200 KB) and
    representations (51 MB) and every rank forms the softmax-over-clients weighted sum - the single exchange step of
    the path.  With one rank this is `aggregate([own_vec], ...)`."""
    score = ops.conw_score(ops.to_bf16(own_vec), global_other16)
    scores, vecs = gather_client_rows(score, own_vec)
    return ops.conw_reduce([vecs[r] for r in range(vecs.shape[0])], scores)
