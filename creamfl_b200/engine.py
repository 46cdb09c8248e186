"""Step drivers of the CreamFL hot path: what the reference's trainers do per batch, on the creamfl_b200 kernels.

  ServerEngine.train_step    - retrieval_trainer.TrainerEngine.train body (retrieval_trainer.py:192-214)
  ServerEngine.extract       - global public representations (MMFL.py:194-221)
  ServerEngine.distill_step  - MMFL.distill body (MMFL.py:346-391)
  MMClient.private_step      - MMClientTrainer.train_epoch, private pass (MMClientTrainer.py:118-143)
  MMClient.contrast_step     - inter + intra contrast pass (MMClientTrainer.py:154-222)
  MMClient.generate          - MMClientTrainer.generate_logits (MMClientTrainer.py:326-359)
  UnimodalClient.*           - ClientTrainer.tra / extract_pub_feature (ClientTrainer.py:307-510, 631-664)
  aggregate                  - MMFL.distill.aggregation, `con_w` (MMFL.py:298-335)
  exchange_*                 - the one real exchange of the path when clients are sharded over GPUs: all-gather of
                               the public-set representations into the server ensemble

Everything numerical happens in libcreamfl_b200.so; this file sequences calls and owns buffers.  Apex O2 loss
scaling has no counterpart (bf16 needs none).

CUDA graphs.  A step function (`use_graphs=True`) is captured once per input shape and replayed:
  * learning-rate / hyper-parameter changes reach the captured optimizer kernels because `optimizer.prepare()`
    re-uploads the device hyper-parameter buffer before every replay;
  * the warm-up run that precedes a capture executes the real step, so everything the step mutates (parameters,
    optimizer moments and step counter, BatchNorm running statistics, criterion parameters, dropout RNG step) is
    snapshotted before and restored after it: the first batch of a new graph is applied exactly once;
  * caption lengths travel as a device int32 tensor (a graph input), so a graph is keyed by shapes only; the public
    banks of the contrast step live in persistent buffers refreshed in place; the cache is bounded (LRU);
  * all graphs of a device share one memory pool (replays are serialised on the launching stream).
"""
from __future__ import annotations

import copy
import weakref
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import ops
from .clients import ClientPCME, ImageClient, text_supervised_loss, unimodal_supervised_loss
from .criterions import get_criterion
from .optim import FusedOptimizer
from .text_towers import TextClient
from .towers import PCME, fork_stream

PCME_CRITERION_CFG = {'init_shift': 15, 'init_negative_scale': 15, 'num_samples': 7}    # coco.yaml:41-47
MAX_GRAPHS = 12           # cached graphs per engine before the least recently used one is dropped


def default_device(index: Optional[int] = None) -> torch.device:
    """The CUDA device the engines run on (current device unless an index is given).  There is no CPU path: without a
    CUDA device this raises instead of handing back a CPU device."""
    if not torch.cuda.is_available():
        raise RuntimeError('creamfl_b200 engines need a CUDA device (the hot path has no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device() if index is None else index)


def _features(output: Dict[str, torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    return output['image_features'], output['caption_features']


_POOLS: Dict[tuple, list] = {}        # (device index, lane) -> [pool handle, number of live graphs in the pool]


class lane:
    """`with lane(device, k): <steps>` - the steps are enqueued on lane k's own CUDA stream with lane-private scratch
    buffers and graph memory pool, so steps of different lanes (different clients hosted by one GPU: they share
    nothing inside a round, MMFL.py:226-247 trains them one after the other) may overlap on the device.  Steps of one
    lane stay ordered.  `lane.fork` makes the lanes wait for the work already enqueued on the current stream,
    `lane.join` makes the current stream wait for the lanes."""
    _streams: Dict[tuple, torch.cuda.Stream] = {}

    def __init__(self, device: torch.device, k: int):
        self.device, self.k = device, int(k)

    @staticmethod
    def stream(device: torch.device, k: int) -> torch.cuda.Stream:
        key = (device.index, int(k))
        if key not in lane._streams:
            lane._streams[key] = torch.cuda.Stream(device)
        return lane._streams[key]

    def __enter__(self):
        from . import tower_ops
        self.prev = tower_ops.set_lane(self.k)
        self.ctx = torch.cuda.stream(lane.stream(self.device, self.k))
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        from . import tower_ops
        self.ctx.__exit__(*exc)
        tower_ops.set_lane(self.prev)
        return False

    @staticmethod
    def fork(device: torch.device, lanes) -> None:
        cur = torch.cuda.current_stream(device)
        for k in lanes:
            lane.stream(device, k).wait_stream(cur)

    @staticmethod
    def join(device: torch.device, lanes) -> None:
        cur = torch.cuda.current_stream(device)
        for k in lanes:
            cur.wait_stream(lane.stream(device, k))


def _graph_pool(device: torch.device) -> list:
    """The memory pool shared by every captured step of a (device, lane) - graphs of one pool must never replay
    concurrently, graphs of different lanes may.  A pool dies with its last graph, so the handle is renewed once no
    graph of the previous pool is alive (engines come and go in tests)."""
    from . import tower_ops
    key = (device.index if device.index is not None else torch.cuda.current_device(), tower_ops.current_lane())
    entry = _POOLS.get(key)
    if entry is None or entry[1] == 0:
        entry = _POOLS[key] = [torch.cuda.graph_pool_handle(), 0]
    return entry


def _release_pool(entry: list) -> None:
    entry[1] -= 1


class GraphedStep:
    """One step function captured as a CUDA graph (the client steps issue ~1500 short kernels and are otherwise
    bound by launch overhead).  Inputs are copied into static buffers, the graph is replayed, outputs are static.

    `state`: callable returning the tensors the step mutates; they are restored after the warm-up run so that the
    step's side effects happen once per call (never during graph construction)."""

    def __init__(self, fn, example: Dict[str, torch.Tensor], state=None, warmup: int = 1):
        self.static = {k: v.clone() for k, v in example.items()}
        saved = [(t, t.clone()) for t in state()] if state is not None else []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(**self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if state is not None:
            now = state()
            if len(now) != len(saved) or any(a.data_ptr() != b[0].data_ptr() for a, b in zip(now, saved)):
                raise RuntimeError('GraphedStep: the step re-allocated part of its state during warm-up')
            for t, keep in saved:
                t.copy_(keep)
        del saved
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launches()
        dev = next(iter(self.static.values())).device
        pool = _graph_pool(dev)
        with torch.cuda.graph(self.graph, pool=pool[0], capture_error_mode='thread_local'):
            self.out = fn(**self.static)
        pool[1] += 1
        weakref.finalize(self, _release_pool, pool)
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


class _GraphCache:
    """Shape-keyed LRU cache of GraphedSteps shared by the engines."""

    def __init__(self, owner):
        self.owner = owner
        self.graphs: 'OrderedDict[tuple, GraphedStep]' = OrderedDict()

    def run(self, name, fn, mutates: bool, **tensors):
        from . import tower_ops
        # a graph bakes in its lane's scratch buffers and memory pool: one capture per lane the step is run in
        key = (name, tower_ops.current_lane(), tuple((k, tuple(t.shape), t.dtype) for k, t in tensors.items()))
        own = self.owner
        if own.optimizer is not None:
            own.optimizer.prepare()          # hyper-parameters (lr schedule) reach the captured kernels: re-uploaded
        g = self.graphs.get(key)             # outside the capture, before every replay
        if g is None:
            if len(self.graphs) >= MAX_GRAPHS:
                self.graphs.popitem(last=False)
            g = self.graphs[key] = GraphedStep(fn, tensors, state=own.state_tensors if mutates else None)
        else:
            self.graphs.move_to_end(key)
        return g(**tensors)


def _module_state(model, criterion, optimizer) -> List[torch.Tensor]:
    st = model.store()
    ts = [st.flat, st.shadow, st.grad] + [dst for dst, _ in st.padded] + [b for b in model.buffers()]
    if criterion is not None:
        ts += [p.data for p in criterion.parameters()]
    if optimizer is not None:
        ts += optimizer.state_tensors()
    bert = getattr(model, 'txt_enc', None)
    if bert is not None and hasattr(bert, 'dropout_state') and bert.dropout_p > 0.0:
        ts.append(bert.dropout_state(st.flat.device).rng)
    return ts


class ServerEngine:
    """The server of MMFL: PCME(ResNet101 + BERT) by default; `not_bert=True` builds the reference's
    `--not_bert` server, PCME(ResNet50 + GRU) (MMFL.py:82-85, pcme.py:28-29,37-38), whose text input is the
    (sentences, lengths) pair instead of BERT tokens."""

    def __init__(self, embed_dim: int = 256, cnn_type: str = 'resnet101', lr: float = 2e-4, grad_clip: float = 2.0,
                 kd_weight: float = 0.3, device: Optional[torch.device] = None, data_parallel: bool = False,
                 use_graphs: bool = False, bert_dropout: float = 0.1, not_bert: bool = False,
                 vocab_size: int = 11755):
        self.device = device or default_device()
        self.not_bert = bool(not_bert)
        if self.not_bert:
            self.model = ClientPCME(vocab_size, embed_dim, cnn_type=cnn_type).to(self.device)
        else:
            self.model = PCME(None, {'embed_dim': embed_dim, 'cnn_type': cnn_type, 'not_bert': False,
                                     'bert_dropout': bert_dropout}).to(self.device)
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
        self._cache = _GraphCache(self)

    def state_tensors(self) -> List[torch.Tensor]:
        """Everything a train / distill step mutates (restored after a graph warm-up run)."""
        self.optimizer.prepare()
        return _module_state(self.model, self.criterion, self.optimizer)

    # ------------------------------------------------------------------------------------------------ inputs
    def _text_inputs(self, tokens) -> Dict[str, torch.Tensor]:
        """Graph inputs of the text tower: BERT ids + mask, or (sentences, lengths) for the GRU server."""
        dev = self.device
        if self.not_bert:
            sentences, lengths = tokens
            return {'sentences': sentences.to(dev, non_blocking=True), 'len32': _len32(lengths, dev)}
        return {'ids': tokens['input_ids'].to(dev, non_blocking=True),
                'mask': tokens['attention_mask'].to(dev, non_blocking=True)}

    def _forward(self, images, txt: Dict[str, torch.Tensor]):
        if self.not_bert:
            return self.model(images, txt['sentences'], None, txt['len32'])
        return self.model(images, None, {'input_ids': txt['ids'], 'attention_mask': txt['mask']}, None)

    def _graphable(self, tokens) -> bool:
        if not self.use_graphs:
            return False
        if self.not_bert:
            return isinstance(tokens, (tuple, list)) and len(tokens) == 2 and torch.is_tensor(tokens[0])
        return isinstance(tokens, dict)

    def _eager_inputs(self, tokens):
        if self.not_bert or isinstance(tokens, dict):
            return self._text_inputs(tokens), None
        return None, tokens           # reference-style caption strings: tokenised inside PCME.forward

    def _sync_grads(self) -> None:
        """Replicated server, batches sharded over ranks: average the flat gradient buffer (one NCCL call)."""
        if self.data_parallel:
            average_gradients(self.model.store().grad)
            for p in self.criterion.parameters():
                average_gradients(p.grad)

    # ------------------------------------------------------------------------------------------------ train
    def train_step(self, images, tokens) -> torch.Tensor:
        """tokens: {'input_ids', 'attention_mask'} (BERT server; a tuple of caption strings is tokenised on the
        host like pcme.py:40-42 and runs un-graphed) or (sentences, lengths) for the `not_bert` server."""
        if self._graphable(tokens):
            fn = self._train_fwd_bwd if self.data_parallel else self._train_step
            loss = self._cache.run('train', lambda images, **txt: fn(images, txt), True, images=images,
                                   **self._text_inputs(tokens))
            if self.data_parallel:
                self._sync_grads()
                self.optimizer.step()
            return loss
        txt, raw = self._eager_inputs(tokens)
        return self._train_step(images, txt, raw)

    def _loss_fwd_bwd(self, images, txt, raw, loss_fn) -> torch.Tensor:
        self.model.train()
        output = self._forward(images, txt) if raw is None else self.model(images, None, raw, None)
        loss = loss_fn(output)
        self.optimizer.zero_grad()
        loss.backward()
        return loss.detach()

    def _train_fwd_bwd(self, images, txt, raw=None) -> torch.Tensor:
        return self._loss_fwd_bwd(images, txt, raw, lambda output: self.criterion(**output)[0])

    def _train_step(self, images, txt, raw=None) -> torch.Tensor:
        loss = self._train_fwd_bwd(images, txt, raw)
        self._sync_grads()
        self.optimizer.step()
        return loss

    # ------------------------------------------------------------------------------------------------ extract
    def extract(self, images, tokens) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._graphable(tokens):
            return self._cache.run('extract', lambda images, **txt: self._extract(images, txt), False, images=images,
                                   **self._text_inputs(tokens))
        txt, raw = self._eager_inputs(tokens)
        return self._extract(images, txt, raw)

    @torch.no_grad()
    def _extract(self, images, txt, raw=None) -> Tuple[torch.Tensor, torch.Tensor]:
        self.model.eval()
        return _features(self._forward(images, txt) if raw is None else self.model(images, None, raw, None))

    # ------------------------------------------------------------------------------------------------ distill
    def distill_step(self, images, tokens, d_idx, agg_img, agg_txt, img_terms: int = 1, txt_terms: int = 1):
        """`img_terms` / `txt_terms`: how many times the reference adds the image / text MSE - once per client type
        that carries the modality (MMFL.py:361-378; 2 each when image, text and multimodal clients all exist)."""
        if self._graphable(tokens) and (agg_img is not None or agg_txt is not None):
            # aggregated targets are passed as graph inputs (they are re-created every round)
            fn = self._distill_fwd_bwd if self.data_parallel else self._distill_step
            aggs = {k: v for k, v in (('agg_img', agg_img), ('agg_txt', agg_txt)) if v is not None}

            def run(images, d_idx, agg_img=None, agg_txt=None, **txt):
                return fn(images, txt, None, d_idx, agg_img, agg_txt, img_terms, txt_terms)
            loss = self._cache.run(('distill', img_terms, txt_terms), run, True, images=images, d_idx=d_idx, **aggs,
                                   **self._text_inputs(tokens))
            if self.data_parallel:
                self._sync_grads()
                self.optimizer.step()
            return loss
        txt, raw = self._eager_inputs(tokens)
        return self._distill_step(images, txt, raw, d_idx, agg_img, agg_txt, img_terms, txt_terms)

    def _distill_step(self, images, txt, raw, d_idx, agg_img, agg_txt, img_terms: int = 1, txt_terms: int = 1):
        loss = self._distill_fwd_bwd(images, txt, raw, d_idx, agg_img, agg_txt, img_terms, txt_terms)
        self._sync_grads()
        self.optimizer.step()
        return loss

    def _distill_fwd_bwd(self, images, txt, raw, d_idx, agg_img, agg_txt, img_terms: int = 1, txt_terms: int = 1):
        def kd(output):
            out_img, out_txt = _features(output)
            loss = 0
            if agg_img is not None and img_terms:
                loss = loss + (self.kd_weight * img_terms) * ops.mse_gather_loss(out_img, agg_img, d_idx)
            if agg_txt is not None and txt_terms:
                loss = loss + (self.kd_weight * txt_terms) * ops.mse_gather_loss(out_txt, agg_txt, d_idx)
            return loss
        return self._loss_fwd_bwd(images, txt, raw, kd)

    # ------------------------------------------------------------------------------------------------ checkpoint
    def save_checkpoint(self, path: str) -> None:
        """MMFL.py:281,284: torch.save({'net': model.state_dict()}, path) - keys and shapes equal the reference's."""
        torch.save({'net': {k: v.detach().cpu() for k, v in self.model.state_dict().items()}}, path)

    def load_checkpoint(self, path: str) -> None:
        state = torch.load(path, map_location='cpu')
        self.model.load_state_dict(state['net'] if 'net' in state else state, strict=True)
        self.model.sync_shadow()


class _Banked:
    """Persistent copies of the server's public features (fp32 + bf16) a client contrasts against.  The reference
    re-uploads them every epoch (MMClientTrainer.py:151); here they are refreshed in place so that captured graphs
    keep valid addresses."""

    def __init__(self):
        self._bank = {}
        self._bank_src = {}

    def set_bank(self, name: str, feats: torch.Tensor, device) -> Tuple[torch.Tensor, torch.Tensor]:
        src = (feats.data_ptr(), feats._version, tuple(feats.shape))
        cur = self._bank.get(name)
        if cur is not None and self._bank_src.get(name) == src:
            return cur
        if cur is None or cur[0].shape != feats.shape:
            cur = (torch.empty(feats.shape, dtype=torch.float32, device=device),
                   torch.empty(feats.shape, dtype=torch.bfloat16, device=device))
            self._bank[name] = cur
        if cur[0].data_ptr() != feats.data_ptr():
            cur[0].copy_(feats, non_blocking=True)
        ops.cast_into(cur[0].view(-1), cur[1].view(-1))
        self._bank_src[name] = src
        return cur


def _len32(lengths, device) -> torch.Tensor:
    if torch.is_tensor(lengths):
        return lengths.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()
    return torch.tensor([int(v) for v in lengths], dtype=torch.int32).to(device, non_blocking=True)


class MMClient(_Banked):
    def __init__(self, embed_dim: int = 256, lr: float = 2e-4, grad_clip: float = 2.0, interintra_weight: float = 0.5,
                 vocab_size: int = 11755, device: Optional[torch.device] = None, use_graphs: bool = False):
        super().__init__()
        self.device = device or default_device()
        self.use_graphs = use_graphs
        self.model = ClientPCME(vocab_size, embed_dim).to(self.device)
        self.criterion = get_criterion('pcme', PCME_CRITERION_CFG).to(self.device)
        self.model.store()
        params = [p for p in self.model.parameters() if p.requires_grad] + list(self.criterion.parameters())
        self.optimizer = FusedOptimizer(params, lr=lr, max_norm=grad_clip, mode='adamp',
                                        no_clip=list(self.criterion.parameters())).attach_stores(self.model)
        self.w = interintra_weight
        self.old_model = None
        self._cache = _GraphCache(self)

    def state_tensors(self) -> List[torch.Tensor]:
        self.optimizer.prepare()
        return _module_state(self.model, self.criterion, self.optimizer)

    def begin_round(self) -> None:
        """old_model = deepcopy(model) (MMClientTrainer.py:92-93); after the first round the copy is refreshed in
        place so that its addresses - and any captured graph - stay valid."""
        if self.old_model is None:
            self.old_model = copy.deepcopy(self.model).eval()
            self.old_model.store()
        else:
            self.old_model.copy_weights_from(self.model)
        self.model.train()

    def private_step(self, images, captions, lengths) -> torch.Tensor:
        len32 = _len32(lengths, self.device)
        if self.use_graphs:
            return self._cache.run('private', self._private_step, True, images=images, captions=captions,
                                   lengths=len32)
        return self._private_step(images, captions, len32)

    def _private_step(self, images, captions, lengths) -> torch.Tensor:
        self.model.train()
        output = self.model(images, captions, None, lengths)
        loss, _ = self.criterion(**output)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def contrast_step(self, images, captions, lengths, d_idx, g_img, g_txt, g_img16=None, g_txt16=None,
                      intra: bool = True, inter: bool = True, loss_scale: bool = False) -> torch.Tensor:
        """g_img / g_txt: fp32 [N_pub, D] server features (g_*16: optional bf16 copies made once per round; with
        CUDA graphs the client keeps its own persistent fp32 + bf16 copies, refreshed when the sources change)."""
        len32 = _len32(lengths, self.device)
        if self.use_graphs:
            bi, bt = self.set_bank('img', g_img, self.device), self.set_bank('txt', g_txt, self.device)
            return self._cache.run(
                ('contrast', intra, inter, loss_scale, bi[0].data_ptr(), bt[0].data_ptr()),
                lambda images, captions, lengths, d_idx: self._contrast_step(
                    images, captions, lengths, d_idx, bi[0], bt[0], bi[1], bt[1], intra, inter, loss_scale),
                True, images=images, captions=captions, lengths=len32, d_idx=d_idx)
        if g_img16 is None:
            g_img16, g_txt16 = ops.to_bf16(g_img), ops.to_bf16(g_txt)
        return self._contrast_step(images, captions, len32, d_idx, g_img, g_txt, g_img16, g_txt16, intra, inter,
                                   loss_scale)

    def _contrast_step(self, images, captions, lengths, d_idx, g_img, g_txt, g_img16, g_txt16, intra: bool = True,
                       inter: bool = True, loss_scale: bool = False) -> torch.Tensor:
        self.model.train()
        self.optimizer.zero_grad()
        b = images.shape[0]
        loss_intra = loss_inter = None
        # the old model's forward (no grad) is a third independent branch next to the two towers of the model: it runs
        # on its own forked stream (inside a captured graph: a parallel branch of the graph)
        with fork_stream(images.device, intra and getattr(self, 'overlap_old_model', True), slot=1) as side:
            if intra:
                with side, torch.no_grad():
                    self.old_model.overlap_towers = False
                    old_img, old_txt = _features(self.old_model(images, captions, None, lengths))
            out_img, out_txt = _features(self.model(images, captions, None, lengths))
        if intra:
            loss_intra = ops.moon_intra_loss(out_img, old_img, g_img, d_idx, 2.0, 2 * b) + \
                ops.moon_intra_loss(out_txt, old_txt, g_txt, d_idx, 2.0, 2 * b)
        if inter:
            loss_inter = ops.infonce_loss(out_img, g_txt16, d_idx, 2.0) + ops.infonce_loss(out_txt, g_img16, d_idx, 2.0)
        loss = combine_contrast(loss_intra, loss_inter, self.w, loss_scale)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def generate(self, images, captions, lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        len32 = _len32(lengths, self.device)
        if self.use_graphs:
            return self._cache.run('generate', self._generate, False, images=images, captions=captions, lengths=len32)
        return self._generate(images, captions, len32)

    @torch.no_grad()
    def _generate(self, images, captions, lengths) -> Tuple[torch.Tensor, torch.Tensor]:
        self.model.eval()
        return _features(self.model(images, captions, None, lengths))


def combine_contrast(loss_intra, loss_inter, weight: float, loss_scale: bool):
    """MMClientTrainer.py:203-206,264,308 / ClientTrainer.py:416-419: `(intra + inter) * w`, with `--loss_scale` the
    inter term is rescaled to the intra term's magnitude; a single active term carries no weight."""
    if loss_intra is not None and loss_inter is not None:
        if not loss_scale:
            return (loss_intra + loss_inter) * weight
        return (loss_intra + loss_inter / (loss_inter / loss_intra).detach()) * weight
    return loss_intra if loss_intra is not None else loss_inter


class UnimodalClient(_Banked):
    """A unimodal client of ClientTrainer.py: ResNet18 image classifier (`kind='image'`, resnet_client.py) or GRU text
    classifier (`kind='text'`, language_model.py) trained with SGD(lr 1e-4, momentum 0.9, wd 5e-5)
    (ClientTrainer.py:287-288); the contrast pass embeds the PUBLIC images (captions) with the same trunk."""

    def __init__(self, kind: str, num_class: int, embed_dim: int = 256, lr: float = 1e-4, scale: int = 128,
                 interintra_weight: float = 0.5, inter_distance: float = 4.0, vocab_size: int = 11755,
                 device: Optional[torch.device] = None, use_graphs: bool = False):
        super().__init__()
        if kind not in ('image', 'text'):
            raise ValueError(f'unknown client kind {kind}')
        self.kind, self.is_image = kind, kind == 'image'
        self.device = device or default_device()
        self.use_graphs = use_graphs
        if self.is_image:
            self.model = ImageClient(num_class=num_class, embed_dim=embed_dim, scale=scale).to(self.device)
        else:
            self.model = TextClient(vocab_size=vocab_size, embed_dim=embed_dim, num_class=num_class,
                                    scale=scale).to(self.device)
        self.model.store()
        self.criterion = None
        self.optimizer = FusedOptimizer(self.model.parameters(), lr=lr, momentum=0.9, weight_decay=5e-5,
                                        mode='sgd').attach_stores(self.model)
        self.w, self.inter_distance = interintra_weight, inter_distance
        self.old_model = None
        self._cache = _GraphCache(self)

    def state_tensors(self) -> List[torch.Tensor]:
        self.optimizer.prepare()
        return _module_state(self.model, None, self.optimizer)

    def begin_round(self) -> None:
        """old_model = deepcopy(model) (ClientTrainer.py:194-196), refreshed in place after the first round."""
        if self.old_model is None:
            self.old_model = copy.deepcopy(self.model)
            self.old_model.store()
        else:
            a, b = self.old_model.store(), self.model.store()
            a.flat.copy_(b.flat)
            a.shadow.copy_(b.shadow)
            for (dst, _), (src, _) in zip(a.padded, b.padded):
                dst.copy_(src)
            for p, q in zip(self.old_model.buffers(), self.model.buffers()):
                p.copy_(q)
        self.old_model.eval()

    @staticmethod
    def _mode(model, extract: bool) -> None:
        model.phase, model.is_train = ('extract_conv_feature', False) if extract else ('None', True)

    def _embed(self, model, x, len32):
        return model(x) if self.is_image else model(x, len32)

    # ---- supervised pass (ClientTrainer.py:322-363)
    def supervised_step(self, inputs, labels, lengths=None) -> torch.Tensor:
        tensors = {'inputs': inputs, 'labels': labels}
        if not self.is_image:
            tensors['lengths'] = _len32(lengths, self.device)
        if self.use_graphs:
            return self._cache.run('supervised', self._supervised_step, True, **tensors)
        return self._supervised_step(**tensors)

    def _supervised_step(self, inputs, labels, lengths=None) -> torch.Tensor:
        self.model.train()
        self._mode(self.model, False)
        self.optimizer.zero_grad()
        if self.is_image:
            loss, _ = unimodal_supervised_loss(self.model, inputs, labels, self.inter_distance)
        else:
            loss, _ = text_supervised_loss(self.model, inputs, lengths, labels, self.inter_distance)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    # ---- contrast pass over the public loader (ClientTrainer.py:365-421)
    def contrast_step(self, x, lengths, d_idx, g_same, g_other, intra: bool = True, inter: bool = True,
                      loss_scale: bool = False) -> torch.Tensor:
        """x: public images (image client) or public vocabulary-id captions (text client); g_same / g_other: the
        server's public features of the client's own / the opposite modality, fp32 [N_pub, D]."""
        tensors = {'x': x, 'd_idx': d_idx}
        if not self.is_image:
            tensors['lengths'] = _len32(lengths, self.device)
        if self.use_graphs:
            bs, bo = self.set_bank('same', g_same, self.device), self.set_bank('other', g_other, self.device)
            return self._cache.run(
                ('contrast', intra, inter, loss_scale, bs[0].data_ptr(), bo[0].data_ptr()),
                lambda x, d_idx, lengths=None: self._contrast_step(x, lengths, d_idx, bs[0], bo[1], intra, inter,
                                                                   loss_scale), True, **tensors)
        return self._contrast_step(x, tensors.get('lengths'), d_idx, g_same, ops.to_bf16(g_other), intra, inter,
                                   loss_scale)

    def _contrast_step(self, x, lengths, d_idx, g_same, g_other16, intra, inter, loss_scale) -> torch.Tensor:
        self.model.train()
        for m in (self.model, self.old_model):
            self._mode(m, True)                                                          # :372-375
        self.optimizer.zero_grad()
        with fork_stream(x.device, intra and getattr(self, 'overlap_old_model', True), slot=1) as side:
            if intra:
                with side, torch.no_grad():
                    old = self._embed(self.old_model, x, lengths)
            feat = self._embed(self.model, x, lengths)
        loss_inter = loss_moon = None
        if inter:
            loss_inter = ops.infonce_loss(feat, g_other16, d_idx, 2.0)                    # :388,398-401
        if intra:
            loss_moon = ops.moon_intra_loss(feat, old, g_same, d_idx, 2.0, feat.shape[0])  # :404-414
        loss = combine_contrast(loss_moon, loss_inter, self.w, loss_scale)
        loss.backward()
        self.optimizer.step()
        for m in (self.model, self.old_model):
            self._mode(m, False)
        return loss.detach()

    # ---- public representations (ClientTrainer.py:631-664; no .eval(): BatchNorm keeps batch statistics)
    def generate(self, x, lengths=None) -> torch.Tensor:
        tensors = {'x': x}
        if not self.is_image:
            tensors['lengths'] = _len32(lengths, self.device)
        if self.use_graphs:
            # BatchNorm running statistics move in this pass (train-mode BN): it mutates state
            return self._cache.run('generate', self._generate, True, **tensors)
        return self._generate(**tensors)

    @torch.no_grad()
    def _generate(self, x, lengths=None) -> torch.Tensor:
        self._mode(self.model, True)
        f = self._embed(self.model, x, lengths)
        self._mode(self.model, False)
        return f


def aggregate(vecs: Sequence[torch.Tensor], global_other: torch.Tensor) -> torch.Tensor:
    """con_w aggregation of one modality on one device (MMFL.py:298-314)."""
    return ops.conw_aggregate(list(vecs), global_other)


def gather_client_rows(score: torch.Tensor, vec: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The exchange step: every rank contributes contrastive scores [.., N_pub] and representations [.., N_pub, D];
    every rank receives them stacked in rank order (leading axis = world size).  Pure torch.distributed (NCCL on GPUs,
    gloo in the CPU tests); with one process it just adds the rank axis."""
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
    """One client per rank, every rank holds the modality: score locally (tcgen05, 1.28 TFLOP), all-gather scores
    (200 KB) and representations (51 MB), softmax-over-clients weighted sum on every rank.  With one
    rank this is `aggregate([own_vec], ...)`."""
    out = exchange_and_aggregate_clients([own_vec], global_other16)
    assert out is not None
    return out


def exchange_and_aggregate_clients(own_vecs: Sequence[Optional[torch.Tensor]], global_other16: torch.Tensor,
                                   layout: Optional[Sequence[Sequence[bool]]] = None) -> Optional[torch.Tensor]:
    """con_w aggregation of ONE modality when each rank hosts a list of clients, some of which do not carry the
    modality (image-only clients return txt = None and vice versa: ClientTrainer.py:622-629; the server then
    aggregates over the clients that returned it, MMFL.py:226-247,298-331).

    own_vecs : this rank's clients in order, fp32 [N_pub, D] or None (every rank passes the same number of slots)
    layout   : layout[rank][slot] = slot carries the modality, for ALL ranks; None = exchanged here (one tiny
               all-gather + host read)
    Every rank scores its own present clients against the replicated server features, the ranks all-gather the
    slot-stacked scores and representations (absent slots travel as masked zero slots - NCCL all-gather is
    fixed-size), and the reduce runs over the present slots only.  Returns None when no client has the modality."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    slots = len(own_vecs)
    present = [v is not None for v in own_vecs]
    if layout is None:
        if world > 1:
            flags = torch.tensor([int(p) for p in present], dtype=torch.int32, device=global_other16.device)
            allf = torch.empty(world * slots, dtype=torch.int32, device=flags.device)
            dist.all_gather_into_tensor(allf, flags)
            layout = allf.view(world, slots).bool().tolist()
        else:
            layout = [present]
    if list(layout[rank]) != present:
        raise ValueError('exchange_and_aggregate_clients: layout disagrees with the representations passed in')
    n_total = sum(int(p) for row in layout for p in row)
    if n_total == 0:
        return None
    n_pub, d = global_other16.shape
    dev = global_other16.device
    if world == 1:
        vecs = [v for v in own_vecs if v is not None]
        scores = torch.stack([ops.conw_score(ops.to_bf16(v), global_other16) for v in vecs], dim=0)
        return ops.conw_reduce(vecs, scores)
    # only slots some rank fills travel: a column of the layout that is empty on every rank is skipped
    live = [s for s in range(slots) if any(row[s] for row in layout)]
    score = torch.zeros((len(live), n_pub), dtype=torch.float32, device=dev)
    vec = torch.zeros((len(live), n_pub, d), dtype=torch.float32, device=dev)
    for k, s in enumerate(live):
        if own_vecs[s] is not None:
            vec[k].copy_(own_vecs[s])
            score[k].copy_(ops.conw_score(ops.to_bf16(own_vecs[s]), global_other16))
    scores, vecs = gather_client_rows(score, vec)             # [world, live, N_pub], [world, live, N_pub, D]
    pick = [(r, k) for r in range(world) for k, s in enumerate(live) if layout[r][s]]
    sel_scores = torch.stack([scores[r, k] for r, k in pick], dim=0)
    return ops.conw_reduce([vecs[r, k] for r, k in pick], sel_scores)
