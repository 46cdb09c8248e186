"""Fused optimizer of the hot path: one C-ABI call (four kernel launches) per step for any number of tensors.

FusedOptimizer mirrors what the reference builds in src/algorithms/optimizers.py:7-31 (`adamp.AdamP`, third-party
adamp==0.3.0) followed by `nn.utils.clip_grad_norm_` in the step loops (retrieval_trainer.py:211-214,
MMClientTrainer.py:133-135): global-norm clipping is folded into the same pass.  mode='sgd' serves the unimodal
clients (torch.optim.SGD(lr 1e-4, momentum 0.9, weight_decay 5e-5), ClientTrainer.py:287-288).

It subclasses torch.optim.Optimizer so that the reference's lr schedulers (CosineAnnealingLR, optimizers.py:53-55)
drive it unchanged; hyper-parameters and the step counter live on the device.  A captured CUDA graph of the step reads
the device hyper-parameter buffer, which `prepare()` re-uploads whenever `param_groups` changed: the engines call it
before EVERY graph replay (engine._GraphCache.run), so a scheduler step reaches the captured kernels.
"""
from __future__ import annotations

from typing import Iterable, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import ops as _ops
from .ops import _p, _stream

_ROW = np.dtype([('p', '<u8'), ('g', '<u8'), ('m', '<u8'), ('v', '<u8'), ('shadow', '<u8'), ('len', '<i4'),
                 ('tensor', '<i4')])
_TENSOR = np.dtype([('row_begin', '<i4'), ('row_end', '<i4'), ('project', '<i4'), ('clip', '<i4'), ('numel', '<i8')])
_MODES = {'adamp': 0, 'adam': 1, 'sgd': 2}


class FusedOptimizer(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, delta: float = 0.1, wd_ratio: float = 0.1, momentum: float = 0.9,
                 max_norm: float = 0.0, mode: str = 'adamp', no_clip: Sequence[torch.nn.Parameter] = (),
                 chunk: int = 8192):
        if mode not in _MODES:
            raise ValueError(f'unknown mode {mode}')
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, delta=delta,
                                      wd_ratio=wd_ratio, momentum=momentum, max_norm=max_norm))
        self.mode = mode
        self._no_clip = {id(p) for p in no_clip}
        self._chunk = chunk
        self._built = False

    # ------------------------------------------------------------------------------------------------ tables
    def _build(self) -> None:
        params = [p for g in self.param_groups for p in g['params']]
        if len(self.param_groups) != 1:
            raise NotImplementedError('FusedOptimizer supports a single parameter group')
        dev = params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FusedOptimizer needs CUDA parameters (no CPU fallback exists)')
        rows, tensors = [], []
        self._keep = []
        self._loose = []
        for ti, p in enumerate(params):
            n = p.numel()
            if hasattr(p, '_g2d'):                      # lives in a ParamStore: memory-contiguous fp32 + flat grad
                pm = p._g2d.new_empty(0)                # placeholder to get dtype/device
                p_ptr = p.data_ptr()
                g_ptr = p._g2d.data_ptr()
                w16 = p._w16
                if p.grad is None:
                    p.grad = p._gview
            else:
                if not p.data.is_contiguous():
                    p.data = p.data.contiguous()
                gbuf = torch.zeros_like(p.data)
                if p.grad is not None:
                    gbuf.copy_(p.grad)
                p.grad = gbuf
                self._loose.append((p, gbuf))
                p_ptr, g_ptr, w16 = p.data_ptr(), gbuf.data_ptr(), None
            m = torch.zeros(n, dtype=torch.float32, device=dev)
            v = torch.zeros(n, dtype=torch.float32, device=dev) if self.mode != 'sgd' else m
            self._keep += [m, v]
            st = self.state[p]
            st['exp_avg'], st['exp_avg_sq'] = m, v
            project = 1 if p.dim() > 1 else 0
            row_begin = len(rows)
            if project:
                nrows, rlen = p.shape[0], n // p.shape[0]
                for r in range(nrows):
                    sh = 0
                    if w16 is not None:
                        sh = w16.data_ptr() + r * w16.stride(0) * 2
                    rows.append((p_ptr + 4 * r * rlen, g_ptr + 4 * r * rlen, m.data_ptr() + 4 * r * rlen,
                                 v.data_ptr() + 4 * r * rlen, sh, rlen, ti))
            else:
                for s in range(0, n, self._chunk):
                    ln = min(self._chunk, n - s)
                    sh = (w16.data_ptr() + 2 * s) if w16 is not None else 0
                    rows.append((p_ptr + 4 * s, g_ptr + 4 * s, m.data_ptr() + 4 * s, v.data_ptr() + 4 * s, sh, ln, ti))
            tensors.append((row_begin, len(rows), project, 0 if id(p) in self._no_clip else 1, n))
        self._rows = torch.from_numpy(np.array(rows, dtype=_ROW).view(np.uint8)).to(dev)
        self._tensors = torch.from_numpy(np.array(tensors, dtype=_TENSOR).view(np.uint8)).to(dev)
        self._n_rows, self._n_tensors = len(rows), len(tensors)
        self._hyper = torch.zeros(9, dtype=torch.float32, device=dev)
        self._hyper_host = None
        self._state = torch.zeros(3, dtype=torch.float32, device=dev)
        self._total = torch.zeros(1, dtype=torch.float64, device=dev)
        self._stats = torch.empty(3 * self._n_rows, dtype=torch.float32, device=dev)
        self._flag = torch.zeros(self._n_tensors, dtype=torch.int32, device=dev)
        self._tnorm = torch.zeros(self._n_tensors, dtype=torch.float32, device=dev)
        self._lacc = torch.zeros(self._n_tensors, dtype=torch.float32, device=dev)
        self._built = True

    def _sync_hyper(self) -> None:
        g = self.param_groups[0]
        b1 = g['momentum'] if self.mode == 'sgd' else g['betas'][0]
        host = (float(g['lr']), float(b1), float(g['betas'][1]), float(g['eps']), float(g['weight_decay']),
                float(g['delta']), float(g['wd_ratio']), float(g['max_norm']), float(_MODES[self.mode]))
        if host != self._hyper_host:
            self._hyper.copy_(torch.tensor(host, dtype=torch.float32), non_blocking=False)
            self._hyper_host = host

    # ------------------------------------------------------------------------------------------------ API
    def prepare(self) -> None:
        """Build the device tables and upload hyper-parameters (call before CUDA-graph capture)."""
        if not self._built:
            self._build()
        self._sync_hyper()

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError('closures are not supported')
        self.prepare()
        for p, gbuf in self._loose:                     # autograd may have re-allocated .grad after set_to_none
            if p.grad is None:
                gbuf.zero_()
                p.grad = gbuf
            elif p.grad.data_ptr() != gbuf.data_ptr():
                gbuf.copy_(p.grad)
                p.grad = gbuf
        _lib.check(_lib.load().creamfl_optimizer_step(
            _p(self._rows), self._n_rows, _p(self._tensors), self._n_tensors, _p(self._hyper), _p(self._state),
            _p(self._total), _p(self._stats), _p(self._flag), _p(self._tnorm), _p(self._lacc), _stream()),
            'optimizer_step')
        _ops._launches += 4

    def state_tensors(self):
        """Every device tensor a step() mutates besides the parameters themselves: moments, the step counter / clip
        state and scratch, and the gradient buffers of parameters outside a ParamStore (used to snapshot / restore
        around the warm-up run of a CUDA-graph capture)."""
        self.prepare()
        return list(self._keep) + [self._state, self._total, self._flag, self._tnorm, self._lacc] + \
            [g for _, g in self._loose]

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Zeroes in place (the kernels accumulate into persistent gradient buffers)."""
        if not self._built:
            self._build()
        done = set()
        for g in self.param_groups:
            for p in g['params']:
                if hasattr(p, '_g2d'):
                    continue
                if p.grad is not None:
                    p.grad.zero_()
        for st in self._stores():
            if id(st) not in done:
                st.zero_grad()
                done.add(id(st))

    def _stores(self):
        return getattr(self, '_param_stores', [])

    def attach_stores(self, *models) -> 'FusedOptimizer':
        """Models owning a ParamStore whose flat gradient buffer zero_grad() should clear with one memset."""
        self._param_stores = [m.store() for m in models]
        return self

    @property
    def grad_norm(self) -> torch.Tensor:
        """Total gradient norm seen by the last step (device scalar)."""
        return self._state[2]
