"""Pinned-memory prefetch of step inputs: host -> device copies run on their own CUDA stream into a ring of device
buffers, so the copy of step k + 1 overlaps the kernels of step k (the input side of the hot
path; the reference moves every batch with a blocking `.cuda()` inside the step: MMClientTrainer.py:157-160,
retrieval_trainer.py:190-196).

    pf = Prefetcher(device, depth=2)
    pf.submit(first_host_batch)
    for nxt in rest:
        batch = pf.next()          # current stream waits for the copy event; tensors live until `depth` submits later
        pf.submit(nxt)             # copy of the next step starts now, on the copy stream
        step(batch)

A host batch is a (nested) dict / list / tuple of CPU tensors; pageable tensors are staged through pinned buffers of
the same ring slot (once pinned, a tensor is copied directly).  Plumbing only: no arithmetic, no fallback - a CPU
device raises.
"""
from __future__ import annotations

from collections import deque
from typing import Any

import torch


def _map(obj: Any, fn):
    if torch.is_tensor(obj):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map(v, fn) for v in obj)
    return obj


class Prefetcher:
    def __init__(self, device: torch.device, depth: int = 2):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('Prefetcher copies to a CUDA device (there is no CPU path in creamfl_b200)')
        if depth < 1:
            raise ValueError('depth must be >= 1')
        self.device, self.depth = device, depth
        self.stream = torch.cuda.Stream(device)
        self._slots = [dict() for _ in range(depth)]           # per ring slot: (shape, dtype, index) -> device buffer
        self._free_events = [None] * depth                     # consumer finished with the slot
        self._pending = deque()                                # (slot, device batch, copy-done event)
        self._next_slot = 0
        self.bytes_copied = 0

    def _buffer(self, slot: int, t: torch.Tensor, counter: list) -> torch.Tensor:
        key = (tuple(t.shape), t.dtype, counter[0])
        counter[0] += 1
        buf = self._slots[slot].get(key)
        if buf is None:
            buf = self._slots[slot][key] = torch.empty(t.shape, dtype=t.dtype, device=self.device)
        return buf

    def submit(self, host_batch: Any) -> None:
        """Start copying `host_batch` to the device on the copy stream."""
        if len(self._pending) >= self.depth:
            raise RuntimeError(f'Prefetcher: {self.depth} batches already in flight; call next() first')
        slot = self._next_slot
        self._next_slot = (slot + 1) % self.depth
        counter = [0]
        with torch.cuda.stream(self.stream):
            if self._free_events[slot] is not None:
                self.stream.wait_event(self._free_events[slot])     # the step that used this slot has finished reading

            def copy(t: torch.Tensor) -> torch.Tensor:
                if t.device.type != 'cpu':
                    return t
                src = t if t.is_pinned() else t.pin_memory()
                dst = self._buffer(slot, t, counter)
                dst.copy_(src, non_blocking=True)
                self.bytes_copied += t.numel() * t.element_size()
                return dst
            dev = _map(host_batch, copy)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._pending.append((slot, dev, done))

    def next(self) -> Any:
        """Oldest submitted batch; the CURRENT stream waits for its copies (no host synchronisation)."""
        if not self._pending:
            raise RuntimeError('Prefetcher.next() without a submitted batch')
        slot, dev, done = self._pending.popleft()
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(done)
        self._last = (slot, cur)
        return dev

    def release(self) -> None:
        """Mark the batch returned by the last next() as consumed by everything enqueued on the current stream so far
        (its ring slot may be overwritten by a later submit)."""
        slot, cur = self._last
        ev = torch.cuda.Event()
        ev.record(cur)
        self._free_events[slot] = ev
