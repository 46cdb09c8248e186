"""CUDA-event timing of every C-ABI call of one eager step, grouped into kernel families - measurement plumbing of
bench.py and scripts/ (no arithmetic here).

`CallTimer` swaps the ctypes handle of creamfl_b200/_lib.py for a proxy that records a CUDA event on the launching
stream before and after each `creamfl_*` call.  The caller first stalls the GPU (`stall()`), enqueues the step, and
reads the events afterwards: the host runs ahead of the device, so consecutive kernels start back to back and an
event pair brackets device time only (no launch latency inside the interval).  Algorithmic FLOPs / bytes per call
are derived from the call's own dimension arguments (SURVEY.md 8d / DESIGN.md section 3 conventions).
"""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Tuple

import torch

from . import _lib


def _gemm_cost(args, drop=False):
    # creamfl_gemm_bf16(a, lda, a_mn, b, ldb, b_mn, M, N, K, ...)
    a_mn, b_mn, m, n, k = int(args[2]), int(args[5]), int(args[6]), int(args[7]), int(args[8])
    fam = 'gemm_tc_wgrad' if a_mn else ('gemm_tc_dgrad' if b_mn else 'gemm_tc_fwd')
    out_bytes = 4 if a_mn else 2                     # weight gradients accumulate in fp32
    return fam, 2.0 * m * n * k, 2.0 * (m * k + n * k) + out_bytes * m * n, f'{m}x{n}x{k}'


def _conv_cost(name, args):
    # creamfl_conv2d_*(x|dy, w|x, [..], n, h, w, cin, cout, r, s, stride, pad, ...): dims start at index 2 (fprop,
    # dgrad) or 3 (wgrad: dy, x, null).  Families are named after the KERNEL that runs (capi_nn.cu picks it): 1x1
    # stride-1 convolutions are plain GEMMs on gemm_tc, stride-1 "same" 3x3 run the implicit GEMM conv_tc, strided
    # ones go through an explicit im2col / col2im + gemm_tc.
    i = 3 if name.endswith('wgrad') else 2
    n, h, w, cin, cout, r, s, stride, pad = (int(v) for v in args[i:i + 9])
    ho, wo = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1
    flops = 2.0 * n * ho * wo * cout * r * s * cin
    byt = 2.0 * (n * h * w * cin + n * ho * wo * cout) + (4.0 if name.endswith('wgrad') else 2.0) * cout * r * s * cin
    mode = name.replace('creamfl_conv2d_', '').replace('fprop', 'fwd')
    if r == 1 and s == 1 and stride == 1:
        fam = 'gemm_tc_' + mode
    elif stride == 1 and r == s and (r & 1) and cin % 64 == 0 and cout % 64 == 0:
        fam = 'conv_tc_' + mode
    else:
        fam = 'gemm_tc_im2col_' + mode
    return fam, flops, byt, f'{n}x{h}x{w} {cin}->{cout} {r}x{s}/{stride}'


def _bn_cost(name, args):
    if name in ('creamfl_bn_train_fwd', 'creamfl_bn_eval_fwd', 'creamfl_bn_train_fwd_mask'):
        p, c = int(args[1]), int(args[2])
        return 'bn_fwd', 0.0, 2.0 * p * c * 2, f'{p}x{c}'
    p, c = int(args[3]), int(args[4])                # bn_train_bwd[_mask](dy, y | mask, x, p, c, ...)
    return 'bn_bwd', 0.0, 2.0 * p * c * 3, f'{p}x{c}'


_FAMILY = {
    'creamfl_optimizer_step': 'optimizer', 'creamfl_layernorm_fwd': 'layernorm', 'creamfl_layernorm_bwd': 'layernorm',
    'creamfl_layernorm_fwd_drop': 'layernorm', 'creamfl_layernorm_bwd_drop': 'layernorm',
    'creamfl_attn_fwd': 'attention', 'creamfl_attn_bwd': 'attention', 'creamfl_attn_fwd_drop': 'attention',
    'creamfl_attn_bwd_drop': 'attention', 'creamfl_embed_fwd': 'embedding', 'creamfl_embed_bwd': 'embedding',
    'creamfl_maxpool_fwd': 'pool', 'creamfl_maxpool_bwd': 'pool', 'creamfl_im2col_nchw_f32': 'im2col',
    'creamfl_stem_fprop': 'stem_tc', 'creamfl_stem_wgrad': 'stem_tc',
    'creamfl_pcme_fwd': 'loss', 'creamfl_pcme_bwd': 'loss', 'creamfl_infonce_fwd': 'infonce',
    'creamfl_infonce_bwd': 'infonce', 'creamfl_conw_score': 'conw_score', 'creamfl_conw_reduce': 'conw_reduce',
}


def classify(name: str, args) -> Tuple[str, float, float, str]:
    """(family, algorithmic FLOPs, algorithmic bytes, shape tag) of one C-ABI call."""
    if name in ('creamfl_gemm_bf16', 'creamfl_gemm_bf16_drop'):
        return _gemm_cost(args)
    if name.startswith('creamfl_conv2d_') and not name.endswith('bytes'):
        return _conv_cost(name, args)
    if name in ('creamfl_bn_train_fwd', 'creamfl_bn_eval_fwd', 'creamfl_bn_train_bwd', 'creamfl_bn_train_fwd_mask',
                'creamfl_bn_train_bwd_mask'):
        return _bn_cost(name, args)
    return _FAMILY.get(name, 'other'), 0.0, 0.0, ''


class _Proxy:
    def __init__(self, lib, timer):
        self._lib, self._timer = lib, timer

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.startswith('creamfl_') or name.endswith('_bytes') or name in (
                'creamfl_last_error', 'creamfl_abi_version', 'creamfl_plan_split_k'):
            return fn
        timer = self._timer

        def timed(*args):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            timer.calls.append((name, classify(name, args), a, b))
            return rc
        return timed


class CallTimer:
    """with CallTimer() as t: t.stall(); step(); -> t.families()"""

    def __init__(self):
        self.calls: List[tuple] = []
        self._saved = None

    def __enter__(self):
        self._saved = _lib.load()
        _lib._lib = _Proxy(self._saved, self)
        return self

    def __exit__(self, *exc):
        _lib._lib = self._saved
        return False

    @staticmethod
    def stall(ms: float = 80.0) -> None:
        """Keep the stream busy for ~ms so that the host enqueues the whole step ahead of the device."""
        torch.cuda._sleep(int(ms * 1e-3 * 1.9e9))

    def overhead_ms(self) -> float:
        """Device time between two back-to-back event records with nothing in between (median of 16 pairs, the
        stream kept busy first): subtracted from every call's interval."""
        if getattr(self, '_overhead', None) is None:
            self.stall(5.0)
            pairs = []
            for _ in range(16):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                b.record()
                pairs.append((a, b))
            torch.cuda.synchronize()
            ts = sorted(a.elapsed_time(b) for a, b in pairs)
            self._overhead = ts[len(ts) // 2]
        return self._overhead

    def _ms(self, a, b) -> float:
        return max(0.0, a.elapsed_time(b) - self.overhead_ms())

    def families(self) -> Dict[str, dict]:
        torch.cuda.synchronize()
        fam: Dict[str, dict] = defaultdict(lambda: {'ms': 0.0, 'calls': 0, 'flops': 0.0, 'bytes': 0.0})
        for _name, (f, flops, byt, _tag), a, b in self.calls:
            d = fam[f]
            d['ms'] += self._ms(a, b)
            d['calls'] += 1
            d['flops'] += flops
            d['bytes'] += byt
        return dict(fam)

    def shapes(self, family: str) -> Dict[str, dict]:
        """Per-shape breakdown of one family (ms, calls, FLOPs)."""
        out: Dict[str, dict] = defaultdict(lambda: {'ms': 0.0, 'calls': 0, 'flops': 0.0, 'bytes': 0.0})
        for _name, (f, flops, byt, tag), a, b in self.calls:
            if f == family:
                d = out[tag]
                d['ms'] += self._ms(a, b)
                d['calls'] += 1
                d['flops'] += flops
                d['bytes'] += byt
        return dict(out)
