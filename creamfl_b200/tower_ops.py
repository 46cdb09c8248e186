"""Thin tensor-level wrappers of the encoder-tower entry points of the C ABI (include/creamfl_b200.h).

No autograd here: these are the forward / backward primitives the block-level autograd Functions in
creamfl_b200/towers.py are composed of.  Activations are NHWC bf16 (4-D tensors [N, H, W, C], contiguous) or
[rows, features] bf16; every function raises if handed a CPU tensor.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from . import ops as _ops
from .ops import _p, _stream, _need_cuda

BF16 = torch.bfloat16

_ws_cache: dict = {}
_lane = 0


def set_lane(k: int) -> int:
    """Execution lane of the work enqueued from now on (engine.lane): scratch buffers and forked streams are per lane,
    so that steps of different lanes - e.g. two clients hosted by one GPU - may run concurrently, also as replays of
    captured graphs (a graph bakes in the addresses of the scratch buffers it was captured with).  Returns the
    previous lane."""
    global _lane
    old, _lane = _lane, int(k)
    return old


def current_lane() -> int:
    return _lane


def workspace(nbytes: int, device) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream, lane): reuse is stream-ordered (contents are dead once the
    consuming kernel has been enqueued), and towers running concurrently on forked streams never share one."""
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream if device.type == 'cuda' else 0,
           _lane)
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _ws_cache[key] = buf
    return buf


def _chk(rc: int, what: str, launches: int = 1) -> None:
    _lib.check(rc, what)
    _ops._launches += launches


# --------------------------------------------------------------------------------------------------- dropout
class DropoutState:
    """Device-side RNG state of the fused dropouts: int64[2] = {seed, step} (include/creamfl_b200.h, "BERT dropout").
    `tick()` advances the step with a one-thread kernel (capturable: a replayed CUDA graph draws new masks);
    `snapshot()` is the frozen copy one forward/backward pair reads; `keep_mask` exports the mask of a site."""

    def __init__(self, seed: int, p: float, device):
        self.p = float(p)
        self.rng = torch.tensor([int(seed) & 0x7fffffffffffffff, 0], dtype=torch.int64, device=device)

    def tick(self) -> None:
        _need_cuda(self.rng)
        _chk(_lib.load().creamfl_rng_tick(_p(self.rng), _stream()), "rng_tick")

    def snapshot(self) -> "DropoutState":
        snap = DropoutState.__new__(DropoutState)
        snap.p, snap.rng = self.p, self.rng.clone()
        return snap

    def keep_mask(self, site: int, n: int) -> torch.Tensor:
        """uint8 [n]: 1 where element e of `site` survives at the current step."""
        _need_cuda(self.rng)
        out = torch.empty(n, dtype=torch.uint8, device=self.rng.device)
        _chk(_lib.load().creamfl_dropout_mask(_p(self.rng), int(site), int(n), self.p, _p(out), _stream()),
             "dropout_mask")
        return out


def _drop_args(drop):
    """(rng pointer, site, p) of an optional (DropoutState, site) pair."""
    if drop is None:
        return None, -1, 0.0
    state, site = drop
    return _p(state.rng), int(site), float(state.p)


def gemm_drop(a, b, bias, add, drop, out=None):
    """out = dropout(a b^T + bias) + add  (bf16; HF BertSelfOutput / BertOutput: dense -> dropout -> + residual)."""
    _need_cuda(a, b, bias, add, out)
    m, k = a.shape
    n = b.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=BF16, device=a.device)
    rng, site, pd = _drop_args(drop)
    _chk(_lib.load().creamfl_gemm_bf16_drop(_p(a), a.stride(0), 0, _p(b), b.stride(0), 0, m, n, k, _p(out),
                                            out.stride(0), 1, _p(bias), _p(add), add.stride(0) if add is not None else 0,
                                            1, rng, site, pd, _stream()), "gemm_bf16_drop")
    return out


# --------------------------------------------------------------------------------------------------- convolution
def conv_out_hw(h: int, w: int, r: int, s: int, stride: int, pad: int):
    return (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1


def conv_fprop(x: torch.Tensor, w2d: torch.Tensor, r: int, s: int, stride: int, pad: int,
               bn_sums: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [N,H,W,Cin] bf16, w2d [Cout, >= R*S*Cin] bf16 (row = one filter in (r, s, cin) order).  bn_sums (fp64
    [2*Cout], zero on entry): receives the BatchNorm statistics of the output."""
    _need_cuda(x, w2d)
    n, h, w, cin = x.shape
    cout = w2d.shape[0]
    ho, wo = conv_out_hw(h, w, r, s, stride, pad)
    y = torch.empty((n, ho, wo, cout), dtype=BF16, device=x.device)
    lib = _lib.load()
    nb = lib.creamfl_conv2d_workspace_bytes(n, h, w, cin, cout, r, s, stride, pad)
    ws = workspace(nb, x.device) if nb else None
    _chk(lib.creamfl_conv2d_fprop(_p(x), _p(w2d), n, h, w, cin, cout, r, s, stride, pad, w2d.stride(0), _p(y),
                                  _p(bn_sums), _p(ws), nb, _stream()), "conv2d_fprop", 2 if nb else 1)
    return y


def conv_fprop_affine(x: torch.Tensor, w2d: torch.Tensor, r: int, s: int, stride: int, pad: int, bias: torch.Tensor,
                      add: Optional[torch.Tensor] = None, relu: bool = True) -> torch.Tensor:
    """Inference convolution with the following BatchNorm folded in: y = [relu](conv(x, w2d) + bias [+ add])
    (w2d / bias from bn_fold_layers)."""
    _need_cuda(x, w2d, bias, add)
    n, h, w, cin = x.shape
    cout = w2d.shape[0]
    ho, wo = conv_out_hw(h, w, r, s, stride, pad)
    y = torch.empty((n, ho, wo, cout), dtype=BF16, device=x.device)
    lib = _lib.load()
    nb = lib.creamfl_conv2d_workspace_bytes(n, h, w, cin, cout, r, s, stride, pad)
    ws = workspace(nb, x.device) if nb else None
    _chk(lib.creamfl_conv2d_fprop_affine(_p(x), _p(w2d), n, h, w, cin, cout, r, s, stride, pad, w2d.stride(0), _p(bias),
                                         _p(add), int(relu), _p(y), _p(ws), nb, _stream()), "conv2d_fprop_affine",
         2 if nb else 1)
    return y


def bn_fold_layers(layers: torch.Tensor, row_start: torch.Tensor, total_rows: int, eps: float, pairs=None) -> None:
    """layers [n, 10] int64 / row_start [n + 1] int64 device tables (include/creamfl_b200.h: creamfl_bn_fold_layers).
    `pairs` (the modules behind the table) is unused here; the kernel reads the addresses in the table."""
    _need_cuda(layers, row_start)
    _chk(_lib.load().creamfl_bn_fold_layers(_p(layers), _p(row_start), layers.shape[0], int(total_rows), eps, _stream()),
         "bn_fold_layers", 1)


def conv_dgrad(dy: torch.Tensor, w2d: torch.Tensor, x_shape, r: int, s: int, stride: int, pad: int,
               add: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(dy, w2d, add)
    n, h, w, cin = x_shape
    cout = w2d.shape[0]
    dx = torch.empty((n, h, w, cin), dtype=BF16, device=dy.device)
    lib = _lib.load()
    nb = lib.creamfl_conv2d_workspace_bytes(n, h, w, cin, cout, r, s, stride, pad)
    ws = workspace(nb, dy.device) if nb else None
    _chk(lib.creamfl_conv2d_dgrad(_p(dy), _p(w2d), n, h, w, cin, cout, r, s, stride, pad, w2d.stride(0), _p(add),
                                  _p(dx), _p(ws), nb, _stream()), "conv2d_dgrad", 2 if nb else 1)
    return dx


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, r: int, s: int, stride: int, pad: int) -> None:
    """dw (fp32, [Cout, R*S*Cin] contiguous memory) += dy^T patches(x)."""
    _need_cuda(dy, x, dw)
    n, h, w, cin = x.shape
    cout = dy.shape[-1]
    lib = _lib.load()
    nb = lib.creamfl_conv2d_workspace_bytes(n, h, w, cin, cout, r, s, stride, pad)
    ws = workspace(nb, dy.device) if nb else None
    _chk(lib.creamfl_conv2d_wgrad(_p(dy), _p(x), None, n, h, w, cin, cout, r, s, stride, pad, _p(dw), _p(ws), nb,
                                  _stream()), "conv2d_wgrad", 2 if nb else 1)


def im2col_images(images: torch.Tensor, r: int, s: int, stride: int, pad: int, pitch: int) -> torch.Tensor:
    """fp32 NCHW images -> bf16 patch matrix [N*Ho*Wo, pitch] in the shared workspace (valid until the next
    workspace user is enqueued... callers consume it immediately)."""
    _need_cuda(images)
    n, c, h, w = images.shape
    ho, wo = conv_out_hw(h, w, r, s, stride, pad)
    col = torch.empty((n * ho * wo, pitch), dtype=BF16, device=images.device)
    _chk(_lib.load().creamfl_im2col_nchw_f32(_p(images), n, c, h, w, r, s, stride, pad, pitch, _p(col), _stream()),
         "im2col_nchw_f32")
    return col


def stem_supported(h: int, w: int) -> bool:
    return bool(_lib.load().creamfl_stem_supported(int(h), int(w)))


def stem_fprop(images: torch.Tensor, w16: torch.Tensor) -> torch.Tensor:
    """fp32 NCHW images, bf16 filters [64, pitch >= 147] -> bf16 NHWC [N, Ho, Wo, 64] (conv 7x7/2 pad 3), patches
    assembled in shared memory."""
    _need_cuda(images, w16)
    n, c, h, w = images.shape
    ho, wo = conv_out_hw(h, w, 7, 7, 2, 3)
    y = torch.empty((n, ho, wo, 64), dtype=BF16, device=images.device)
    _chk(_lib.load().creamfl_stem_fprop(_p(images), n, h, w, _p(w16), w16.stride(0), _p(y), _stream()), "stem_fprop")
    return y


def stem_wgrad(images: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor) -> None:
    """dw (fp32 [64, 147], contiguous) += dy^T patches(images); dy bf16 NHWC [N, Ho, Wo, 64]."""
    _need_cuda(images, dy, dw)
    n, c, h, w = images.shape
    _chk(_lib.load().creamfl_stem_wgrad(_p(images), _p(dy), n, h, w, _p(dw), _stream()), "stem_wgrad")


# --------------------------------------------------------------------------------------------------- BatchNorm
class BNScratch:
    """Per-layer scratch of the BatchNorm kernels (fp64 sums, per-channel affine / backward coefficients)."""

    def __init__(self, c: int, device):
        self.sums = torch.zeros(2 * c, dtype=torch.float64, device=device)
        self.scale = torch.empty(c, dtype=torch.float32, device=device)
        self.shift = torch.empty(c, dtype=torch.float32, device=device)
        self.coef = torch.empty(5 * c, dtype=torch.float32, device=device)


def bn_train_fwd(x, gamma, beta, running_mean, running_var, sc: BNScratch, eps, momentum, res=None, relu=True,
                 stats_ready=False, num_batches_tracked=None, want_mask=False):
    """Returns (y, mean, rstd) or, with want_mask, (y, mean, rstd, mask): mask = uint8 [P*C/8], bit i of byte t =
    (y[8t + i] > 0) - the ReLU gate for bn_train_bwd when a residual was added before the ReLU."""
    c = x.shape[-1]
    p = x.numel() // c
    mean = torch.empty(c, dtype=torch.float32, device=x.device)
    rstd = torch.empty(c, dtype=torch.float32, device=x.device)
    y = torch.empty_like(x)
    lib = _lib.load()
    if want_mask:
        mask = torch.empty(p * c // 8, dtype=torch.uint8, device=x.device)
        _chk(lib.creamfl_bn_train_fwd_mask(_p(x), p, c, _p(gamma), _p(beta), eps, momentum, _p(running_mean),
                                           _p(running_var), _p(sc.sums), _p(mean), _p(rstd), _p(sc.scale), _p(sc.shift),
                                           _p(res), int(relu), int(stats_ready), _p(num_batches_tracked), _p(y), _p(mask),
                                           _stream()), "bn_train_fwd_mask", 2 if stats_ready else 3)
        return y, mean, rstd, mask
    _chk(lib.creamfl_bn_train_fwd(_p(x), p, c, _p(gamma), _p(beta), eps, momentum, _p(running_mean),
                                  _p(running_var), _p(sc.sums), _p(mean), _p(rstd), _p(sc.scale), _p(sc.shift),
                                  _p(res), int(relu), int(stats_ready), _p(num_batches_tracked), _p(y),
                                  _stream()), "bn_train_fwd",
         2 if stats_ready else 3)
    return y, mean, rstd


def bn_eval_fwd(x, gamma, beta, running_mean, running_var, sc: BNScratch, eps, res=None, relu=True):
    c = x.shape[-1]
    p = x.numel() // c
    y = torch.empty_like(x)
    _chk(_lib.load().creamfl_bn_eval_fwd(_p(x), p, c, _p(gamma), _p(beta), eps, _p(running_mean), _p(running_var),
                                         _p(sc.scale), _p(sc.shift), _p(res), int(relu), _p(y), _stream()),
         "bn_eval_fwd", 2)
    return y


def bn_train_bwd(dy, y_mask, x, gamma, mean, rstd, sc: BNScratch, dgamma, dbeta, want_g=False, beta=None,
                 relu_from_x=False, mask=None):
    """Returns (dx, g) where g = dy * gate (only if want_g).  gate = the bits of `mask` (from bn_train_fwd(want_mask))
    if given; (y_mask > 0) if y_mask is given; recomputed from x (no residual, ReLU applied) if relu_from_x and beta
    are given; 1 otherwise."""
    c = x.shape[-1]
    p = x.numel() // c
    dx = torch.empty_like(x)
    g = torch.empty_like(x) if want_g else None
    lib = _lib.load()
    if mask is not None:
        _chk(lib.creamfl_bn_train_bwd_mask(_p(dy), _p(mask), _p(x), p, c, _p(gamma), _p(mean), _p(rstd), _p(sc.sums),
                                           _p(sc.coef), _p(dgamma), _p(dbeta), _p(dx), _p(g), _stream()),
             "bn_train_bwd_mask", 3)
        return dx, g
    _chk(lib.creamfl_bn_train_bwd(_p(dy), _p(y_mask), _p(x), p, c, _p(gamma), _p(beta), int(relu_from_x),
                                  _p(mean), _p(rstd), _p(sc.sums), _p(sc.coef), _p(dgamma), _p(dbeta), _p(dx),
                                  _p(g), _stream()),
         "bn_train_bwd", 3)
    return dx, g


# --------------------------------------------------------------------------------------------------- pooling
def maxpool_fwd(x, want_idx=True):
    n, h, w, c = x.shape
    ho, wo = conv_out_hw(h, w, 3, 3, 2, 1)
    y = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device) if want_idx else None
    _chk(_lib.load().creamfl_maxpool_fwd(_p(x), n, h, w, c, _p(y), _p(idx), _stream()), "maxpool_fwd")
    return y, idx


def stem_tail_fwd(o, bn, training: bool, want_idx: bool):
    """Fused BatchNorm -> ReLU -> maxpool 3x3/2 of the stem's raw convolution output `o` [N, H, W, C]: the normalised
    map is never written.  bn: towers.BN.  Returns (pooled y, winning taps or None, (mean, rstd) or None)."""
    n, h, w, c = o.shape
    sc = bn.scratch()
    lib = _lib.load()
    saved = None
    if training:
        mean = torch.empty(c, dtype=torch.float32, device=o.device)
        rstd = torch.empty(c, dtype=torch.float32, device=o.device)
        _chk(lib.creamfl_bn_train_stats(_p(o), n * h * w, c, _p(bn.weight), _p(bn.bias), bn.eps, bn.momentum,
                                        _p(bn.running_mean), _p(bn.running_var), _p(sc.sums), _p(mean), _p(rstd),
                                        _p(sc.scale), _p(sc.shift), 0, _p(bn.num_batches_tracked), _stream()),
             "bn_train_stats", 2)
        saved = (mean, rstd)
    else:
        _chk(lib.creamfl_bn_eval_affine(c, _p(bn.weight), _p(bn.bias), bn.eps, _p(bn.running_mean), _p(bn.running_var),
                                        _p(sc.scale), _p(sc.shift), _stream()), "bn_eval_affine", 1)
    ho, wo = conv_out_hw(h, w, 3, 3, 2, 1)
    y = torch.empty((n, ho, wo, c), dtype=BF16, device=o.device)
    idx = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=o.device) if want_idx else None
    _chk(lib.creamfl_maxpool_affine_fwd(_p(o), _p(sc.scale), _p(sc.shift), n, h, w, c, _p(y), _p(idx), _stream()),
         "maxpool_affine_fwd")
    return y, idx, saved


def stem_tail_bwd(dy, idx, o, bn, saved, dgamma, dbeta):
    """Backward of stem_tail_fwd: d(raw convolution output) from the pooled gradient; the 112 x 112 gradient map of the
    pooling input is gathered on the fly inside the BatchNorm backward kernels."""
    n, h, w, c = o.shape
    mean, rstd = saved
    sc = bn.scratch()
    do = torch.empty_like(o)
    _chk(_lib.load().creamfl_bn_pool_bwd(_p(dy), _p(idx), _p(o), n, h, w, c, _p(bn.weight), _p(bn.bias), _p(mean),
                                         _p(rstd), _p(sc.sums), _p(sc.coef), _p(dgamma), _p(dbeta), _p(do), _stream()),
         "bn_pool_bwd", 3)
    return do


def maxpool_bwd(dy, idx, x_shape):
    n, h, w, c = x_shape
    dx = torch.empty((n, h, w, c), dtype=BF16, device=dy.device)
    _chk(_lib.load().creamfl_maxpool_bwd(_p(dy), _p(idx), n, h, w, c, _p(dx), _stream()), "maxpool_bwd")
    return dx


# --------------------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, eps, res=None, drop=None):
    """drop = (DropoutState, site): y = dropout(LayerNorm(x + res))."""
    r, d = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(r, dtype=torch.float32, device=x.device)
    rstd = torch.empty(r, dtype=torch.float32, device=x.device)
    if drop is None:
        _chk(_lib.load().creamfl_layernorm_fwd(_p(x), _p(res), _p(gamma), _p(beta), eps, r, d, int(x.dtype == BF16),
                                               _p(y), _p(mean), _p(rstd), _stream()), "layernorm_fwd")
    else:
        rng, site, pd = _drop_args(drop)
        _chk(_lib.load().creamfl_layernorm_fwd_drop(_p(x), _p(res), _p(gamma), _p(beta), eps, r, d,
                                                    int(x.dtype == BF16), _p(y), _p(mean), _p(rstd), rng, site, pd,
                                                    _stream()), "layernorm_fwd_drop")
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, res=None, dx_colsum=None, drop_in=None, drop_out=None):
    """drop_in: dy is the gradient of dropout(LayerNorm(.)) (masked on load).  drop_out: returns (dx, dx_drop) with
    dx_drop = dropout'(dx), the gradient of the dense layer feeding this LayerNorm; dx_colsum then sums dx_drop."""
    r, d = x.shape
    lib = _lib.load()
    dx = torch.empty_like(x)
    nb = lib.creamfl_layernorm_bwd_workspace_bytes(d)
    ws = torch.empty(nb, dtype=torch.uint8, device=x.device)
    if drop_in is None and drop_out is None:
        _chk(lib.creamfl_layernorm_bwd(_p(dy), _p(x), _p(res), _p(gamma), _p(mean), _p(rstd), r, d,
                                       int(x.dtype == BF16), _p(dx), _p(dgamma), _p(dbeta), _p(dx_colsum), _p(ws), nb,
                                       _stream()), "layernorm_bwd", 2)
        return dx
    state = (drop_in or drop_out)[0]
    dx_drop = torch.empty_like(x) if drop_out is not None else None
    _chk(lib.creamfl_layernorm_bwd_drop(_p(dy), _p(x), _p(res), _p(gamma), _p(mean), _p(rstd), r, d,
                                        int(x.dtype == BF16), _p(dx), _p(dx_drop), _p(dgamma), _p(dbeta),
                                        _p(dx_colsum), _p(ws), nb, _p(state.rng),
                                        drop_in[1] if drop_in is not None else -1,
                                        drop_out[1] if drop_out is not None else -1, float(state.p), _stream()),
         "layernorm_bwd_drop", 2)
    return (dx, dx_drop) if drop_out is not None else dx


def colsum_into(x, out):
    m, n = x.shape
    _chk(_lib.load().creamfl_colsum_bf16(_p(x), m, n, x.stride(0), _p(out), _stream()), "colsum_bf16")


def add_bf16(a, b):
    y = torch.empty_like(a)
    _chk(_lib.load().creamfl_add_bf16(_p(a), _p(b), a.numel(), _p(y), _stream()), "add_bf16")
    return y


def act_bwd(dy, y, kind):
    out = torch.empty(dy.shape, dtype=BF16, device=dy.device)
    _chk(_lib.load().creamfl_act_bwd_f32(_p(dy), _p(y), dy.numel(), kind, _p(out), _stream()), "act_bwd_f32")
    return out


# --------------------------------------------------------------------------------------------------- BERT pieces
def embed_fwd(ids, token_type, word, pos, typ, seq_len):
    t = ids.numel()
    d = word.shape[1]
    out = torch.empty((t, d), dtype=BF16, device=ids.device)
    _chk(_lib.load().creamfl_embed_fwd(_p(ids), _p(token_type), _p(word), _p(pos), _p(typ), t, seq_len, d, _p(out),
                                       _stream()), "embed_fwd")
    return out


def embed_bwd(ids, token_type, dh, seq_len, dword, dpos, dtyp):
    t, d = dh.shape
    _chk(_lib.load().creamfl_embed_bwd(_p(ids), _p(token_type), _p(dh), t, seq_len, d, _p(dword), _p(dpos), _p(dtyp),
                                       _stream()), "embed_bwd")


def attn_fwd(qkv, mask, b, l, heads, drop=None):
    """drop = (DropoutState, site): ctx = dropout(softmax(...)) v; `probs` keeps the un-dropped probabilities."""
    ctx = torch.empty((b * l, heads * 64), dtype=BF16, device=qkv.device)
    probs = torch.empty((b, heads, l, l), dtype=BF16, device=qkv.device)
    if drop is None:
        _chk(_lib.load().creamfl_attn_fwd(_p(qkv), _p(mask), b, l, heads, 64, _p(ctx), _p(probs), _stream()),
             "attn_fwd")
    else:
        rng, site, pd = _drop_args(drop)
        _chk(_lib.load().creamfl_attn_fwd_drop(_p(qkv), _p(mask), b, l, heads, 64, _p(ctx), _p(probs), rng, site, pd,
                                               _stream()), "attn_fwd_drop")
    return ctx, probs


def attn_bwd(qkv, probs, dctx, b, l, heads, dbias=None, drop=None):
    dqkv = torch.empty_like(qkv)
    if drop is None:
        _chk(_lib.load().creamfl_attn_bwd(_p(qkv), _p(probs), _p(dctx), b, l, heads, 64, _p(dqkv), _p(dbias),
                                          _stream()), "attn_bwd")
    else:
        rng, site, pd = _drop_args(drop)
        _chk(_lib.load().creamfl_attn_bwd_drop(_p(qkv), _p(probs), _p(dctx), b, l, heads, 64, _p(dqkv), _p(dbias), rng,
                                               site, pd, _stream()), "attn_bwd_drop")
    return dqkv


# --------------------------------------------------------------------------------------------------- PIE pooling
def pie_pool_fwd(x, h, w2):
    b, p, c = x.shape
    hd = h.shape[-1]
    attn = torch.empty((b, p), dtype=torch.float32, device=x.device)
    r = torch.empty((b, c), dtype=BF16, device=x.device)
    pooled = torch.empty((b, c), dtype=BF16, device=x.device)
    _chk(_lib.load().creamfl_pie_pool_fwd(_p(x), _p(h), _p(w2), b, p, c, hd, _p(attn), _p(r), _p(pooled), _stream()),
         "pie_pool_fwd")
    return attn, r, pooled


def pie_pool_bwd(x, h, w2, attn, d_r, d_pooled, dw2):
    b, p, c = x.shape
    hd = h.shape[-1]
    dx = torch.empty_like(x)
    dpre = torch.empty_like(h)
    _chk(_lib.load().creamfl_pie_pool_bwd(_p(x), _p(h), _p(w2), _p(attn), _p(d_r), _p(d_pooled), b, p, c, hd, _p(dx),
                                          _p(dpre), _p(dw2), _stream()), "pie_pool_bwd")
    return dx, dpre


# --------------------------------------------------------------------------------------------------- unimodal heads
def avgpool_fwd(x, scale=1.0):
    """x [N, H, W, C] bf16 -> (fp32 [N, C], bf16 [N, C]) = scale * mean over H*W."""
    n, h, w, c = x.shape
    y = torch.empty((n, c), dtype=torch.float32, device=x.device)
    y16 = torch.empty((n, c), dtype=BF16, device=x.device)
    _chk(_lib.load().creamfl_avgpool_fwd(_p(x), n, h * w, c, float(scale), _p(y), _p(y16), _stream()), "avgpool_fwd")
    return y, y16


def avgpool_bwd(dy16, shape, scale=1.0):
    n, h, w, c = shape
    dx = torch.empty(shape, dtype=BF16, device=dy16.device)
    _chk(_lib.load().creamfl_avgpool_bwd(_p(dy16), n, h * w, c, float(scale), _p(dx), _stream()), "avgpool_bwd")
    return dx


def relu_inplace(master, shadow):
    _chk(_lib.load().creamfl_relu_inplace(_p(master), _p(shadow), master.numel(), _stream()), "relu_inplace")


# --------------------------------------------------------------------------------------------------- GRU text towers
def pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def wemb_gather(ids, table, pitch):
    """ids int64 [T] -> bf16 [T, pitch]: rows of the fp32 table, zero tail (pitch % 8 == 0 for TMA)."""
    _need_cuda(ids, table)
    t = ids.numel()
    v, dw = table.shape
    out = torch.empty((t, pitch), dtype=BF16, device=ids.device)
    _chk(_lib.load().creamfl_wemb_gather_fwd(_p(ids), _p(table), t, v, dw, pitch, _p(out), _stream()),
         "wemb_gather_fwd")
    return out


def wemb_scatter(ids, dx16, dtable):
    """dtable[ids[t], :] += dx16[t, :Dw]  (dx16 bf16 [T, pitch])."""
    if dx16.dtype != BF16:
        raise TypeError('wemb_scatter: dx must be bf16')
    t = ids.numel()
    v, dw = dtable.shape
    _chk(_lib.load().creamfl_wemb_scatter_bwd(_p(ids), _p(dx16), t, v, dw, dx16.stride(0), _p(dtable), _stream()),
         "wemb_scatter_bwd")


def gru_fwd(xproj, w_hh, b_hh, lengths32, b, l, h, rev_steps=0, want_seq=True, want_last=True, want_gates=True):
    """xproj fp32 [B*L, 6H]; w_hh fp32 [2, 3H, H] (contiguous), b_hh fp32 [2*3H]; lengths int32 [B]."""
    _need_cuda(xproj, w_hh, b_hh, lengths32)
    dev = xproj.device
    hseq = torch.empty((b, l, 2 * h), dtype=torch.float32, device=dev) if want_seq else None
    hlast = torch.empty((b, 2 * h), dtype=torch.float32, device=dev) if want_last else None
    gates = torch.empty((b, l, 2, 4, h), dtype=torch.float32, device=dev) if want_gates else None
    _chk(_lib.load().creamfl_gru_fwd(_p(xproj), _p(w_hh), _p(b_hh), _p(lengths32), b, l, h, int(rev_steps), _p(hseq),
                                     _p(hlast), _p(gates), _stream()), "gru_fwd")
    return hseq, hlast, gates


def gru_bwd(gates, hseq, w_hh, lengths32, dhseq, dhlast, b, l, h, rev_steps=0):
    dev = gates.device
    dxp = torch.empty((b * l, 6 * h), dtype=BF16, device=dev)
    dgh = torch.empty((b * l, 6 * h), dtype=BF16, device=dev)
    hprev = torch.empty((b * l, 2 * h), dtype=BF16, device=dev)
    _chk(_lib.load().creamfl_gru_bwd(_p(gates), _p(hseq), _p(w_hh), _p(lengths32), _p(dhseq), _p(dhlast), b, l, h,
                                     int(rev_steps), _p(dxp), _p(dgh), _p(hprev), _stream()), "gru_bwd")
    return dxp, dgh, hprev


def seq_pool_fwd(x, hid, w2, lengths32, c, hd):
    """x bf16 [B, L, pitch] (c valid columns), hid bf16 [B, L, hpitch] (hd valid), w2 fp32 [hd]."""
    b, l, pitch = x.shape
    attn = torch.empty((b, l), dtype=torch.float32, device=x.device)
    r = torch.empty((b, pitch), dtype=BF16, device=x.device)
    _chk(_lib.load().creamfl_seq_pool_fwd(_p(x), _p(hid), _p(w2), _p(lengths32), b, l, c, pitch, hd, hid.shape[-1],
                                          _p(attn), _p(r), _stream()), "seq_pool_fwd")
    return attn, r


def seq_pool_bwd(x, hid, w2, attn, d_r, lengths32, c, hd, dw2):
    b, l, pitch = x.shape
    dx = torch.empty_like(x)
    dpre = torch.empty_like(hid)
    _chk(_lib.load().creamfl_seq_pool_bwd(_p(x), _p(hid), _p(w2), _p(attn), _p(d_r), _p(lengths32), b, l, c, pitch, hd,
                                          hid.shape[-1], _p(dx), _p(dpre), _p(dw2), _stream()), "seq_pool_bwd")
    return dx, dpre


def scale_relu_fwd(x, scale):
    y = torch.empty_like(x)
    _chk(_lib.load().creamfl_scale_relu_fwd(_p(x), x.numel(), float(scale), _p(y), _stream()), "scale_relu_fwd")
    return y


def scale_relu_bwd(dy, y, scale):
    dx = torch.empty_like(y)
    _chk(_lib.load().creamfl_scale_relu_bwd(_p(dy), _p(y), y.numel(), float(scale), _p(dx), _stream()),
         "scale_relu_bwd")
    return dx
