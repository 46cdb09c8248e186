"""PyTorch-facing wrappers of the creamfl_b200 C ABI.

PyTorch supplies device memory (caching allocator), the current stream and autograd bookkeeping; every
floating-point operation of the wrapped ops happens inside libcreamfl_b200.so.  All tensors must live on a CUDA
device - there is no CPU path (calling with CPU tensors raises).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH, ACT_DGELU, ACT_DRELU, ACT_SIGMOID = range(7)

# kernel launches issued through this module since the last reset (bench.py reports it as gpu_launches)
_launches = 0
_LAUNCHES_PER_CALL = {
    "gemm": 1, "infonce_fwd": 3, "infonce_bwd": 4, "conw_score": 2, "conw_reduce": 1, "pcme_fwd": 2,
    "pcme_bwd": 2, "moon_fwd": 2, "moon_bwd": 1, "mse_fwd": 2, "mse_bwd": 1, "l2norm_fwd": 1, "l2norm_bwd": 1,
    "cast": 1, "recall": 4,
}


def launches() -> int:
    return _launches


def reset_launches() -> None:
    global _launches
    _launches = 0


def _count(kind: str) -> None:
    global _launches
    _launches += _LAUNCHES_PER_CALL[kind]


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("creamfl_b200 ops run on CUDA tensors only (no CPU fallback exists)")


def _contig(t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    if t.dtype != dtype:
        raise TypeError(f"expected {dtype}, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# --------------------------------------------------------------------------------------------------- casts
def to_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy through the library's cast kernel (bf16 input is returned unchanged)."""
    if x.dtype == torch.bfloat16:
        return x if x.is_contiguous() else x.contiguous()
    _need_cuda(x)
    x = _contig(x, torch.float32)
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().creamfl_cast_f32_bf16(_p(x), x.numel(), _p(y), _stream()), "cast_f32_bf16")
    _count("cast")
    return y


def cast_into(src_f32: torch.Tensor, dst_bf16: torch.Tensor) -> None:
    """dst (bf16, same numel) <- src (fp32), both contiguous 1-D views; used to refresh parameter shadows."""
    _need_cuda(src_f32, dst_bf16)
    if src_f32.numel() != dst_bf16.numel() or src_f32.dtype != torch.float32 or dst_bf16.dtype != torch.bfloat16:
        raise ValueError("cast_into: shape / dtype mismatch")
    _lib.check(_lib.load().creamfl_cast_f32_bf16(_p(src_f32), src_f32.numel(), _p(dst_bf16), _stream()),
               "cast_f32_bf16")
    _count("cast")


def l2norm_raw(x32: torch.Tensor):
    """(y, inv_norm) of F.normalize over the last dim of an fp32 [R, D] tensor; no autograd."""
    R, D = x32.shape
    y = torch.empty_like(x32)
    inv = torch.empty(R, dtype=torch.float32, device=x32.device)
    _lib.check(_lib.load().creamfl_l2norm_fwd(_p(x32), R, D, _p(y), None, _p(inv), _stream()), "l2norm_fwd")
    _count("l2norm_fwd")
    return y, inv


def l2norm_bwd_raw(gy: torch.Tensor, y: torch.Tensor, inv: torch.Tensor) -> torch.Tensor:
    R, D = y.shape
    dx = torch.empty_like(y)
    _lib.check(_lib.load().creamfl_l2norm_bwd(_p(gy), _p(y), _p(inv), R, D, _p(dx), _stream()), "l2norm_bwd")
    _count("l2norm_bwd")
    return dx


# --------------------------------------------------------------------------------------------------- GEMM
def gemm_bf16(a: torch.Tensor, b: torch.Tensor, *, a_mn: bool = False, b_mn: bool = False,
              bias: Optional[torch.Tensor] = None, act: int = ACT_NONE, alpha: float = 1.0,
              add: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None,
              out_dtype: torch.dtype = torch.bfloat16, want_preact: bool = False, split_k: int = 1,
              out: Optional[torch.Tensor] = None, accumulate: bool = False, n_cols: Optional[int] = None):
    """out[M,N] = act(alpha * A B^T + bias + add) on tcgen05.

    a: [M,K] (or [K,M] if a_mn), b: [N,K] (or [K,N] if b_mn); both bf16, 2-D, unit inner stride.
    """
    _need_cuda(a, b, bias, add, aux, out)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise TypeError("gemm_bf16 operands must be bf16")
    if a.dim() != 2 or b.dim() != 2 or a.stride(1) != 1 or b.stride(1) != 1:
        raise ValueError("gemm_bf16 operands must be 2-D with unit inner stride")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb:
        raise ValueError(f"gemm_bf16: K mismatch {K} vs {Kb}")
    if n_cols is not None:
        if n_cols > N:
            raise ValueError("gemm_bf16: n_cols exceeds the operand")
        N = n_cols
    if out is None:
        if split_k > 1:
            out = torch.zeros((M, N), dtype=torch.float32, device=a.device)
        else:
            out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    pre = torch.empty((M, N), dtype=torch.bfloat16, device=a.device) if want_preact else None
    if bias is not None:
        bias = _contig(bias, torch.float32)
    add_bf16 = 0
    if add is not None:
        add_bf16 = 1 if add.dtype == torch.bfloat16 else 0
        if add.stride(1) != 1:
            add = add.contiguous()
    if aux is not None:
        aux = _contig(aux, torch.bfloat16)
    rc = _lib.load().creamfl_gemm_bf16(
        _p(a), a.stride(0), int(a_mn), _p(b), b.stride(0), int(b_mn), M, N, K, _p(out), out.stride(0),
        1 if out.dtype == torch.bfloat16 else 0, _p(pre), _p(bias), act, float(alpha), _p(add),
        add.stride(0) if add is not None else 0, add_bf16, _p(aux), aux.stride(0) if aux is not None else 0,
        int(split_k), int(accumulate), _stream())
    _lib.check(rc, "gemm_bf16")
    _count("gemm")
    return (out, pre) if want_preact else out


# --------------------------------------------------------------------------------------------------- InfoNCE
class _InfoNCE(torch.autograd.Function):
    """CE(inv_tau * Q G^T, labels), mean over rows; G (the public-feature bank) carries no gradient.

    Reference: MMClientTrainer.py:193-201,301-308; ClientTrainer.py:388-401,493-502.
    """

    @staticmethod
    def forward(ctx, q, bank_bf16, labels, inv_tau):
        _need_cuda(q, bank_bf16, labels)
        lib = _lib.load()
        qb = to_bf16(q.detach())
        B, D = qb.shape
        N = bank_bf16.shape[0]
        labels = _contig(labels, torch.int64)
        loss = torch.empty(1, dtype=torch.float32, device=q.device)
        row = torch.empty(B, dtype=torch.float32, device=q.device)
        lse2 = torch.empty(B, dtype=torch.float32, device=q.device)
        nbytes = lib.creamfl_rowlse_workspace_bytes(B, N)
        ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=q.device)
        _lib.check(lib.creamfl_infonce_fwd(_p(qb), _p(bank_bf16), _p(labels), B, N, D, float(inv_tau), _p(loss),
                                           _p(row), _p(lse2), _p(ws), nbytes, _stream()), "infonce_fwd")
        _count("infonce_fwd")
        ctx.save_for_backward(qb, bank_bf16, labels, lse2)
        ctx.inv_tau = float(inv_tau)
        ctx.q_dtype = q.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        qb, bank, labels, lse2 = ctx.saved_tensors
        lib = _lib.load()
        B, D = qb.shape
        N = bank.shape[0]
        g = gout.detach().to(torch.float32).reshape(1).contiguous()
        dq = torch.empty((B, D), dtype=torch.float32, device=qb.device)
        nbytes = lib.creamfl_infonce_bwd_workspace_bytes(B, N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=qb.device)
        _lib.check(lib.creamfl_infonce_bwd(_p(qb), _p(bank), _p(labels), _p(lse2), B, N, D, ctx.inv_tau, _p(g),
                                           _p(dq), _p(ws), nbytes, _stream()), "infonce_bwd")
        _count("infonce_bwd")
        return dq.to(ctx.q_dtype), None, None, None


def infonce_loss(q: torch.Tensor, bank_bf16: torch.Tensor, labels: torch.Tensor, inv_tau: float = 2.0):
    if bank_bf16.dtype != torch.bfloat16:
        raise TypeError("the public-feature bank must be bf16 (use to_bf16 once per round)")
    return _InfoNCE.apply(q, bank_bf16, labels, inv_tau)


# --------------------------------------------------------------------------------------------------- con_w
def conw_score(v_bf16: torch.Tensor, g_bf16: torch.Tensor) -> torch.Tensor:
    """score[n] = <V[n],G[n]> - log sum_j exp <V[n],G[j]>  (MMFL.py:304-307)."""
    _need_cuda(v_bf16, g_bf16)
    lib = _lib.load()
    N, D = v_bf16.shape
    if g_bf16.shape != v_bf16.shape:
        raise ValueError("conw_score: client and global representations must have the same shape")
    score = torch.empty(N, dtype=torch.float32, device=v_bf16.device)
    nbytes = lib.creamfl_rowlse_workspace_bytes(N, N)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=v_bf16.device)
    _lib.check(lib.creamfl_conw_score(_p(_contig(v_bf16, torch.bfloat16)), _p(_contig(g_bf16, torch.bfloat16)),
                                      N, D, _p(score), _p(ws), nbytes, _stream()), "conw_score")
    _count("conw_score")
    return score


def conw_reduce(vecs: Sequence[torch.Tensor], scores: torch.Tensor, want_weights: bool = False):
    """out[n] = sum_c softmax_c(scores[:, n])[c] * vecs[c][n]  (MMFL.py:311-314)."""
    _need_cuda(scores, *vecs)
    lib = _lib.load()
    Cn = len(vecs)
    vecs = [_contig(v, torch.float32) for v in vecs]
    N, D = vecs[0].shape
    scores = _contig(scores, torch.float32)
    if scores.shape != (Cn, N):
        raise ValueError(f"conw_reduce: scores must be [{Cn}, {N}]")
    out = torch.empty((N, D), dtype=torch.float32, device=scores.device)
    w = torch.empty((Cn, N), dtype=torch.float32, device=scores.device) if want_weights else None
    arr = (C.c_void_p * Cn)(*[v.data_ptr() for v in vecs])
    _lib.check(lib.creamfl_conw_reduce(arr, _p(scores), Cn, N, D, _p(out), _p(w), _stream()), "conw_reduce")
    _count("conw_reduce")
    return (out, w) if want_weights else out


def conw_aggregate(vecs: Sequence[torch.Tensor], global_other: torch.Tensor, want_weights: bool = False):
    """Full con_w aggregation of one modality: score every client against the server's opposite-modality
    features, softmax over clients, weighted sum (MMFL.py:298-335)."""
    g = to_bf16(global_other)
    scores = torch.stack([conw_score(to_bf16(v), g) for v in vecs], dim=0)
    return conw_reduce(vecs, scores, want_weights)


# --------------------------------------------------------------------------------------------------- PCME
class _PCME(torch.autograd.Function):
    """MCSoftContrastiveLoss.forward (src/criterions/probemb.py:221-256), both directions, reduction='sum'."""

    @staticmethod
    def forward(ctx, img, txt, shift, neg_scale):
        _need_cuda(img, txt, shift, neg_scale)
        lib = _lib.load()
        img32 = _contig(img.detach().float(), torch.float32)
        txt32 = _contig(txt.detach().float(), torch.float32)
        if img32.shape != txt32.shape:
            raise RuntimeError(f"# anchors ({tuple(img32.shape)}) != # candidates ({tuple(txt32.shape)})")
        N, D = img32.shape
        sh = _contig(shift.detach().float().reshape(1), torch.float32)
        ns = _contig(neg_scale.detach().float().reshape(1), torch.float32)
        dist = torch.empty((N, N), dtype=torch.float32, device=img.device)
        out3 = torch.empty(3, dtype=torch.float32, device=img.device)
        nbytes = lib.creamfl_pcme_workspace_bytes(N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=img.device)
        _lib.check(lib.creamfl_pcme_fwd(_p(img32), _p(txt32), N, D, _p(sh), _p(ns), _p(dist), _p(out3), _p(ws),
                                        nbytes, _stream()), "pcme_fwd")
        _count("pcme_fwd")
        ctx.save_for_backward(img32, txt32, dist, sh, ns)
        ctx.dtypes = (img.dtype, txt.dtype, shift.dtype, neg_scale.dtype)
        ctx.mark_non_differentiable(out3)
        return out3[0].clone(), out3

    @staticmethod
    def backward(ctx, gout, _g3):
        img32, txt32, dist, sh, ns = ctx.saved_tensors
        lib = _lib.load()
        N, D = img32.shape
        g = gout.detach().float().reshape(1).contiguous()
        d_img = torch.empty_like(img32)
        d_txt = torch.empty_like(txt32)
        d_sh = torch.empty(1, dtype=torch.float32, device=img32.device)
        d_ns = torch.empty(1, dtype=torch.float32, device=img32.device)
        nbytes = lib.creamfl_pcme_workspace_bytes(N)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=img32.device)
        _lib.check(lib.creamfl_pcme_bwd(_p(img32), _p(txt32), _p(dist), N, D, _p(sh), _p(ns), _p(g), _p(d_img),
                                        _p(d_txt), _p(d_sh), _p(d_ns), _p(ws), nbytes, _stream()), "pcme_bwd")
        _count("pcme_bwd")
        t = ctx.dtypes
        return d_img.to(t[0]), d_txt.to(t[1]), d_sh.to(t[2]), d_ns.to(t[3])


def pcme_loss(img, txt, shift, neg_scale):
    """Returns (loss, parts) with parts = [loss, per-direction positive part, per-direction negative part]."""
    return _PCME.apply(img, txt, shift, neg_scale)


# --------------------------------------------------------------------------------------------------- MOON
class _Moon(torch.autograd.Function):
    """sum_r CE([<z,bank[idx]>, <z,zold>] * inv_tau, 0) / denom  (MMClientTrainer.py:169-191)."""

    @staticmethod
    def forward(ctx, z, zold, bank, idx, inv_tau, denom):
        _need_cuda(z, zold, bank, idx)
        lib = _lib.load()
        z32 = _contig(z.detach().float(), torch.float32)
        zo = _contig(zold.detach().float(), torch.float32)
        bank = _contig(bank, torch.float32)
        idx = _contig(idx, torch.int64)
        R, D = z32.shape
        rows = torch.empty(R, dtype=torch.float32, device=z.device)
        coef = torch.empty(R, dtype=torch.float32, device=z.device)
        loss = torch.empty(1, dtype=torch.float32, device=z.device)
        _lib.check(lib.creamfl_moon_fwd(_p(z32), _p(zo), _p(bank), _p(idx), R, D, float(inv_tau), float(denom),
                                        _p(rows), _p(coef), _p(loss), _stream()), "moon_fwd")
        _count("moon_fwd")
        ctx.save_for_backward(zo, bank, idx, coef)
        ctx.z_dtype = z.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        zo, bank, idx, coef = ctx.saved_tensors
        lib = _lib.load()
        R, D = zo.shape
        g = gout.detach().float().reshape(1).contiguous()
        dz = torch.empty_like(zo)
        _lib.check(lib.creamfl_moon_bwd(_p(zo), _p(bank), _p(idx), _p(coef), _p(g), R, D, _p(dz), _stream()),
                   "moon_bwd")
        _count("moon_bwd")
        return dz.to(ctx.z_dtype), None, None, None, None, None


def moon_intra_loss(z, zold, bank, idx, inv_tau: float = 2.0, denom: Optional[float] = None):
    return _Moon.apply(z, zold, bank, idx, inv_tau, float(z.shape[0] if denom is None else denom))


# --------------------------------------------------------------------------------------------------- distill MSE
class _MseGather(torch.autograd.Function):
    """nn.MSELoss()(x, bank[idx])  (MMFL.py:296,355-378)."""

    @staticmethod
    def forward(ctx, x, bank, idx):
        _need_cuda(x, bank, idx)
        lib = _lib.load()
        x32 = _contig(x.detach().float(), torch.float32)
        bank = _contig(bank, torch.float32)
        idx = _contig(idx, torch.int64)
        R, D = x32.shape
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        nbytes = lib.creamfl_mse_workspace_bytes()
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.creamfl_mse_gather_fwd(_p(x32), _p(bank), _p(idx), R, D, _p(loss), _p(ws), nbytes,
                                              _stream()), "mse_gather_fwd")
        _count("mse_fwd")
        ctx.save_for_backward(x32, bank, idx)
        ctx.x_dtype = x.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        x32, bank, idx = ctx.saved_tensors
        lib = _lib.load()
        R, D = x32.shape
        g = gout.detach().float().reshape(1).contiguous()
        dx = torch.empty_like(x32)
        _lib.check(lib.creamfl_mse_gather_bwd(_p(x32), _p(bank), _p(idx), _p(g), R, D, _p(dx), _stream()),
                   "mse_gather_bwd")
        _count("mse_bwd")
        return dx.to(ctx.x_dtype), None, None


def mse_gather_loss(x, bank, idx):
    return _MseGather.apply(x, bank, idx)


# --------------------------------------------------------------------------------------------------- cross-entropy
class _CrossEntropy(torch.autograd.Function):
    """nn.CrossEntropyLoss()(x - margin * onehot(labels), labels)  (ClientTrainer.py:346-352; labels None -> arange,
    the class-centre loss on W W^T)."""

    @staticmethod
    def forward(ctx, x, labels, margin):
        _need_cuda(x, labels)
        x32 = x.detach().float()
        if x32.stride(1) != 1:
            x32 = x32.contiguous()
        R, Cn = x32.shape
        if labels is not None:
            labels = _contig(labels, torch.int64)
        rows = torch.empty(R, dtype=torch.float32, device=x.device)
        dlog = torch.empty((R, Cn), dtype=torch.float32, device=x.device)
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().creamfl_ce_fwd(_p(x32), x32.stride(0), _p(labels), R, Cn, float(margin), _p(rows),
                                              _p(dlog), _p(loss), _stream()), "ce_fwd")
        global _launches
        _launches += 2
        ctx.save_for_backward(dlog)
        ctx.x_dtype = x.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        (dlog,) = ctx.saved_tensors
        return (dlog * gout).to(ctx.x_dtype), None, None


def cross_entropy(x: torch.Tensor, labels: Optional[torch.Tensor], margin: float = 0.0) -> torch.Tensor:
    return _CrossEntropy.apply(x, labels, margin)


# --------------------------------------------------------------------------------------------------- L2 norm
class _L2Norm(torch.autograd.Function):
    """F.normalize(x, p=2, dim=-1)  (src/utils/tensor_utils.py:25-27)."""

    @staticmethod
    def forward(ctx, x):
        _need_cuda(x)
        lib = _lib.load()
        x32 = _contig(x.detach().float(), torch.float32)
        R, D = x32.shape
        y = torch.empty_like(x32)
        inv = torch.empty(R, dtype=torch.float32, device=x.device)
        _lib.check(lib.creamfl_l2norm_fwd(_p(x32), R, D, _p(y), None, _p(inv), _stream()), "l2norm_fwd")
        _count("l2norm_fwd")
        ctx.save_for_backward(y, inv)
        ctx.x_dtype = x.dtype
        return y

    @staticmethod
    def backward(ctx, gy):
        y, inv = ctx.saved_tensors
        lib = _lib.load()
        R, D = y.shape
        gy = _contig(gy.detach().float(), torch.float32)
        dx = torch.empty_like(y)
        _lib.check(lib.creamfl_l2norm_bwd(_p(gy), _p(y), _p(inv), R, D, _p(dx), _stream()), "l2norm_bwd")
        _count("l2norm_bwd")
        return dx.to(ctx.x_dtype)


def l2_normalize(x: torch.Tensor) -> torch.Tensor:
    if x.dim() != 2:
        raise ValueError("l2_normalize expects [rows, features]")
    return _L2Norm.apply(x)


# --------------------------------------------------------------------------------------------------- Recall@K
def recall_ranks(q: torch.Tensor, g: torch.Tensor, q_labels: torch.Tensor, g_labels: torch.Tensor) -> torch.Tensor:
    """0-based rank of the best positive of every query (eval_coco.py:273-334), int32 [Nq]."""
    _need_cuda(q, g, q_labels, g_labels)
    lib = _lib.load()
    q = _contig(q.float(), torch.float32)
    g = _contig(g.float(), torch.float32)
    q_labels = _contig(q_labels.to(torch.int64), torch.int64)
    g_labels = _contig(g_labels.to(torch.int64), torch.int64)
    Nq, D = q.shape
    Ng = g.shape[0]
    if len(q_labels) != Nq or len(g_labels) != Ng:
        raise RuntimeError(f"length mismatch {tuple(q.shape)}, {tuple(q_labels.shape)}")
    ranks = torch.empty(Nq, dtype=torch.int32, device=q.device)
    nbytes = lib.creamfl_recall_workspace_bytes(Nq)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    _lib.check(lib.creamfl_recall_ranks(_p(q), _p(g), _p(q_labels), _p(g_labels), Nq, Ng, D, _p(ranks), _p(ws),
                                        nbytes, _stream()), "recall_ranks")
    _count("recall")
    return ranks
