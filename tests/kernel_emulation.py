"""Torch (CPU) emulation of the libcreamfl_b200 entry points the GRU text towers call - TEST INFRASTRUCTURE ONLY.

The build container has no GPU; the host-side sequencing of the text towers (creamfl_b200/text_towers.py: operand
layouts, padded pitches, fused gradient views, accumulation targets) is nevertheless checkable here by swapping the
ctypes wrappers for functions that follow the C ABI's documented semantics (include/creamfl_b200.h) and, for the
GRU, the per-step arithmetic of csrc/text_ops.cu line by line (same gate order, same saved quantities, same bf16
roundings on the outputs).  tests/test_cpu_text_tower_host.py installs it with monkeypatch; nothing in the product
imports this file.  The CUDA kernels themselves are tested on the GPU (tests/test_gpu_text_tower.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def gemm_bf16(a, b, *, a_mn=False, b_mn=False, bias=None, act=0, alpha=1.0, add=None, aux=None,
              out_dtype=BF16, want_preact=False, split_k=1, out=None, accumulate=False, n_cols=None):
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(1) == 1 and b.stride(1) == 1
    assert a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0 and (a.stride(0) * 2) % 16 == 0 and (b.stride(0) * 2) % 16 == 0
    A = a.float().t() if a_mn else a.float()
    Bm = b.float() if b_mn else b.float().t()
    assert A.shape[1] == Bm.shape[0], (A.shape, Bm.shape)
    acc = alpha * (A @ Bm)
    if n_cols is not None:
        acc = acc[:, :n_cols]
    if bias is not None:
        assert bias.numel() == acc.shape[1]
        acc = acc + bias.float()
    if add is not None:
        assert add.shape == acc.shape
        acc = acc + add.float()
    pre = acc.to(BF16) if want_preact else None
    if act == 3:
        acc = torch.tanh(acc)
    elif act == 6:
        acc = torch.sigmoid(acc)
    elif act == 2:
        acc = torch.relu(acc)
    elif act != 0:
        raise NotImplementedError(act)
    if out is None:
        out = acc.to(torch.float32 if split_k > 1 else out_dtype)
    else:
        assert out.shape == acc.shape, (out.shape, acc.shape)
        if accumulate or split_k > 1:
            assert out.dtype == torch.float32 and act == 0
            out += acc
        else:
            out.copy_(acc.to(out.dtype))
    return (out, pre) if want_preact else out


def cast_into(src, dst):
    dst.copy_(src.to(BF16))


def to_bf16(x):
    return x.to(BF16).contiguous()


def l2_normalize(x):
    return F.normalize(x, p=2, dim=-1)


# ---- tower_ops
def pad8(n):
    return (n + 7) // 8 * 8


def wemb_gather(ids, table, pitch):
    out = torch.zeros((ids.numel(), pitch), dtype=BF16)
    out[:, :table.shape[1]] = table[ids].to(BF16)
    return out


def wemb_scatter(ids, dx16, dtable):
    assert dx16.dtype == BF16
    dtable.index_add_(0, ids, dx16[:, :dtable.shape[1]].float())


def gru_fwd(xproj, w_hh, b_hh, lengths32, b, l, h, rev_steps=0, want_seq=True, want_last=True, want_gates=True):
    xp = xproj.view(b, l, 2, 3 * h)
    hseq = torch.zeros(b, l, 2 * h)
    hlast = torch.zeros(b, 2 * h)
    gates = torch.zeros(b, l, 2, 4, h)
    bh = b_hh.view(2, 3 * h)
    for bi in range(b):
        ln = int(lengths32[bi])
        for d in range(2):
            nst = rev_steps if (d == 1 and 0 < rev_steps < ln) else ln
            hs = torch.zeros(h)
            for s in range(nst):
                t = s if d == 0 else ln - 1 - s
                gh = w_hh[d] @ hs + bh[d]
                x = xp[bi, t, d]
                r = torch.sigmoid(x[:h] + gh[:h])
                z = torch.sigmoid(x[h:2 * h] + gh[h:2 * h])
                n = torch.tanh(x[2 * h:] + r * gh[2 * h:])
                hs = (1 - z) * n + z * hs
                hseq[bi, t, d * h:(d + 1) * h] = hs
                gates[bi, t, d, 0], gates[bi, t, d, 1], gates[bi, t, d, 2], gates[bi, t, d, 3] = r, z, n, gh[2 * h:]
                if t == ln - 1:
                    hlast[bi, d * h:(d + 1) * h] = hs
    return (hseq if want_seq else None), (hlast if want_last else None), (gates if want_gates else None)


def gru_bwd(gates, hseq, w_hh, lengths32, dhseq, dhlast, b, l, h, rev_steps=0):
    dxp = torch.zeros(b, l, 2, 3 * h)
    dgh = torch.zeros(b, l, 2, 3 * h)
    hprev = torch.zeros(b, l, 2, h)
    for bi in range(b):
        ln = int(lengths32[bi])
        for d in range(2):
            nst = rev_steps if (d == 1 and 0 < rev_steps < ln) else ln
            dh = torch.zeros(h)
            for s in range(nst - 1, -1, -1):
                t = s if d == 0 else ln - 1 - s
                if dhseq is not None:
                    dh = dh + dhseq[bi, t, d * h:(d + 1) * h]
                if dhlast is not None and t == ln - 1:
                    dh = dh + dhlast[bi, d * h:(d + 1) * h]
                r, z, n, hn = gates[bi, t, d]
                hp = torch.zeros(h)
                if s > 0:
                    tp = t - 1 if d == 0 else t + 1
                    hp = hseq[bi, tp, d * h:(d + 1) * h]
                dn = dh * (1 - z)
                dz = dh * (hp - n)
                dnp = dn * (1 - n * n)
                dzp = dz * z * (1 - z)
                drp = dnp * hn * r * (1 - r)
                dxp[bi, t, d] = torch.cat([drp, dzp, dnp])
                v = torch.cat([drp, dzp, dnp * r])
                dgh[bi, t, d] = v
                hprev[bi, t, d] = hp
                dh = dh * z + w_hh[d].t() @ v
    return dxp.view(b * l, 6 * h).to(BF16), dgh.view(b * l, 6 * h).to(BF16), hprev.view(b * l, 2 * h).to(BF16)


def seq_pool_fwd(x, hid, w2, lengths32, c, hd):
    b, l, pitch = x.shape
    attn = torch.zeros(b, l)
    r = torch.zeros(b, pitch)
    for bi in range(b):
        ln = int(lengths32[bi])
        a = hid[bi, :ln, :hd].float() @ w2
        a = torch.softmax(a, dim=0)
        attn[bi, :ln] = a
        r[bi, :c] = a @ x[bi, :ln, :c].float()
    return attn, r.to(BF16)


def seq_pool_bwd(x, hid, w2, attn, d_r, lengths32, c, hd, dw2):
    b, l, pitch = x.shape
    dx = torch.zeros(b, l, pitch)
    dpre = torch.zeros(hid.shape)
    for bi in range(b):
        ln = int(lengths32[bi])
        a = attn[bi]
        dr = d_r[bi, :c].float()
        dattn = x[bi, :, :c].float() @ dr
        dattn[ln:] = 0
        s = (a * dattn).sum()
        da = a * (dattn - s)
        dx[bi, :, :c] = a[:, None] * dr[None, :]
        hv = hid[bi, :, :hd].float()
        dpre[bi, :, :hd] = da[:, None] * w2[None, :] * (1 - hv * hv)
        dw2 += da @ hv
    return dx.to(BF16), dpre.to(BF16)


def scale_relu_fwd(x, scale):
    return torch.relu(x * scale)


def scale_relu_bwd(dy, y, scale):
    return torch.where(y > 0, dy * scale, torch.zeros_like(dy))


def layernorm_fwd(x, gamma, beta, eps, res=None):
    s = x.float() + (res.float() if res is not None else 0)
    mean = s.mean(-1)
    var = s.var(-1, unbiased=False)
    rstd = (var + eps).rsqrt()
    y = (s - mean[:, None]) * rstd[:, None] * gamma + beta
    return y.to(x.dtype), mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, res=None, dx_colsum=None):
    s = x.float() + (res.float() if res is not None else 0)
    xhat = (s - mean[:, None]) * rstd[:, None]
    dy = dy.float()
    dgamma += (dy * xhat).sum(0)
    dbeta += dy.sum(0)
    g = dy * gamma
    dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    if dx_colsum is not None:
        dx_colsum += dx.sum(0)
    return dx.to(x.dtype)


def act_bwd(dy, y, kind):
    if kind == 6:
        return (dy * y * (1 - y)).to(BF16)
    if kind == 3:
        return (dy * (1 - y * y)).to(BF16)
    raise NotImplementedError(kind)


def colsum_into(x, out):
    assert x.shape[1] % 8 == 0 and x.stride(0) % 8 == 0
    out += x.float().sum(0)


def relu_inplace(master, shadow):
    master.clamp_(min=0)
    if shadow is not None:
        shadow.copy_(master.to(BF16))


def install(monkeypatch):
    """Swap the ctypes wrappers used by creamfl_b200.text_towers / clients._LinearFn for the emulations above and
    lift the CUDA-only guard of the ParamStore (tests only)."""
    from creamfl_b200 import ops, tower_ops, towers
    monkeypatch.setattr(towers, '_require_cuda', lambda dev: None)
    for name in ('gemm_bf16', 'cast_into', 'to_bf16', 'l2_normalize'):
        monkeypatch.setattr(ops, name, globals()[name])
    for name in ('wemb_gather', 'wemb_scatter', 'gru_fwd', 'gru_bwd', 'seq_pool_fwd', 'seq_pool_bwd', 'scale_relu_fwd',
                 'scale_relu_bwd', 'layernorm_fwd', 'layernorm_bwd', 'act_bwd', 'colsum_into', 'relu_inplace'):
        monkeypatch.setattr(tower_ops, name, globals()[name])
