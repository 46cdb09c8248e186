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

BF16 = torch.bfloat16        # storage type of the "bf16" tensors; install(exact=True) switches it to fp32
EXACT = False


def gemm_bf16(a, b, *, a_mn=False, b_mn=False, bias=None, act=0, alpha=1.0, add=None, aux=None,
              out_dtype=None, want_preact=False, split_k=1, out=None, accumulate=False, n_cols=None):
    if out_dtype is None or out_dtype == torch.bfloat16:
        out_dtype = BF16
    assert a.dtype == BF16 and b.dtype == BF16 and a.stride(1) == 1 and b.stride(1) == 1
    if not EXACT:      # TMA operand rules: 16-byte aligned base and row pitch
        assert a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
        assert (a.stride(0) * 2) % 16 == 0 and (b.stride(0) * 2) % 16 == 0
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
    if act == 1:
        acc = F.gelu(acc)                                   # erf GELU
    elif act == 2:
        acc = torch.relu(acc)
    elif act == 3:
        acc = torch.tanh(acc)
    elif act == 4:                                          # acc * gelu'(aux)
        xa = aux.float()
        acc = acc * (0.5 * (1 + torch.erf(xa / 2 ** 0.5)) + xa * torch.exp(-0.5 * xa * xa) / (2 * torch.pi) ** 0.5)
    elif act == 5:                                          # acc * (aux > 0)
        acc = acc * (aux.float() > 0)
    elif act == 6:
        acc = torch.sigmoid(acc)
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


# ---- dropout (include/creamfl_b200.h "BERT dropout"): Philox4x32-10 restated in numpy, independent of csrc/philox.cuh
def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): uint32 arrays in, four uint32 arrays out."""
    import numpy as np
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    m32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m32
        n1 = p1 & m32
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m32
        n3 = p0 & m32
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & m32
        k1 = (k1 + np.uint64(0xBB67AE85)) & m32
    return c0, c1, c2, c3


def keep_mask_np(seed, step, site, n, p):
    """uint8 [n] keep mask of a dropout site: element e is kept iff 16-bit field (e & 7) of Philox(counter
    {e >> 3, site, step}, key seed) >= round(p * 65536)."""
    import numpy as np
    blocks = (n + 7) // 8
    e8 = np.arange(blocks, dtype=np.uint64)
    r = philox4x32_10(e8 & np.uint64(0xFFFFFFFF), e8 >> np.uint64(32), np.full(blocks, site, dtype=np.uint64),
                      np.full(blocks, step & 0xFFFFFFFF, dtype=np.uint64), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    thresh = np.uint64(int(p * 65536.0 + 0.5))
    fields = np.stack([f for w in r for f in (w & np.uint64(0xFFFF), w >> np.uint64(16))], axis=1)   # [blocks, 8]
    return (fields >= thresh).astype(np.uint8).reshape(-1)[:n]


def _dropstate_tick(self):
    self.rng[1] += 1


def _dropstate_keep_mask(self, site, n):
    seed, step = (int(v) for v in self.rng.tolist())
    return torch.from_numpy(keep_mask_np(seed, step, int(site), int(n), self.p))


def _mask_scale(drop, shape):
    state, site = drop
    n = 1
    for d in shape:
        n *= int(d)
    return _dropstate_keep_mask(state, site, n).view(*shape).float() / (1.0 - state.p)


def gemm_drop(a, b, bias, add, drop, out=None):
    acc = a.float() @ b.float().t()
    if bias is not None:
        acc = acc + bias.float()
    if drop is not None:
        acc = acc * _mask_scale(drop, acc.shape)
    if add is not None:
        acc = acc + add.float()
    if out is None:
        return acc.to(BF16)
    out.copy_(acc.to(out.dtype))
    return out


def layernorm_fwd(x, gamma, beta, eps, res=None, drop=None):
    s = x.float() + (res.float() if res is not None else 0)
    mean = s.mean(-1)
    var = s.var(-1, unbiased=False)
    rstd = (var + eps).rsqrt()
    y = (s - mean[:, None]) * rstd[:, None] * gamma + beta
    if drop is not None:
        y = y * _mask_scale(drop, y.shape)
    return y.to(x.dtype), mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, res=None, dx_colsum=None, drop_in=None, drop_out=None):
    s = x.float() + (res.float() if res is not None else 0)
    xhat = (s - mean[:, None]) * rstd[:, None]
    dy = dy.float()
    if drop_in is not None:
        dy = dy * _mask_scale(drop_in, dy.shape)
    dgamma += (dy * xhat).sum(0)
    dbeta += dy.sum(0)
    g = dy * gamma
    dx = rstd[:, None] * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    dxd = dx * _mask_scale(drop_out, dx.shape) if drop_out is not None else dx
    if dx_colsum is not None:
        dx_colsum += dxd.sum(0)
    if drop_out is not None:
        return dx.to(x.dtype), dxd.to(x.dtype)
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




# ---- image tower / BERT entry points (tower_ops)
def _nchw(x):
    return x.float().permute(0, 3, 1, 2)


def _filters(w2d, cout, r, s_, cin):
    return w2d[:, :r * s_ * cin].float().reshape(cout, r, s_, cin).permute(0, 3, 1, 2)


def conv_out_hw(h, w, r, s_, stride, pad):
    return (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s_) // stride + 1


def conv_fprop(x, w2d, r, s_, stride, pad, bn_sums=None):
    n, h, w, cin = x.shape
    cout = w2d.shape[0]
    y = F.conv2d(_nchw(x), _filters(w2d, cout, r, s_, cin), stride=stride, padding=pad).permute(0, 2, 3, 1)
    y = y.contiguous().to(BF16)
    if bn_sums is not None:                     # statistics of the bf16 output, accumulated
        y2 = y.double().reshape(-1, cout)
        bn_sums[:cout] += y2.sum(0)
        bn_sums[cout:] += (y2 * y2).sum(0)
    return y


def conv_dgrad(dy, w2d, x_shape, r, s_, stride, pad, add=None):
    n, h, w, cin = x_shape
    cout = w2d.shape[0]
    dx = torch.nn.grad.conv2d_input((n, cin, h, w), _filters(w2d, cout, r, s_, cin), _nchw(dy), stride=stride,
                                    padding=pad).permute(0, 2, 3, 1)
    if add is not None:
        dx = dx + add.float()
    return dx.contiguous().to(BF16)


def conv_wgrad(dy, x, dw, r, s_, stride, pad):
    n, h, w, cin = x.shape
    cout = dy.shape[-1]
    gw = torch.nn.grad.conv2d_weight(_nchw(x), (cout, cin, r, s_), _nchw(dy), stride=stride, padding=pad)
    dw += gw.permute(0, 2, 3, 1).reshape(cout, r * s_ * cin)


def im2col_images(images, r, s_, stride, pad, pitch):
    n, c, h, w = images.shape
    cols = F.unfold(images.float(), (r, s_), padding=pad, stride=stride)           # [N, C*r*s, L], (c, r, s) order
    l = cols.shape[-1]
    cols = cols.view(n, c, r * s_, l).permute(0, 3, 2, 1).reshape(n * l, r * s_ * c)
    out = torch.zeros((n * l, pitch), dtype=BF16)
    out[:, :r * s_ * c] = cols.to(BF16)
    return out


def stem_supported(h, w):
    return w <= 256 and w % 4 == 0 and (w + 6 - 7) // 2 + 1 <= 128


def stem_fprop(images, w16):
    """conv 7x7 / stride 2 / pad 3 from fp32 NCHW images; w16 [64, >= 147] in (r, s, c) column order."""
    wt = w16[:, :147].float().view(64, 7, 7, 3).permute(0, 3, 1, 2)
    x = images.to(BF16).float()                     # the kernel rounds the patch operand to bf16
    return F.conv2d(x, wt, stride=2, padding=3).permute(0, 2, 3, 1).contiguous().to(BF16)


def stem_wgrad(images, dy, dw):
    x = images.to(BF16).float()
    cols = F.unfold(x, (7, 7), padding=3, stride=2)                                   # [N, 3*49, L] in (c, r, s) order
    n, _, l = cols.shape
    cols = cols.view(n, 3, 49, l).permute(0, 3, 2, 1).reshape(n * l, 147)             # (r, s, c) order
    dw += dy.float().reshape(-1, 64).t() @ cols


def bn_train_fwd(x, gamma, beta, running_mean, running_var, sc, eps, momentum, res=None, relu=True, stats_ready=False,
                 num_batches_tracked=None, want_mask=False):
    c = x.shape[-1]
    x2 = x.double().reshape(-1, c)
    p = x2.shape[0]
    if not stats_ready:
        sc.sums[:c] += x2.sum(0)
        sc.sums[c:] += (x2 * x2).sum(0)
    m = sc.sums[:c] / p
    var = (sc.sums[c:] / p - m * m).clamp_min(0)
    sc.sums.zero_()
    rstd = (1.0 / torch.sqrt(var + eps)).float()
    mean = m.float()
    running_mean.mul_(1 - momentum).add_(momentum * mean)
    running_var.mul_(1 - momentum).add_(momentum * (var * p / max(p - 1, 1)).float())
    if num_batches_tracked is not None:
        num_batches_tracked += 1
    scale = gamma * rstd
    y = x.float() * scale + (beta - mean * scale)
    if res is not None:
        y = y + res.float()
    if want_mask:                                # gate bits of the pre-ReLU value, 8 channels per byte, LSB first
        bits = (y > 0).reshape(-1, 8).to(torch.uint8)
        mask = (bits << torch.arange(8, dtype=torch.uint8)).sum(1).to(torch.uint8)
    if relu:
        y = torch.relu(y)
    return (y.to(BF16), mean, rstd, mask) if want_mask else (y.to(BF16), mean, rstd)


def bn_eval_fwd(x, gamma, beta, running_mean, running_var, sc, eps, res=None, relu=True):
    scale = gamma * torch.rsqrt(running_var + eps)
    y = x.float() * scale + (beta - running_mean * scale)
    if res is not None:
        y = y + res.float()
    if relu:
        y = torch.relu(y)
    return y.to(BF16)


def stem_tail_fwd(o, bn, training, want_idx):
    """creamfl_bn_train_stats / creamfl_bn_eval_affine + creamfl_maxpool_affine_fwd: maxpool(relu(BatchNorm(o)))."""
    c = o.shape[-1]
    if training:
        a, mean, rstd = bn_train_fwd(o, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.scratch(), bn.eps,
                                     bn.momentum, relu=True, num_batches_tracked=bn.num_batches_tracked)
        saved = (mean, rstd)
    else:
        a, saved = bn_eval_fwd(o, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.scratch(), bn.eps, relu=True), None
    y, idx = maxpool_fwd(a, want_idx=want_idx)
    return y, idx, saved


def stem_tail_bwd(dy, idx, o, bn, saved, dgamma, dbeta):
    """creamfl_bn_pool_bwd: BatchNorm backward (ReLU gate from x) of the max-pooling backward of dy."""
    mean, rstd = saved
    da = maxpool_bwd(dy, idx, o.shape)
    do, _ = bn_train_bwd(da, None, o, bn.weight, mean, rstd, bn.scratch(), dgamma, dbeta, beta=bn.bias, relu_from_x=True)
    return do


def bn_fold_layers(layers, row_start, total_rows, eps, pairs=None):
    """creamfl_bn_fold_layers: w_out[c, :] = w[c, :] * s_c, bias[c] = beta[c] - mean[c] * s_c (the table rows hold the
    addresses of exactly these tensors; the emulation walks the modules behind them)."""
    assert layers.shape == (len(pairs), 10) and int(row_start[-1]) == total_rows
    with torch.no_grad():
        for conv, bn, wv, bv in pairs:
            o = conv.weight.shape[0]
            s = bn.weight * torch.rsqrt(bn.running_var + eps)
            wv.copy_((conv.weight.permute(0, 2, 3, 1).reshape(o, -1) * s[:, None]).to(wv.dtype))
            bv.copy_(bn.bias - bn.running_mean * s)


def conv_fprop_affine(x, w2d, r, s_, stride, pad, bias, add=None, relu=True):
    n, h, w, cin = x.shape
    cout = w2d.shape[0]
    y = F.conv2d(_nchw(x), _filters(w2d, cout, r, s_, cin), stride=stride, padding=pad).permute(0, 2, 3, 1).float() + bias
    if add is not None:
        y = y + add.float()
    if relu:
        y = torch.relu(y)
    return y.to(BF16)


def bn_train_bwd(dy, y_mask, x, gamma, mean, rstd, sc, dgamma, dbeta, want_g=False, beta=None, relu_from_x=False,
                 mask=None):
    c = x.shape[-1]
    xhat = (x.float() - mean) * rstd
    g = dy.float()
    if mask is not None:
        gate = ((mask[:, None] >> torch.arange(8, dtype=torch.uint8)) & 1).reshape(x.shape)
        g = g * gate
    elif y_mask is not None:
        g = g * (y_mask.float() > 0)
    elif relu_from_x and beta is not None:
        g = g * ((gamma * xhat + beta) > 0)
    g2, xh2 = g.reshape(-1, c), xhat.reshape(-1, c)
    p = g2.shape[0]
    s1, s2 = g2.sum(0), (g2 * xh2).sum(0)
    dbeta += s1
    dgamma += s2
    dx = gamma * rstd * (g - s1 / p - xhat * (s2 / p))
    return dx.to(BF16), (g.to(BF16) if want_g else None)


def maxpool_fwd(x, want_idx=True):
    y, idx = F.max_pool2d(_nchw(x), 3, 2, 1, return_indices=True)
    return y.permute(0, 2, 3, 1).contiguous().to(BF16), (idx if want_idx else None)     # idx: opaque to the caller


def maxpool_bwd(dy, idx, x_shape):
    n, h, w, c = x_shape
    dx = torch.zeros(n, c, h * w)
    dx.scatter_add_(2, idx.flatten(2), _nchw(dy).flatten(2))          # overlapping windows: contributions add up
    return dx.view(n, c, h, w).permute(0, 2, 3, 1).contiguous().to(BF16)


def embed_fwd(ids, token_type, word, pos, typ, seq_len):
    t = ids.numel()
    posidx = torch.arange(t) % seq_len
    tt = token_type.reshape(-1) if token_type is not None else torch.zeros(t, dtype=torch.long)
    return (word[ids.reshape(-1)] + pos[posidx] + typ[tt]).to(BF16)


def embed_bwd(ids, token_type, dh, seq_len, dword, dpos, dtyp):
    t = dh.shape[0]
    d = dh.float()
    dword.index_add_(0, ids.reshape(-1), d)
    dpos.index_add_(0, torch.arange(t) % seq_len, d)
    tt = token_type.reshape(-1) if token_type is not None else torch.zeros(t, dtype=torch.long)
    dtyp.index_add_(0, tt, d)


def _split_heads(qkv, b, l, heads):
    q, k, v = qkv.float().view(b, l, 3, heads, 64).permute(2, 0, 3, 1, 4)                # each [B, H, L, 64]
    return q, k, v


def attn_fwd(qkv, mask, b, l, heads, drop=None):
    q, k, v = _split_heads(qkv, b, l, heads)
    s = q @ k.transpose(-1, -2) * 0.125
    s = s + torch.where(mask > 0.5, 0.0, -3.0e38)[:, None, None, :]                     # HF extended mask
    p = torch.softmax(s, dim=-1).to(BF16)          # rounded once: the kernel multiplies the bf16 probabilities
    pd = p.float() * _mask_scale(drop, p.shape) if drop is not None else p.float()
    ctx = (pd @ v).permute(0, 2, 1, 3).reshape(b * l, heads * 64)
    return ctx.to(BF16), p


def attn_bwd(qkv, probs, dctx, b, l, heads, dbias=None, drop=None):
    q, k, v = _split_heads(qkv, b, l, heads)
    p = probs.float()
    do = dctx.float().view(b, l, heads, 64).permute(0, 2, 1, 3)
    dp = do @ v.transpose(-1, -2)
    pd = p
    if drop is not None:
        m = _mask_scale(drop, p.shape)
        dp, pd = dp * m, p * m
    ds = p * (dp - (dp * p).sum(-1, keepdim=True)) * 0.125
    dq, dk, dv = ds @ k, ds.transpose(-1, -2) @ q, pd.transpose(-1, -2) @ do
    dqkv = torch.stack([dq, dk, dv], 0).permute(1, 3, 0, 2, 4).reshape(b * l, 3 * heads * 64)
    if dbias is not None:
        dbias += dqkv.sum(0)
    return dqkv.to(BF16)


def pie_pool_fwd(x, h, w2):
    a = torch.softmax(h.float() @ w2, dim=1)                                            # [B, P]
    r = torch.einsum('bp,bpc->bc', a, x.float())
    return a, r.to(BF16), x.float().mean(1).to(BF16)


def pie_pool_bwd(x, h, w2, attn, d_r, d_pooled, dw2):
    xf, hf, dr = x.float(), h.float(), d_r.float()
    p = x.shape[1]
    dx = attn[:, :, None] * dr[:, None, :] + d_pooled.float()[:, None, :] / p
    dattn = torch.einsum('bpc,bc->bp', xf, dr)
    da = attn * (dattn - (attn * dattn).sum(1, keepdim=True))
    dpre = da[:, :, None] * w2[None, None, :] * (1 - hf * hf)
    dw2 += torch.einsum('bp,bph->h', da, hf)
    return dx.to(BF16), dpre.to(BF16)


def avgpool_fwd(x, scale=1.0):
    n, h, w, c = x.shape
    y = x.float().reshape(n, h * w, c).mean(1) * scale
    return y, y.to(BF16)


def avgpool_bwd(dy16, shape, scale=1.0):
    n, h, w, c = shape
    return (dy16.float()[:, None, None, :] * (scale / (h * w))).expand(n, h, w, c).contiguous().to(BF16)


def l2norm_raw(x32):
    inv = 1.0 / x32.norm(dim=-1).clamp_min(1e-12)
    return x32 * inv[:, None], inv


def l2norm_bwd_raw(gy, y, inv):
    return inv[:, None] * (gy - y * (gy * y).sum(-1, keepdim=True))


def cross_entropy(x, labels, margin=0.0):
    lab = torch.arange(x.shape[0]) if labels is None else labels
    return F.cross_entropy(x - margin * F.one_hot(lab, x.shape[1]).to(x.dtype), lab)


_OPS = ('gemm_bf16', 'cast_into', 'to_bf16', 'l2_normalize', 'l2norm_raw', 'l2norm_bwd_raw', 'cross_entropy')
_TOWER_OPS = ('wemb_gather', 'wemb_scatter', 'gru_fwd', 'gru_bwd', 'seq_pool_fwd', 'seq_pool_bwd', 'scale_relu_fwd',
              'scale_relu_bwd', 'layernorm_fwd', 'layernorm_bwd', 'act_bwd', 'colsum_into', 'relu_inplace', 'conv_fprop',
              'conv_dgrad', 'conv_wgrad', 'im2col_images', 'bn_train_fwd', 'bn_eval_fwd', 'bn_train_bwd', 'maxpool_fwd',
              'maxpool_bwd', 'embed_fwd', 'embed_bwd', 'attn_fwd', 'attn_bwd', 'pie_pool_fwd', 'pie_pool_bwd',
              'avgpool_fwd', 'avgpool_bwd', 'gemm_drop', 'stem_supported', 'stem_fprop', 'stem_wgrad', 'bn_fold_layers',
              'conv_fprop_affine', 'stem_tail_fwd', 'stem_tail_bwd')


def install(monkeypatch, exact=False):
    """Swap the ctypes wrappers used by creamfl_b200.{towers,text_towers,clients} for the emulations above and lift
    the CUDA-only guard of the ParamStore (tests only).  exact=True additionally stores every "bf16" tensor in fp32,
    which turns the comparison with the fp32 oracle into a sharp check of the host-side sequencing (1e-4 instead of
    the bf16 noise floor)."""
    import sys
    from creamfl_b200 import ops, tower_ops, towers
    me = sys.modules[__name__]
    monkeypatch.setattr(me, 'BF16', torch.float32 if exact else torch.bfloat16)
    monkeypatch.setattr(me, 'EXACT', bool(exact))
    if exact:
        for mod in (towers, tower_ops):
            monkeypatch.setattr(mod, 'BF16', torch.float32)
    monkeypatch.setattr(towers, '_require_cuda', lambda dev: None)
    for name in _OPS:
        monkeypatch.setattr(ops, name, getattr(me, name))
    for name in _TOWER_OPS:
        monkeypatch.setattr(tower_ops, name, getattr(me, name))
    monkeypatch.setattr(tower_ops.DropoutState, 'tick', _dropstate_tick)
    monkeypatch.setattr(tower_ops.DropoutState, 'keep_mask', _dropstate_keep_mask)


# ================================================================================================ engine level
# Loss / aggregation / metric / optimizer entry points, emulated with the CPU oracle (oracle/creamfl_oracle.py - the
# restatement of the reference's arithmetic that the CUDA kernels are tested against on the GPU).  Used by
# tests/test_cpu_round_plumbing.py to run the product's orchestrator (src/algorithms) for one communication round on
# the CPU: BASELINE.json configs[0] "plumbing, no GPU".
def _O():
    from oracle import creamfl_oracle
    return creamfl_oracle


def _ste_round(x):
    """bf16 operand rounding with a straight-through gradient (the kernels round the operand, not the gradient)."""
    return x + (x.detach().to(BF16).to(x.dtype) - x.detach())


def pcme_loss(img, txt, shift, neg_scale):
    O = _O()
    i2t = O.pcme_direction_loss(img.float(), txt.float(), shift.float().reshape(()), neg_scale.float().reshape(()))
    t2i = O.pcme_direction_loss(txt.float(), img.float(), shift.float().reshape(()), neg_scale.float().reshape(()))
    loss = i2t['loss'] + t2i['loss']
    parts = torch.stack([loss.detach(), i2t['pos_loss'].detach(), i2t['neg_loss'].detach()]).float()
    return loss, parts


def infonce_loss(q, bank_bf16, labels, inv_tau=2.0):
    return _O().inter_infonce(_ste_round(q.float()), bank_bf16.float(), labels, tau=1.0 / inv_tau)


def moon_intra_loss(z, zold, bank, idx, inv_tau=2.0, denom=None):
    return _O().moon_intra(z.float(), zold.float(), bank[idx].float(), tau=1.0 / inv_tau, denom=denom)


def mse_gather_loss(x, bank, idx):
    return _O().distill_mse(x.float(), bank, idx)


def conw_score(v_bf16, g_bf16):
    return _O().conw_scores(v_bf16.float(), g_bf16.float())


def conw_reduce(vecs, scores, want_weights=False):
    w = torch.softmax(scores.float(), dim=0)
    out = sum(v.float() * w[c].reshape(-1, 1) for c, v in enumerate(vecs))
    return (out, w) if want_weights else out


def conw_aggregate(vecs, global_other, want_weights=False):
    g = global_other.to(BF16).float()
    scores = torch.stack([conw_score(v.to(BF16), g) for v in vecs], dim=0)
    return conw_reduce(vecs, scores, want_weights)


def recall_ranks(q, g, q_labels, g_labels):
    return torch.from_numpy(_O().recall_ranks_count(q.float(), g.float(), q_labels.numpy(), g_labels.numpy())).to(torch.int32)


class _EmuOptimizer:
    """Drop-in bodies for creamfl_b200.optim.FusedOptimizer's device-table methods: global-norm clipping + AdamP /
    Adam / SGD-momentum through the oracle's fp64 restatement, then the bf16 shadow refresh."""

    @staticmethod
    def _build(self):
        params = [p for g in self.param_groups for p in g['params']]
        self._loose = []
        for p in params:
            if hasattr(p, '_g2d'):
                if p.grad is None:
                    p.grad = p._gview
            else:
                gbuf = torch.zeros_like(p.data)
                if p.grad is not None:
                    gbuf.copy_(p.grad)
                p.grad = gbuf
                self._loose.append((p, gbuf))
        self._emu = {'step': 0, 'm': [torch.zeros(p.shape, dtype=torch.float64) for p in params],
                     'v': [torch.zeros(p.shape, dtype=torch.float64) for p in params], 'norm': torch.zeros(())}
        self._built = True

    @staticmethod
    def _sync_hyper(self):
        pass

    @staticmethod
    def step(self, closure=None):
        O = _O()
        self.prepare()
        params = [p for g in self.param_groups for p in g['params']]
        hp = self.param_groups[0]
        grads = [(p.grad if p.grad is not None else torch.zeros_like(p.data)).detach().double().clone() for p in params]
        if hp['max_norm'] > 0:
            self._emu['norm'] = torch.tensor(O.clip_grad_norm(
                [g for p, g in zip(params, grads) if id(p) not in self._no_clip], hp['max_norm']))
        pd = [p.data.double() for p in params]
        st = self._emu
        st['step'] += 1
        if self.mode == 'sgd':
            O.sgd_momentum_step(pd, grads, st['m'], st['step'], hp['lr'], hp['momentum'], hp['weight_decay'])
        else:
            O.adamp_step(pd, grads, st['m'], st['v'], st['step'], hp['lr'], betas=hp['betas'], eps=hp['eps'],
                         weight_decay=hp['weight_decay'], delta=hp['delta'] if self.mode == 'adamp' else -1.0,
                         wd_ratio=hp['wd_ratio'])
        with torch.no_grad():
            for p, new in zip(params, pd):
                p.data.copy_(new.float())
        for store in self._stores():
            store.sync_shadow()


def install_engine(monkeypatch, exact=False):
    """install() plus the engine-level entry points and a CPU device for the engines (tests only)."""
    import sys
    install(monkeypatch, exact=exact)
    from creamfl_b200 import engine, ops, optim
    me = sys.modules[__name__]
    for name in ('pcme_loss', 'infonce_loss', 'moon_intra_loss', 'mse_gather_loss', 'conw_score', 'conw_reduce',
                 'conw_aggregate', 'recall_ranks'):
        monkeypatch.setattr(ops, name, getattr(me, name))
    monkeypatch.setattr(engine, 'default_device', lambda index=None: torch.device('cpu'))
    for name in ('_build', '_sync_hyper', 'step'):
        monkeypatch.setattr(optim.FusedOptimizer, name, getattr(_EmuOptimizer, name))
    monkeypatch.setattr(optim.FusedOptimizer, 'grad_norm', property(lambda self: self._emu['norm']))
