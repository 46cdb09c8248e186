"""GRU text towers of the clients on the creamfl_b200 kernels.

Mirrors
  * caption_encoder.EncoderText (src/networks/models/caption_encoder.py:29-116) - the text tower of the multimodal
    client's PCME (`config.not_bert = True`, forced at MMFL.py:163): Embedding -> packed bi-GRU -> last valid step ->
    PIENet(word embeddings, pad mask) -> LayerNorm -> l2_normalize;
  * language_model.EncoderText (src/networks/language_model.py:28-130) - the unimodal text client: same trunk, then
    `* scale`, ReLU and ReLU-clamped classifier heads (training) or the L2-normalised embedding.
Parameter names and shapes equal torch's (`embed.weight`, `rnn.weight_ih_l0`, `rnn.weight_hh_l0_reverse`, ...,
`pie_net.attention.w_1.weight`, ...), so reference checkpoints load unchanged.

Kernel sequence (forward): wemb_gather -> ONE tcgen05 GEMM x W_ih^T + b_ih for both directions (K = 300 padded to
304) -> gru_fwd (register-resident W_hh, sequential part only) -> GEMM + tanh (PIENet w_1) -> seq_pool_fwd (masked
softmax pooling) -> GEMM + sigmoid (PIENet fc) -> LayerNorm(hlast + residual).  The reverse direction runs ONE step:
the tower consumes rnn_out[b, len_b - 1] only (caption_encoder.py:99-101), which the reverse direction produces in
its first step - the remaining len_b - 1 reverse steps of the reference never reach the output or any gradient.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn

from . import ops, tower_ops as T, towers as _towers
from .towers import PIENet, ParamStore, StoreMixin, _Linear, grad_target


class _GRUParams(nn.Module):
    """Parameter container with nn.GRU(input, hidden, bidirectional=True) names, shapes and default init."""

    def __init__(self, input_size: int, hidden_size: int):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        k = 1.0 / math.sqrt(hidden_size)
        for sfx in ('', '_reverse'):
            for name, shape in (('weight_ih_l0', (3 * hidden_size, input_size)),
                                ('weight_hh_l0', (3 * hidden_size, hidden_size)),
                                ('bias_ih_l0', (3 * hidden_size,)), ('bias_hh_l0', (3 * hidden_size,))):
                p = nn.Parameter(torch.empty(shape))
                nn.init.uniform_(p, -k, k)
                setattr(self, name + sfx, p)

    def groups(self) -> List[List[nn.Parameter]]:
        return [[self.weight_ih_l0, self.weight_ih_l0_reverse], [self.weight_hh_l0, self.weight_hh_l0_reverse],
                [self.bias_ih_l0, self.bias_ih_l0_reverse], [self.bias_hh_l0, self.bias_hh_l0_reverse]]

    def flatten_parameters(self) -> None:   # nn.GRU API; the flat ParamStore already is one buffer
        pass


class _TextTowerFn(torch.autograd.Function):
    """ids [B, L] int64, lengths int32 [B] -> LayerNorm(gru_last + sigmoid(fc(attention-pool(word embeddings))))
    fp32 [B, D] (caption_encoder.py:90-106 before l2_normalize)."""

    @staticmethod
    def forward(ctx, ids, len32, tower, *params):
        tw = tower
        b, l = ids.shape
        t = b * l
        h, dw, hd = tw.hidden, tw.word_dim, tw.word_dim // 2
        kp = T.pad8(dw)
        need = any(ctx.needs_input_grad)
        pie = tw.pie_net
        idsf = ids.reshape(-1).contiguous()
        x16 = T.wemb_gather(idsf, tw.embed.weight.data, kp)                                    # [T, kp]
        xproj = ops.gemm_bf16(x16, tw._wih16, bias=tw._bih, out_dtype=torch.float32)           # [T, 2*3H]
        hseq, hlast, gates = T.gru_fwd(xproj, tw._whh, tw._bhh, len32, b, l, h, rev_steps=1, want_seq=need,
                                       want_last=True, want_gates=need)
        hid = ops.gemm_bf16(x16, pie.attention.w_1.weight._w16p, act=ops.ACT_TANH)             # [T, pad8(hd)]
        w2 = pie.attention.w_2.weight.data.view(-1)
        attn, r16 = T.seq_pool_fwd(x16.view(b, l, kp), hid.view(b, l, -1), w2, len32, dw, hd)
        res = ops.gemm_bf16(r16, pie.fc.weight._w16, bias=pie.fc.bias, act=ops.ACT_SIGMOID, out_dtype=torch.float32)
        ln = pie.layer_norm
        z, mean, rstd = T.layernorm_fwd(hlast, ln.weight, ln.bias, ln.eps, res=res)
        ctx.tower = tw
        ctx.saved = (idsf, len32, x16, hseq, hlast, gates, hid, attn, r16, res, mean, rstd, b, l) if need else None
        return z

    @staticmethod
    def backward(ctx, dz):
        tw = ctx.tower
        idsf, len32, x16, hseq, hlast, gates, hid, attn, r16, res, mean, rstd, b, l = ctx.saved
        ctx.saved = None
        t = b * l
        h, dw, hd = tw.hidden, tw.word_dim, tw.word_dim // 2
        kp = T.pad8(dw)
        pie = tw.pie_net
        ln = pie.layer_norm
        for p_ in tw.tower_params():
            grad_target(p_)
        dsum = T.layernorm_bwd(dz.contiguous().float(), hlast, ln.weight, mean, rstd, grad_target(ln.weight),
                               grad_target(ln.bias), res=res)                                   # d(hlast) = d(res)
        # PIENet residual: res = sigmoid(r fc^T + b)
        d_respre = T.act_bwd(dsum, res, ops.ACT_SIGMOID)
        ops.gemm_bf16(d_respre, r16, a_mn=True, b_mn=True, out=grad_target(pie.fc.weight), split_k=0,
                      accumulate=True, n_cols=dw)
        T.colsum_into(d_respre, grad_target(pie.fc.bias))
        d_r = ops.gemm_bf16(d_respre, pie.fc.weight._w16, b_mn=True)                            # [B, kp]
        w2 = pie.attention.w_2.weight.data.view(-1)
        dx_part, dpre = T.seq_pool_bwd(x16.view(b, l, kp), hid.view(b, l, -1), w2, attn, d_r, len32, dw, hd,
                                       grad_target(pie.attention.w_2.weight).view(-1))
        dpre2 = dpre.view(t, -1)
        ops.gemm_bf16(dpre2[:, :hd], x16, a_mn=True, b_mn=True, out=grad_target(pie.attention.w_1.weight), split_k=0,
                      accumulate=True, n_cols=dw)
        dx1 = ops.gemm_bf16(dpre2, pie.attention.w_1.weight._w16p, b_mn=True, add=dx_part.view(t, kp))
        # GRU: only rnn_out[b, len_b - 1] is consumed, so d(hlast) is the whole upstream gradient
        dxp, dgh, hprev = T.gru_bwd(gates, hseq, tw._whh, len32, None, dsum, b, l, h, rev_steps=1)
        ops.gemm_bf16(dxp, x16, a_mn=True, b_mn=True, out=tw._gwih, split_k=0, accumulate=True, n_cols=dw)
        T.colsum_into(dxp, tw._gbih)
        T.colsum_into(dgh, tw._gbhh)
        for d_ in range(2):
            ops.gemm_bf16(dgh[:, d_ * 3 * h:(d_ + 1) * 3 * h], hprev[:, d_ * h:(d_ + 1) * h], a_mn=True, b_mn=True,
                          out=tw._gwhh[d_], split_k=0, accumulate=True)
        dx = ops.gemm_bf16(dxp, tw._wih16, b_mn=True, add=dx1)                                  # bf16 [T, kp]
        T.wemb_scatter(idsf, dx, grad_target(tw.embed.weight))
        return (None, None, None) + (None,) * (len(ctx.needs_input_grad) - 3)


class _ScaleReluFn(torch.autograd.Function):
    """relu(x * scale) (language_model.py:111-112)."""

    @staticmethod
    def forward(ctx, x, scale):
        y = T.scale_relu_fwd(x.contiguous().float(), scale)
        ctx.save_for_backward(y)
        ctx.scale = scale
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return T.scale_relu_bwd(dy.contiguous().float(), y, ctx.scale), None


class _TextTrunk(nn.Module):
    """Embedding + bi-GRU + PIENet parameters and the store bindings shared by both text towers."""

    def __init__(self, vocab_size: int, word_dim: int, embed_dim: int):
        super().__init__()
        if embed_dim % 2 or (embed_dim // 2) not in (32, 64, 128):
            raise NotImplementedError(
                f'GRU hidden size {embed_dim // 2} per direction: the recurrent kernels hold W_hh in registers and '
                'support 32, 64 or 128 (feature_dim 64, 128 or 256)')
        self.embed_dim, self.word_dim, self.hidden = embed_dim, word_dim, embed_dim // 2
        self.embed = nn.Embedding(vocab_size, word_dim)
        self.rnn = _GRUParams(word_dim, embed_dim // 2)
        self.pie_net = PIENet(1, word_dim, embed_dim, word_dim // 2)
        nn.init.xavier_uniform_(self.embed.weight)          # wemb_type None branch (caption_encoder.py:61-62)
        self._len_cache = {}
        self._bound = None

    # ---- parameter store plumbing
    def adjacent_groups(self):
        return self.rnn.groups()

    def tower_params(self):
        pie = self.pie_net
        return [self.embed.weight] + [p for g in self.rnn.groups() for p in g] + \
            [pie.attention.w_1.weight, pie.attention.w_2.weight, pie.fc.weight, pie.fc.bias, pie.layer_norm.weight,
             pie.layer_norm.bias]

    def bind(self, st: ParamStore) -> None:
        """Views of the owning ParamStore the kernels read / accumulate into (rebuilt with the store)."""
        h, dw = self.hidden, self.word_dim
        g_wih, g_whh, g_bih, g_bhh = self.rnn.groups()
        self._wih16 = st.group_pad[id(g_wih[0])]                         # bf16 [6H, pad8(dw)]
        _, self._gwih = st.fused(g_wih, 6 * h, dw)                       # fp32 grad [6H, dw]
        o = st.offsets[id(g_whh[0])]
        self._whh = st.flat[o:o + 6 * h * h].view(2, 3 * h, h)
        self._gwhh = st.grad[o:o + 6 * h * h].view(2, 3 * h, h)
        o = st.offsets[id(g_bih[0])]
        self._bih, self._gbih = st.flat[o:o + 6 * h], st.grad[o:o + 6 * h]
        o = st.offsets[id(g_bhh[0])]
        self._bhh, self._gbhh = st.flat[o:o + 6 * h], st.grad[o:o + 6 * h]
        self._bound = st

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        skip = {'_len_cache': {}, '_bound': None, '_wih16': None, '_gwih': None, '_whh': None, '_gwhh': None,
                '_bih': None, '_gbih': None, '_bhh': None, '_gbhh': None, '_store': None}
        for k, v in self.__dict__.items():
            new.__dict__[k] = skip[k] if k in skip else copy.deepcopy(v, memo)
        return new

    def lengths32(self, lengths, device) -> torch.Tensor:
        """int32 device copy of the caption lengths, cached per length profile (built once, so that a CUDA-graph
        capture of the step contains no host-to-device copy)."""
        if torch.is_tensor(lengths) and lengths.is_cuda:
            return lengths.to(torch.int32).contiguous()
        host = lengths.tolist() if torch.is_tensor(lengths) else [int(v) for v in lengths]
        key = (tuple(host), str(device))
        hit = self._len_cache.get(key)
        if hit is None:
            if len(self._len_cache) > 64:
                self._len_cache.clear()
            hit = self._len_cache[key] = torch.tensor(host, dtype=torch.int32).to(device)
        return hit

    def trunk(self, x, lengths) -> torch.Tensor:
        """-> fp32 [B, D]: LayerNorm(gru_last + PIENet residual) (before l2_normalize / scale)."""
        _towers._require_cuda(x.device)
        st = self._bound
        if st is None or not st.intact():
            raise RuntimeError('text tower parameters are not in a ParamStore: call the owning model\'s .store()')
        len32 = self.lengths32(lengths, x.device)
        return _TextTowerFn.apply(x.long(), len32, self, *self.tower_params())


class GRUEncoderText(_TextTrunk):
    """caption_encoder.EncoderText (mlp_local False): forward(x, lengths) -> {'embedding': [B, D] unit rows}.
    Lives inside ClientPCME's ParamStore."""

    def forward(self, x, lengths):
        return {'embedding': ops.l2_normalize(self.trunk(x, lengths))}                # caption_encoder.py:109


class TextModel(StoreMixin, nn.Module):
    """Stand-alone multimodal-client text tower with its own parameter store (tests, microbenchmarks)."""

    def __init__(self, vocab_size=11755, word_dim=300, embed_dim=256):
        super().__init__()
        self.txt_enc = GRUEncoderText(vocab_size, word_dim, embed_dim)

    def _adjacent_groups(self):
        return self.txt_enc.adjacent_groups()

    def _after_store_build(self, st):
        self.txt_enc.bind(st)

    def forward(self, x, lengths):
        self.store()
        return self.txt_enc(x, lengths)['embedding']


class TextClient(StoreMixin, _TextTrunk):
    """Mirror of src/networks/language_model.py EncoderText (:28-130): trunk -> `* scale` -> ReLU -> classifier heads
    with ReLU-clamped weights (training) or the L2-normalised embedding."""

    def __init__(self, vocab_size=11755, word_dim=300, embed_dim=256, num_class=4, scale=128):
        super().__init__(vocab_size, word_dim, embed_dim)
        self.class_fc = _Linear(embed_dim, num_class)
        self.class_fc_2 = _Linear(embed_dim, 80)
        self.is_train, self.phase, self.scale = True, '', scale

    def _adjacent_groups(self):
        return self.adjacent_groups()

    def _after_store_build(self, st):
        self.bind(st)

    __deepcopy__ = _TextTrunk.__deepcopy__

    def forward(self, x, lengths):
        from .clients import _LinearFn
        self.store()
        out = _ScaleReluFn.apply(self.trunk(x, lengths), float(self.scale))          # language_model.py:111-112
        if self.is_train:
            for fc in (self.class_fc, self.class_fc_2):                              # :115-121, in-place clamp
                T.relu_inplace(fc.weight.data, fc.weight._w16)
            x1 = _LinearFn.apply(out, self.class_fc.weight, self.class_fc.bias)
            x2 = _LinearFn.apply(out, self.class_fc_2.weight, self.class_fc_2.bias)
            return x1, x2, self.class_fc.weight, self.class_fc_2.weight
        return ops.l2_normalize(out)                                                 # :128
