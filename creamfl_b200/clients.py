"""Client-side models of the hot path.

ClientPCME mirrors the multimodal client's PCME(ResNet18 + GRU) (src/networks/models/pcme.py with
`config.not_bert = True`, forced at MMFL.py:163): the image tower (towers.EncoderImage) and the GRU text tower
(text_towers.GRUEncoderText: Embedding -> packed bi-GRU -> PIENet, caption_encoder.py:29-116) both run on the
creamfl_b200 kernels and share one flat parameter store.  ImageClient / TextClient mirror the unimodal clients
(src/networks/resnet_client.py, src/networks/language_model.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops, tower_ops as T
from .text_towers import GRUEncoderText, TextClient  # noqa: F401  (re-exported: src/networks/language_model.py)
from .towers import EncoderImage, ResNet, StoreMixin, _Linear, fork_stream, grad_target


class ClientPCME(StoreMixin, nn.Module):
    """PCME(ResNet18 + GRU) of a multimodal client; same forward signature and output dict as pcme.PCME."""

    def __init__(self, vocab_size=11755, embed_dim=256, cnn_type='resnet18', word_dim=300):
        super().__init__()
        self.embed_dim = embed_dim
        self.n_embeddings = 1
        self.img_enc = EncoderImage({'embed_dim': embed_dim, 'cnn_type': cnn_type})
        self.txt_enc = GRUEncoderText(vocab_size, word_dim, embed_dim)

    def _adjacent_groups(self):
        return self.txt_enc.adjacent_groups()

    def _after_store_build(self, st):
        self.txt_enc.bind(st)

    def image_forward(self, images):
        self.store()
        return self.img_enc(images)

    @torch.no_grad()
    def copy_weights_from(self, other: 'ClientPCME') -> None:
        """In-place copy of every parameter and buffer (keeps addresses stable: the per-round `old_model` of
        MMClientTrainer.py:92 is refreshed instead of re-allocated, so captured CUDA graphs stay valid)."""
        a, b = self.store(), other.store()
        a.flat.copy_(b.flat)
        a.shadow.copy_(b.shadow)
        for (dst, _), (src, _) in zip(a.padded, b.padded):
            dst.copy_(src)
        for p, q in zip(self.buffers(), other.buffers()):
            p.copy_(q)

    def forward(self, images, sentences, captions_word, lengths):
        self.store()
        # GRU text tower (a latency-bound recurrence) on a forked stream next to the image tower (towers.fork_stream)
        with fork_stream(images.device, getattr(self, 'overlap_towers', True)) as side:
            with side:
                caption_output = self.txt_enc(sentences, lengths)
            image_output = self.img_enc(images)
        return {
            'image_features': image_output['embedding'], 'image_attentions': None, 'image_residuals': None,
            'image_logsigma': None, 'image_logsigma_att': None,
            'caption_features': caption_output['embedding'], 'caption_attentions': None, 'caption_residuals': None,
            'caption_logsigma': None, 'caption_logsigma_att': None,
        }


# ===================================================================================================== unimodal image client
def _pad_cols(t: torch.Tensor, mult: int = 8) -> torch.Tensor:
    """Zero-pad the last dimension to a multiple of `mult` (TMA needs 16-byte row pitches; class counts such as 100
    or 4 are not multiples of 8).  Head tensors only - a few KB."""
    c = t.shape[1]
    cp = (c + mult - 1) // mult * mult
    if cp == c:
        return t
    out = torch.zeros((t.shape[0], cp), dtype=t.dtype, device=t.device)
    out[:, :c] = t
    return out


def _pad_rows(t: torch.Tensor, mult: int = 8) -> torch.Tensor:
    r = t.shape[0]
    rp = (r + mult - 1) // mult * mult
    if rp == r:
        return t
    out = torch.zeros((rp, t.shape[1]), dtype=t.dtype, device=t.device)
    out[:r] = t
    return out


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the tcgen05 GEMM: x fp32/bf16 [R, K] (rounded to bf16), W a ParamStore parameter [N, K],
    y fp32 [R, N].  dW / db accumulate into the flat gradient buffer."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x16 = ops.to_bf16(x.detach()) if x.dtype != torch.bfloat16 else x.detach()
        y = ops.gemm_bf16(x16, weight._w16, bias=bias, out_dtype=torch.float32)
        ctx.save_for_backward(x16)
        ctx.weight, ctx.bias, ctx.x_dtype = weight, bias, x.dtype
        return y

    @staticmethod
    def backward(ctx, dy):
        (x16,) = ctx.saved_tensors
        w = ctx.weight
        n = w.shape[0]
        dy16 = _pad_cols(ops.to_bf16(dy.contiguous().float()))
        if dy16.shape[1] == n:
            ops.gemm_bf16(dy16, x16, a_mn=True, b_mn=True, out=grad_target(w), split_k=0, accumulate=True)
            w16 = w._w16
        else:                                            # ragged class count: go through padded temporaries
            gw = ops.gemm_bf16(dy16, x16, a_mn=True, b_mn=True, out_dtype=torch.float32)
            grad_target(w).add_(gw[:n])
            w16 = _pad_rows(w._w16)
        if ctx.bias is not None:
            if dy16.shape[1] == n:
                T.colsum_into(dy16, grad_target(ctx.bias))
            else:
                tmp = torch.zeros(dy16.shape[1], dtype=torch.float32, device=dy16.device)
                T.colsum_into(dy16, tmp)
                grad_target(ctx.bias).add_(tmp[:n])
        dx = ops.gemm_bf16(dy16, w16, b_mn=True, out_dtype=torch.float32) if ctx.needs_input_grad[0] else None
        return (dx.to(ctx.x_dtype) if dx is not None else None), None, None


class _GramFn(torch.autograd.Function):
    """W W^T for the class-centre loss (ClientTrainer.py:353): fp32 [C, C]; gradient (dG + dG^T) W accumulated."""

    @staticmethod
    def forward(ctx, weight):
        ctx.weight = weight
        return ops.gemm_bf16(weight._w16, weight._w16, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, dg):
        w = ctx.weight
        sym = _pad_cols(ops.to_bf16((dg + dg.t()).contiguous().float()))
        # d/dW of sum(dG * W W^T) = (dG + dG^T) W ; entries the ReLU clamp zeroed carry no gradient (relu'(w) = 0)
        gw = ops.gemm_bf16(sym, _pad_rows(w._w16), b_mn=True, out_dtype=torch.float32)
        grad_target(w).add_(gw * (w.data > 0))
        return None


class _PoolFn(torch.autograd.Function):
    """avg_pool over H*W times `scale` (resnet_client.py:177-179): NHWC bf16 map -> fp32 [N, C]."""

    @staticmethod
    def forward(ctx, x, scale):
        y, _ = T.avgpool_fwd(x, scale)
        ctx.shape, ctx.scale = tuple(x.shape), scale
        return y

    @staticmethod
    def backward(ctx, dy):
        return T.avgpool_bwd(ops.to_bf16(dy.contiguous().float()), ctx.shape, ctx.scale), None


class ImageClient(StoreMixin, nn.Module):
    """Mirror of src/networks/resnet_client.py ResNet (resnet18_client, :220-232): trunk on the creamfl_b200
    kernels, `x * scale`, Linear(512, D); `phase == 'extract_conv_feature'` returns the L2-normalised embedding,
    otherwise the classifier heads with ReLU-clamped weights (:193-201)."""

    def __init__(self, num_class=100, embed_dim=256, scale=128, is_train=True, phase='none', arch='resnet18'):
        super().__init__()
        trunk = ResNet(arch)
        # same attribute names as the reference so that state_dict keys match
        self.conv1, self.bn1 = trunk.conv1, trunk.bn1
        self.layer1, self.layer2, self.layer3, self.layer4 = trunk.layer1, trunk.layer2, trunk.layer3, trunk.layer4
        self._trunk = [trunk]                      # not a registered submodule (parameters are registered above)
        self.embed_dim = embed_dim
        if embed_dim != 512:
            self.linear = _Linear(512, embed_dim)
        self.class_fc_2 = _Linear(embed_dim, num_class)
        self.class_fc_22 = _Linear(embed_dim, 80)
        self.is_train, self.scale, self.phase = bool(is_train), int(scale), str(phase)

    def forward(self, x):
        self.store()
        fmap = self._trunk[0](x)
        feat = _PoolFn.apply(fmap, float(self.scale))
        if self.embed_dim != 512:
            feat = _LinearFn.apply(feat, self.linear.weight, self.linear.bias)
        if self.phase == 'extract_conv_feature':
            return ops.l2_normalize(feat)
        if self.is_train:
            for fc in (self.class_fc_2, self.class_fc_22):
                T.relu_inplace(fc.weight.data, fc.weight._w16)
            x1 = _LinearFn.apply(feat, self.class_fc_2.weight, self.class_fc_2.bias)
            x2 = _LinearFn.apply(feat, self.class_fc_22.weight, self.class_fc_22.bias)
            return x1, x2, self.class_fc_2.weight, self.class_fc_22.weight
        return feat


def resnet18_client(pretrained=False, **kwargs):
    """src/networks/resnet_client.py:220-232 (ImageNet weights are not on the box: `pretrained` is ignored)."""
    return ImageClient(num_class=kwargs.get('num_class', 100), embed_dim=kwargs.get('embed_dim', 256),
                       scale=kwargs.get('scale', 128), is_train=kwargs.get('is_train', True),
                       phase=kwargs.get('phase', 'none'))


def unimodal_supervised_loss(model: ImageClient, inputs, labels, inter_distance: float = 4.0):
    """ClientTrainer.tra supervised pass (ClientTrainer.py:322-356): CE(fvec - 4*onehot, y) + 0.5 * CE(W W^T, arange)."""
    fvec, _, class_weight, _ = model(inputs)
    loss = ops.cross_entropy(fvec, labels, inter_distance)
    center = ops.cross_entropy(_GramFn.apply(class_weight), None, 0.0)
    return 0.5 * center + loss, fvec


# ===================================================================================================== unimodal text client
def text_supervised_loss(model: TextClient, captions, lengths, labels, inter_distance: float = 4.0):
    """ClientTrainer.tra supervised pass for the text clients (ClientTrainer.py:335-356)."""
    fvec, _, class_weight, _ = model(captions, lengths)
    loss = ops.cross_entropy(fvec, labels, inter_distance)
    center = ops.cross_entropy(_GramFn.apply(class_weight), None, 0.0)
    return 0.5 * center + loss, fvec
