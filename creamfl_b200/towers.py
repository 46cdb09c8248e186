"""Encoder towers of the CreamFL hot path on the creamfl_b200 kernels.

Mirrors the modules the reference reaches through `get_model` (src/networks/models/__init__.py:6):
  * ResNet           - torchvision ResNet-{18,101} feature extractor as used by EncoderImage
                       (src/networks/models/image_encoder.py:24-32,55): NHWC bf16, convolutions as implicit GEMM on
                       tcgen05, BatchNorm with batch statistics, one autograd node per residual block
  * EncoderImage     - ResNet -> avgpool -> fc -> PIENet -> LayerNorm -> l2_normalize (image_encoder.py:54-71,
                       pie_model.py:28-40,61-67)
  * BertEncoder      - HF BertModel (bert-base) forward/backward restricted to what pcme.py:43-44 consumes
                       (last_hidden_state[:, 0]); state_dict keys equal HF's
Parameter names / shapes equal the reference's so `{'net': state_dict}` checkpoints (MMFL.py:281) load unchanged.

Numerics: fp32 master parameters, bf16 shadows as tensor-core operands, bf16 activations, fp32 accumulation and
statistics (the policy replacing apex O2).  Weight gradients are accumulated by the kernels
directly into a flat fp32 gradient buffer that `param.grad` views.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops, tower_ops as T

BF16 = torch.bfloat16


# ===================================================================================================== parameters
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def _require_cuda(dev: torch.device) -> None:
    if dev.type != 'cuda':
        raise RuntimeError('creamfl_b200 towers run on CUDA only (no CPU fallback exists); call .cuda() first')


class ParamStore:
    """All parameters of a model in one flat fp32 buffer, with a flat fp32 gradient buffer and a flat bf16 shadow.

    4-D (convolution) parameters keep their logical OIHW shape but live in memory as [O, H, W, I] (torch
    channels_last), the order the implicit-GEMM kernels read.  `adjacent` lists parameter groups that must be
    contiguous (BERT query/key/value -> one [2304, 768] operand)."""

    def __init__(self, module: nn.Module, adjacent: Sequence[Sequence[nn.Parameter]] = ()):
        params: List[nn.Parameter] = []
        seen = set()
        for grp in adjacent:
            for p in grp:
                if id(p) not in seen:
                    seen.add(id(p))
                    params.append(p)
        for p in module.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        if not params:
            raise ValueError('ParamStore: module has no parameters')
        dev = params[0].device
        _require_cuda(dev)
        in_group = {id(p) for grp in adjacent for p in grp}
        offs, total = [], 0
        for p in params:
            if id(p) not in in_group:
                total = (total + 63) // 64 * 64
            offs.append(total)
            total += p.numel()
        total = (total + 63) // 64 * 64
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.shadow = torch.zeros(total, dtype=BF16, device=dev)
        self.params = params
        self.offsets = {}
        self.padded = []
        for p, off in zip(params, offs):
            n = p.numel()
            self.offsets[id(p)] = off
            if p.dim() == 4:
                o, i, r, s = p.shape
                view = self.flat[off:off + n].view(o, r, s, i).permute(0, 3, 1, 2)
                gview = self.grad[off:off + n].view(o, r, s, i).permute(0, 3, 1, 2)
                k = r * s * i
                if k % 8 == 0:
                    p._w16 = self.shadow[off:off + n].view(o, k)
                else:  # 7x7x3 stem: rows padded to a 16-byte pitch (TMA), zero tail
                    p._w16 = torch.zeros((o, (k + 7) // 8 * 8), dtype=BF16, device=dev)
                    self.padded.append((p._w16, self.shadow[off:off + n].view(o, k)))
                p._g2d = self.grad[off:off + n].view(o, k)
            else:
                view = self.flat[off:off + n].view(p.shape)
                gview = self.grad[off:off + n].view(p.shape)
                p._w16 = self.shadow[off:off + n].view(p.shape)
                if p.dim() == 2 and p.shape[1] % 8 and id(p) not in in_group:
                    # K = 300 operands of the GRU text towers: row pitch padded to 16 bytes (TMA), rows to 8
                    buf = torch.zeros((_pad8(p.shape[0]), _pad8(p.shape[1])), dtype=BF16, device=dev)
                    self.padded.append((buf[:p.shape[0]], p._w16))
                    p._w16, p._w16p = buf[:p.shape[0]], buf
                p._g2d = gview
            view.copy_(p.data)
            p.data = view
            p._gview = gview
            p.grad = gview if p.requires_grad else None
        # adjacent 2-D groups with an unaligned inner dimension share one padded operand (both directions of the
        # GRU input projection -> one [6H, 304] matrix)
        self.group_pad = {}
        for grp in adjacent:
            if all(p.dim() == 2 for p in grp) and len({p.shape[1] for p in grp}) == 1 and grp[0].shape[1] % 8:
                rows, cols = sum(p.shape[0] for p in grp), grp[0].shape[1]
                buf = torch.zeros((_pad8(rows), _pad8(cols)), dtype=BF16, device=dev)
                r0 = 0
                for p in grp:
                    off = self.offsets[id(p)]
                    self.padded.append((buf[r0:r0 + p.shape[0]], self.shadow[off:off + p.numel()].view(p.shape)))
                    p._w16, p._w16p = buf[r0:r0 + p.shape[0]], buf
                    r0 += p.shape[0]
                self.group_pad[id(grp[0])] = buf[:rows]
        self.first_ptr = params[0].data_ptr()
        self.sync_shadow()

    def fused(self, group: Sequence[nn.Parameter], rows: int, cols: Optional[int] = None):
        """bf16 shadow and fp32 grad of an adjacent group viewed as one tensor."""
        off = self.offsets[id(group[0])]
        n = sum(p.numel() for p in group)
        shape = (rows,) if cols is None else (rows, cols)
        return self.shadow[off:off + n].view(shape), self.grad[off:off + n].view(shape)

    def sync_shadow(self) -> None:
        """Refresh the bf16 shadow after the fp32 masters changed (optimizer step, load_state_dict)."""
        ops.cast_into(self.flat, self.shadow)
        for dst, src in self.padded:
            dst[:, :src.shape[1]].copy_(src)

    def zero_grad(self) -> None:
        self.grad.zero_()
        for p in self.params:
            if p.requires_grad:
                p.grad = p._gview

    def intact(self) -> bool:
        return self.params[0].data_ptr() == self.first_ptr


class fork_stream:
    """`with fork_stream(device) as side: with side: <branch>; <main work>` - runs <branch> on a side stream forked
    from the current stream and joins it on exit (capturable: inside a CUDA-graph capture the fork / join become
    graph edges and the two branches replay concurrently).  `enabled=False` (or a CPU device) runs everything inline."""
    _streams = {}

    def __init__(self, device, enabled: bool = True, slot: int = 0):
        import os
        self.enabled = bool(enabled) and device.type == 'cuda' and os.environ.get('CREAMFL_NO_OVERLAP') != '1'
        self.device, self.slot = device, slot
        self.side = self.main = None

    class _Branch:
        def __init__(self, outer):
            self.outer, self.ctx = outer, None

        def __enter__(self):
            o = self.outer
            if o.enabled:
                o.side.wait_stream(o.main)
                self.ctx = torch.cuda.stream(o.side)
                self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            if self.ctx is not None:
                self.ctx.__exit__(*exc)
            return False

    def __enter__(self):
        if self.enabled:
            key = (self.device.index, self.slot, T.current_lane())
            if key not in fork_stream._streams:
                fork_stream._streams[key] = torch.cuda.Stream(self.device)
            self.side, self.main = fork_stream._streams[key], torch.cuda.current_stream(self.device)
        return fork_stream._Branch(self)

    def __exit__(self, *exc):
        if self.enabled:
            self.main.wait_stream(self.side)
        return False


def grad_target(p: nn.Parameter) -> torch.Tensor:
    """The fp32 buffer a kernel accumulates d(loss)/dp into.  Handles `optimizer.zero_grad(set_to_none=True)`."""
    if p.grad is None:
        p._g2d.zero_()
        p.grad = p._gview
    elif p.grad.data_ptr() != p._gview.data_ptr():
        raise RuntimeError('parameter .grad was replaced; creamfl_b200 accumulates into its flat gradient buffer')
    return p._g2d


class StoreMixin:
    """Modules that own a ParamStore: (re)build lazily - deepcopy / .to() / .cpu() re-home the parameters."""
    _store: Optional[ParamStore] = None

    def _adjacent_groups(self):
        return []

    def store(self) -> ParamStore:
        st = self.__dict__.get('_store')
        if st is None or not st.intact() or st.params[0] is not next(iter(self._ordered_first())):
            st = ParamStore(self, self._adjacent_groups())
            self.__dict__['_store'] = st
            self._after_store_build(st)
        return st

    def _ordered_first(self):
        groups = self._adjacent_groups()
        if groups:
            yield groups[0][0]
        else:
            yield next(self.parameters())

    def _after_store_build(self, st: ParamStore) -> None:
        pass

    def sync_shadow(self) -> None:
        self.store().sync_shadow()

    def zero_grad(self, set_to_none: bool = False) -> None:  # noqa: D401 - nn.Module signature
        self.store().zero_grad()

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == '_store':
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new.__dict__['_store'] = None
        return new


# ===================================================================================================== ResNet
class Conv(nn.Module):
    def __init__(self, cin, cout, k, stride, pad):
        super().__init__()
        self.k, self.stride, self.pad = k, stride, pad
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode='fan_out', nonlinearity='relu')


class BN(nn.Module):
    def __init__(self, c, eps=1e-5, momentum=0.1):
        super().__init__()
        self.eps, self.momentum = eps, momentum
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer('running_mean', torch.zeros(c))
        self.register_buffer('running_var', torch.ones(c))
        self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        self._scratch = None

    def scratch(self) -> T.BNScratch:
        sc = self._scratch
        if sc is None or sc.sums.device != self.weight.device:
            sc = self._scratch = T.BNScratch(self.weight.numel(), self.weight.device)
        return sc

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_scratch' else copy.deepcopy(v, memo)
        return new

    def fwd(self, x, res=None, relu=True, stats_ready=False, keep_gate=False):
        """Returns (y, saved) - saved is (mean, rstd, gate bits or None) in training mode, None in eval mode.
        stats_ready: the producing convolution already accumulated the batch statistics into scratch().sums.
        keep_gate (residual before the ReLU, backward pass will follow): the ReLU gate is kept as one bit per element
        (1/16 of y) for bwd()."""
        if self.training:
            want_mask = keep_gate and res is not None and relu
            out = T.bn_train_fwd(x, self.weight, self.bias, self.running_mean, self.running_var,
                                 self.scratch(), self.eps, self.momentum, res=res, relu=relu,
                                 stats_ready=stats_ready, num_batches_tracked=self.num_batches_tracked,
                                 want_mask=want_mask)
            return out[0], (out[1], out[2], out[3] if want_mask else None)
        return T.bn_eval_fwd(x, self.weight, self.bias, self.running_mean, self.running_var, self.scratch(), self.eps,
                             res=res, relu=relu), None

    def bwd(self, dy, y_mask, x, saved, want_g=False, relu_from_x=False):
        """`y_mask`: the forward output when a residual was added before the ReLU (read only if the forward pass kept
        no gate bits); `relu_from_x`: ReLU without residual - the gate is recomputed from x inside the kernels and the
        output is not read."""
        if saved is None:
            raise RuntimeError('BatchNorm backward needs a training-mode forward (batch statistics)')
        mean, rstd, mask = saved
        return T.bn_train_bwd(dy, None if mask is not None else y_mask, x, self.weight, mean, rstd, self.scratch(),
                              grad_target(self.weight), grad_target(self.bias), want_g=want_g, beta=self.bias,
                              relu_from_x=relu_from_x, mask=mask)


def _conv_f(c: Conv, x, bn: 'BN' = None):
    """Convolution; with `bn` (training mode) the batch statistics of the output are produced on the way out."""
    sums = bn.scratch().sums if (bn is not None and bn.training) else None
    return T.conv_fprop(x, c.weight._w16, c.k, c.k, c.stride, c.pad, bn_sums=sums)


def _conv_b(c: Conv, dy, x, need_dx=True, add=None):
    T.conv_wgrad(dy, x, grad_target(c.weight), c.k, c.k, c.stride, c.pad)
    if not need_dx:
        return None
    return T.conv_dgrad(dy, c.weight._w16, x.shape, c.k, c.k, c.stride, c.pad, add=add)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = Conv(inplanes, planes, 1, 1, 0), BN(planes)
        self.conv2, self.bn2 = Conv(planes, planes, 3, stride, 1), BN(planes)       # torchvision v1.5: stride on 3x3
        self.conv3, self.bn3 = Conv(planes, planes * 4, 1, 1, 0), BN(planes * 4)
        self.downsample = downsample

    def block_params(self):
        return [p for p in self.parameters()]

    def forward(self, x):
        return _BlockFn.apply(x, self, *self.block_params())

    def run_fwd(self, x, save):
        o1 = _conv_f(self.conv1, x, self.bn1)
        a1, s1 = self.bn1.fwd(o1, stats_ready=True)
        o2 = _conv_f(self.conv2, a1, self.bn2)
        a2, s2 = self.bn2.fwd(o2, stats_ready=True)
        o3 = _conv_f(self.conv3, a2, self.bn3)
        if self.downsample is not None:
            od = _conv_f(self.downsample[0], x, self.downsample[1])
            idn, sd = self.downsample[1].fwd(od, relu=False, stats_ready=True)
        else:
            od, idn, sd = None, x, None
        y, s3 = self.bn3.fwd(o3, res=idn, relu=True, stats_ready=True, keep_gate=save)
        if save:
            return y, (x, o1, a1, s1, o2, a2, s2, o3, s3, od, sd, y)
        return y, None

    def run_eval(self, x, fold):
        """Inference: every BatchNorm folded into its convolution (ResNet.fold_eval), ReLU / residual in the epilogues."""
        a1 = fold.conv(self.conv1, x, relu=True)
        a2 = fold.conv(self.conv2, a1, relu=True)
        idn = x if self.downsample is None else fold.conv(self.downsample[0], x, relu=False)
        return fold.conv(self.conv3, a2, add=idn, relu=True)

    def run_bwd(self, saved, dy):
        x, o1, a1, s1, o2, a2, s2, o3, s3, od, sd, y = saved
        do3, g = self.bn3.bwd(dy, y, o3, s3, want_g=True)
        da2 = _conv_b(self.conv3, do3, a2)
        do2, _ = self.bn2.bwd(da2, None, o2, s2, relu_from_x=True)
        da1 = _conv_b(self.conv2, do2, a1)
        do1, _ = self.bn1.bwd(da1, None, o1, s1, relu_from_x=True)
        if self.downsample is not None:
            dod, _ = self.downsample[1].bwd(g, None, od, sd)
            dxd = _conv_b(self.downsample[0], dod, x)
            return _conv_b(self.conv1, do1, x, add=dxd)
        return _conv_b(self.conv1, do1, x, add=g)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1, self.bn1 = Conv(inplanes, planes, 3, stride, 1), BN(planes)
        self.conv2, self.bn2 = Conv(planes, planes, 3, 1, 1), BN(planes)
        self.downsample = downsample

    def block_params(self):
        return [p for p in self.parameters()]

    def forward(self, x):
        return _BlockFn.apply(x, self, *self.block_params())

    def run_fwd(self, x, save):
        o1 = _conv_f(self.conv1, x, self.bn1)
        a1, s1 = self.bn1.fwd(o1, stats_ready=True)
        o2 = _conv_f(self.conv2, a1, self.bn2)
        if self.downsample is not None:
            od = _conv_f(self.downsample[0], x, self.downsample[1])
            idn, sd = self.downsample[1].fwd(od, relu=False, stats_ready=True)
        else:
            od, idn, sd = None, x, None
        y, s2 = self.bn2.fwd(o2, res=idn, relu=True, stats_ready=True, keep_gate=save)
        if save:
            return y, (x, o1, a1, s1, o2, s2, od, sd, y)
        return y, None

    def run_eval(self, x, fold):
        a1 = fold.conv(self.conv1, x, relu=True)
        idn = x if self.downsample is None else fold.conv(self.downsample[0], x, relu=False)
        return fold.conv(self.conv2, a1, add=idn, relu=True)

    def run_bwd(self, saved, dy):
        x, o1, a1, s1, o2, s2, od, sd, y = saved
        do2, g = self.bn2.bwd(dy, y, o2, s2, want_g=True)
        da1 = _conv_b(self.conv2, do2, a1)
        do1, _ = self.bn1.bwd(da1, None, o1, s1, relu_from_x=True)
        if self.downsample is not None:
            dod, _ = self.downsample[1].bwd(g, None, od, sd)
            dxd = _conv_b(self.downsample[0], dod, x)
            return _conv_b(self.conv1, do1, x, add=dxd)
        return _conv_b(self.conv1, do1, x, add=g)


class _BlockFn(torch.autograd.Function):
    """One residual block = one autograd node; the block's own run_fwd/run_bwd sequence the kernels."""

    @staticmethod
    def forward(ctx, x, block, *params):
        need = any(ctx.needs_input_grad)
        y, saved = block.run_fwd(x, need)
        ctx.block, ctx.saved = block, saved
        return y

    @staticmethod
    def backward(ctx, dy):
        if ctx.saved is None:
            raise RuntimeError('block was run without saving activations')
        dx = ctx.block.run_bwd(ctx.saved, dy.contiguous())
        ctx.saved = None
        return (dx, None) + (None,) * (len(ctx.needs_input_grad) - 2)


class _StemFn(torch.autograd.Function):
    """conv 7x7/2 (im2col from fp32 NCHW images + tcgen05 GEMM) -> BN -> ReLU -> maxpool 3x3/2."""

    @staticmethod
    def forward(ctx, images, net, *params):
        n, c, h, w = images.shape
        need = any(ctx.needs_input_grad)
        w16 = net.stem_shadow()
        images = images.contiguous()
        fused = images.dtype == torch.float32 and T.stem_supported(h, w)
        if fused:
            # patches assembled in shared memory inside the implicit GEMM (csrc/stem_tc.cu): no patch matrix in HBM
            col = None
            o = T.stem_fprop(images, w16)
        else:
            col = T.im2col_images(images, 7, 7, 2, 3, w16.shape[1])
            ho, wo = T.conv_out_hw(h, w, 7, 7, 2, 3)
            o = ops.gemm_bf16(col, w16).view(n, ho, wo, 64)
        # BatchNorm -> ReLU -> maxpool in one pass over the raw convolution output (the normalised 112 x 112 map is
        # never written; the backward pass gathers the pooling gradient inside the BatchNorm backward kernels)
        y, idx, s = T.stem_tail_fwd(o, net.bn1, net.bn1.training, want_idx=need)
        ctx.net = net
        ctx.saved = (images if fused else col, fused, o, s, idx) if need else None
        return y

    @staticmethod
    def backward(ctx, dy):
        net = ctx.net
        src, fused, o, s, idx = ctx.saved
        ctx.saved = None
        if s is None:
            raise RuntimeError('BatchNorm backward needs a training-mode forward (batch statistics)')
        if net.fused_stem_backward:
            do = T.stem_tail_bwd(dy.contiguous(), idx, o, net.bn1, s, grad_target(net.bn1.weight),
                                 grad_target(net.bn1.bias))
        else:
            # measured: gathering the pooling gradient inside BOTH BatchNorm backward passes costs more than writing the
            # 112 x 112 gradient map once (client phase 389 -> 410 ms per mini-round), so the backward stays in two steps
            da = T.maxpool_bwd(dy.contiguous(), idx, o.shape)
            do, _ = net.bn1.bwd(da, None, o, s + (None,), relu_from_x=True)
        g = grad_target(net.conv1.weight)                         # [64, 147] fp32
        if fused:
            T.stem_wgrad(src, do, g)                              # patches re-assembled in shared memory from the images
        else:
            ops.gemm_bf16(do.view(-1, 64), src, a_mn=True, b_mn=True, out=g, split_k=0, accumulate=True,
                          n_cols=g.shape[1])
        return (None, None) + (None,) * (len(ctx.needs_input_grad) - 2)


class _EvalFold:
    """Folded inference weights of a ResNet (one bf16 copy of every block convolution scaled by the BatchNorm that
    follows it + the per-channel shifts) and the device tables creamfl_bn_fold_layers reads.  The fold is ONE launch
    and runs at the start of every inference forward, so it always sees the current masters / running statistics
    (also inside a captured CUDA graph)."""

    def __init__(self, net: 'ResNet'):
        pairs = net.block_conv_bn_pairs()
        dev = pairs[0][0].weight.device
        total_w = sum(c.weight.numel() for c, _ in pairs)
        total_c = sum(c.weight.shape[0] for c, _ in pairs)
        self.w16 = torch.empty(total_w, dtype=BF16, device=dev)
        self.bias = torch.empty(total_c, dtype=torch.float32, device=dev)
        self.views, self.pairs, rows, start = {}, [], [], [0]
        ow = oc = 0
        for conv, bn in pairs:
            o = conv.weight.shape[0]
            k = conv.weight.numel() // o
            wv, bv = self.w16[ow:ow + o * k].view(o, k), self.bias[oc:oc + o]
            self.views[id(conv)] = (wv, bv)
            self.pairs.append((conv, bn, wv, bv))
            rows.append([conv.weight.data_ptr(), k, o, wv.data_ptr(), k, bn.weight.data_ptr(), bn.bias.data_ptr(),
                         bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bv.data_ptr()])
            start.append(start[-1] + o)
            ow += o * k
            oc += o
        self.layers = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.row_start = torch.tensor(start, dtype=torch.int64).to(dev)
        self.total_rows = start[-1]
        self.eps = pairs[0][1].eps
        self.key = _EvalFold.key_of(pairs)

    @staticmethod
    def key_of(pairs):
        (c0, b0), (c1, b1) = pairs[0], pairs[-1]
        return (c0.weight.data_ptr(), b0.running_var.data_ptr(), c1.weight.data_ptr(), b1.running_var.data_ptr(),
                b1.weight.data_ptr())

    def refresh(self):
        T.bn_fold_layers(self.layers, self.row_start, self.total_rows, self.eps, self.pairs)

    def conv(self, c: 'Conv', x, add=None, relu=True):
        w, b = self.views[id(c)]
        return T.conv_fprop_affine(x, w, c.k, c.k, c.stride, c.pad, b, add=add, relu=relu)


class ResNet(nn.Module):
    """torchvision.models.resnet{18,101} without avgpool/fc (image_encoder.py:24-32): images fp32 NCHW in,
    final feature map NHWC bf16 out."""

    fold_eval_bn = True      # class-level switch (tests compare against the un-folded inference path)
    fused_stem_backward = False   # creamfl_bn_pool_bwd (kept, tested; slower than maxpool backward + BatchNorm backward)
    CFG = {'resnet18': (BasicBlock, [2, 2, 2, 2]), 'resnet34': (BasicBlock, [3, 4, 6, 3]),
           'resnet50': (Bottleneck, [3, 4, 6, 3]), 'resnet101': (Bottleneck, [3, 4, 23, 3]),
           'resnet152': (Bottleneck, [3, 8, 36, 3])}

    def __init__(self, arch='resnet101'):
        super().__init__()
        block, layers = self.CFG[arch]
        self.inplanes = 64
        self.conv1, self.bn1 = Conv(3, 64, 7, 2, 3), BN(64)
        self.layer1 = self._make_layer(block, 64, layers[0], 1)
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        self.out_dim = 512 * block.expansion

    def _make_layer(self, block, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(Conv(self.inplanes, planes * block.expansion, 1, stride, 0),
                                       BN(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes))
        return nn.Sequential(*layers)

    def stem_shadow(self):
        """[64, 152] bf16 filter matrix of the stem (147 columns padded to a 16-byte pitch by the ParamStore)."""
        return self.conv1.weight._w16

    def block_conv_bn_pairs(self):
        """(convolution, BatchNorm that follows it) of every residual block, in execution order (the stem keeps its
        own path: conv -> BatchNorm -> ReLU -> maxpool)."""
        pairs = []
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                pairs.append((blk.conv1, blk.bn1))
                pairs.append((blk.conv2, blk.bn2))
                if hasattr(blk, 'conv3'):
                    pairs.append((blk.conv3, blk.bn3))
                if blk.downsample is not None:
                    pairs.append((blk.downsample[0], blk.downsample[1]))
        return pairs

    def eval_fold(self) -> '_EvalFold':
        fold = self.__dict__.get('_fold')
        if fold is None or fold.key != _EvalFold.key_of(self.block_conv_bn_pairs()):
            fold = self.__dict__['_fold'] = _EvalFold(self)
        return fold

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_fold' else copy.deepcopy(v, memo)
        return new

    def forward(self, images):
        x = _StemFn.apply(images, self, self.conv1.weight, self.bn1.weight, self.bn1.bias)
        if not self.training and not torch.is_grad_enabled() and self.fold_eval_bn:
            # inference (public-feature extraction, the clients' old model): BatchNorm folded into the convolutions,
            # no normalisation pass between them
            fold = self.eval_fold()
            fold.refresh()
            for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
                for blk in layer:
                    x = blk.run_eval(x, fold)
            return x
        for layer in (self.layer1, self.layer2, self.layer3, self.layer4):
            for blk in layer:
                x = blk(x)
        return x


# ===================================================================================================== image head
class _Linear(nn.Module):
    def __init__(self, din, dout, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(dout, din))
        self.bias = nn.Parameter(torch.zeros(dout)) if bias else None
        nn.init.xavier_uniform_(self.weight)


class _LayerNormP(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class _SelfAttnPool(nn.Module):
    def __init__(self, n_head, d_in, d_hidden):
        super().__init__()
        self.w_1 = _Linear(d_in, d_hidden, bias=False)
        self.w_2 = _Linear(d_hidden, n_head, bias=False)


class PIENet(nn.Module):
    """Parameter container of pie_model.PIENet (n_embeds = 1)."""

    def __init__(self, n_embeds, d_in, d_out, d_h):
        super().__init__()
        if n_embeds != 1:
            raise NotImplementedError('the reference instantiates PIENet with n_embeds = 1 only')
        self.attention = _SelfAttnPool(n_embeds, d_in, d_h)
        self.fc = _Linear(d_in, d_out)
        self.layer_norm = _LayerNormP(d_out, 1e-5)


class _ImageHeadFn(torch.autograd.Function):
    """x7 [B,7,7,C] -> l2_normalize(LayerNorm(fc(avgpool) + sigmoid(pie.fc(attention-pool)))) (image_encoder.py:55-67)."""

    @staticmethod
    def forward(ctx, x7, enc, *params):
        b, h, w, c = x7.shape
        p = h * w
        pie = enc.pie_net
        x2 = x7.reshape(b * p, c)
        hid = ops.gemm_bf16(x2, pie.attention.w_1.weight._w16, act=ops.ACT_TANH)
        attn, r, pooled = T.pie_pool_fwd(x7.view(b, p, c), hid.view(b, p, -1), pie.attention.w_2.weight.view(-1))
        out = ops.gemm_bf16(pooled, enc.fc.weight._w16, bias=enc.fc.bias, out_dtype=torch.float32)
        res = ops.gemm_bf16(r, pie.fc.weight._w16, bias=pie.fc.bias, act=ops.ACT_SIGMOID, out_dtype=torch.float32)
        z, mean, rstd = T.layernorm_fwd(out, pie.layer_norm.weight, pie.layer_norm.bias, pie.layer_norm.eps, res=res)
        emb, inv = ops.l2norm_raw(z)
        ctx.enc = enc
        ctx.saved = (x7, hid, attn, r, pooled, out, res, mean, rstd, emb, inv) if any(ctx.needs_input_grad) else None
        return emb

    @staticmethod
    def backward(ctx, demb):
        enc = ctx.enc
        pie = enc.pie_net
        x7, hid, attn, r, pooled, out, res, mean, rstd, emb, inv = ctx.saved
        ctx.saved = None
        b, h, w, c = x7.shape
        p = h * w
        dz = ops.l2norm_bwd_raw(demb.contiguous().float(), emb, inv)
        dsum = T.layernorm_bwd(dz, out, pie.layer_norm.weight, mean, rstd, grad_target(pie.layer_norm.weight),
                               grad_target(pie.layer_norm.bias), res=res)
        d_out16 = ops.to_bf16(dsum)
        d_respre = T.act_bwd(dsum, res, ops.ACT_SIGMOID)
        # fc: out = pooled Wfc^T + b
        ops.gemm_bf16(d_out16, pooled, a_mn=True, b_mn=True, out=grad_target(enc.fc.weight), split_k=0, accumulate=True)
        T.colsum_into(d_out16, grad_target(enc.fc.bias))
        d_pooled = ops.gemm_bf16(d_out16, enc.fc.weight._w16, b_mn=True)
        # pie.fc: res = sigmoid(r Wp^T + b)
        ops.gemm_bf16(d_respre, r, a_mn=True, b_mn=True, out=grad_target(pie.fc.weight), split_k=0, accumulate=True)
        T.colsum_into(d_respre, grad_target(pie.fc.bias))
        d_r = ops.gemm_bf16(d_respre, pie.fc.weight._w16, b_mn=True)
        dx_part, dpre = T.pie_pool_bwd(x7.view(b, p, c), hid.view(b, p, -1), pie.attention.w_2.weight.view(-1), attn,
                                       d_r, d_pooled, grad_target(pie.attention.w_2.weight).view(-1))
        x2 = x7.reshape(b * p, c)
        dpre2 = dpre.view(b * p, -1)
        ops.gemm_bf16(dpre2, x2, a_mn=True, b_mn=True, out=grad_target(pie.attention.w_1.weight), split_k=0,
                      accumulate=True)
        dx = ops.gemm_bf16(dpre2, pie.attention.w_1.weight._w16, b_mn=True, add=dx_part.view(b * p, c))
        return (dx.view(b, h, w, c), None) + (None,) * (len(ctx.needs_input_grad) - 2)


class EncoderImage(nn.Module):
    """Mirror of src/networks/models/image_encoder.py:17-71 (mlp_local=False)."""

    def __init__(self, config, mlp_local=False):
        super().__init__()
        if mlp_local:
            raise NotImplementedError('mlp_local heads are hard-wired to 512-d in the reference (SURVEY appendix B); '
                                      'not part of the accelerated path')
        embed_dim = config['embed_dim'] if isinstance(config, dict) else config.embed_dim
        cnn_type = config['cnn_type'] if isinstance(config, dict) else config.cnn_type
        self.cnn = ResNet(cnn_type)
        self.cnn_dim = self.cnn.out_dim
        self.fc = _Linear(self.cnn_dim, embed_dim)
        self.pie_net = PIENet(1, self.cnn_dim, embed_dim, self.cnn_dim // 2)
        self.mlp_local = mlp_local

    def head_params(self):
        return [self.fc.weight, self.fc.bias, self.pie_net.attention.w_1.weight, self.pie_net.attention.w_2.weight,
                self.pie_net.fc.weight, self.pie_net.fc.bias, self.pie_net.layer_norm.weight,
                self.pie_net.layer_norm.bias]

    def forward(self, images):
        x7 = self.cnn(images)
        # the reference returns only 'embedding' from EncoderImage.forward (image_encoder.py:69-71)
        return {'embedding': _ImageHeadFn.apply(x7, self, *self.head_params())}


# ===================================================================================================== BERT
class _BertSelfAttention(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.query, self.key, self.value = _Linear(d, d), _Linear(d, d), _Linear(d, d)


class _BertSelfOutput(nn.Module):
    def __init__(self, din, d, eps):
        super().__init__()
        self.dense = _Linear(din, d)
        self.LayerNorm = _LayerNormP(d, eps)


class _BertAttention(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.self = _BertSelfAttention(d)
        self.output = _BertSelfOutput(d, d, eps)


class _BertIntermediate(nn.Module):
    def __init__(self, d, dff):
        super().__init__()
        self.dense = _Linear(d, dff)


class _BertLayer(nn.Module):
    def __init__(self, d, dff, eps):
        super().__init__()
        self.attention = _BertAttention(d, eps)
        self.intermediate = _BertIntermediate(d, dff)
        self.output = _BertSelfOutput(dff, d, eps)


class _BertEmbeddings(nn.Module):
    def __init__(self, vocab, d, max_pos, types, eps):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, d, padding_idx=0)
        self.position_embeddings = nn.Embedding(max_pos, d)
        self.token_type_embeddings = nn.Embedding(types, d)
        self.LayerNorm = _LayerNormP(d, eps)


class _BertEncoderStack(nn.Module):
    def __init__(self, n, d, dff, eps):
        super().__init__()
        self.layer = nn.ModuleList([_BertLayer(d, dff, eps) for _ in range(n)])


class _BertPooler(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.dense = _Linear(d, d)       # dead on this path (pcme.py:44 reads last_hidden_state), kept for checkpoints


class BertEncoder(nn.Module):
    """bert-base-uncased geometry (HF BertConfig defaults), including its train-mode dropout:
    hidden_dropout_prob = attention_probs_dropout_prob = 0.1 at 37 sites per forward (embeddings, and per layer the
    attention probabilities, the attention output dense and the FFN output dense).  The reference trains with it on
    (pcme.py:31 builds HF BertModel with default config; retrieval_trainer.py:187 / MMFL.py:293 call model.train()).
    The masks are Philox-generated inside the kernels (csrc/philox.cuh) and regenerated in the backward; eval mode
    applies none.  `dropout_p = 0` reproduces the frozen-dropout parity protocol of SURVEY 3.2."""

    def __init__(self, vocab=30522, hidden=768, layers=12, heads=12, ffn=3072, max_pos=512, types=2, eps=1e-12,
                 dropout_p=0.1, seed=None):
        super().__init__()
        self.hidden, self.heads = hidden, heads
        self.dropout_p = float(dropout_p)
        self._seed = int(torch.initial_seed() if seed is None else seed)
        self._drop_state = None
        self.embeddings = _BertEmbeddings(vocab, hidden, max_pos, types, eps)
        self.encoder = _BertEncoderStack(layers, hidden, ffn, eps)
        self.pooler = _BertPooler(hidden)
        for m in self.modules():
            if isinstance(m, _Linear):
                nn.init.normal_(m.weight, std=0.02)
            elif isinstance(m, nn.Embedding):
                nn.init.normal_(m.weight, std=0.02)
        with torch.no_grad():
            self.embeddings.word_embeddings.weight[0].zero_()

    def dropout_state(self, device) -> 'T.DropoutState':
        """Device RNG state {seed, step} of this tower's dropouts (created lazily on the tower's device)."""
        st = self._drop_state
        if st is None or st.rng.device != device or st.p != self.dropout_p:
            st = self._drop_state = T.DropoutState(self._seed, self.dropout_p, device)
        return st

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == '_drop_state' else copy.deepcopy(v, memo)
        return new

    def qkv_groups(self):
        groups = []
        for lyr in self.encoder.layer:
            s = lyr.attention.self
            groups.append([s.query.weight, s.key.weight, s.value.weight])
            groups.append([s.query.bias, s.key.bias, s.value.bias])
        return groups


class _BertFn(torch.autograd.Function):
    """ids, mask -> Linear(768, D)(last_hidden_state[:, 0]) fp32 [B, D]  (pcme.py:43-44 before l2_normalize)."""

    @staticmethod
    def forward(ctx, ids, token_type, mask, model, *params):
        bert, lin, store = model.txt_enc, model.linear, model.store()
        b, l = ids.shape
        t = b * l
        d, heads = bert.hidden, bert.heads
        need = any(ctx.needs_input_grad)
        emb = bert.embeddings
        # train-mode dropout: advance the tower's RNG step, freeze a copy for this forward/backward pair
        drop_state = None
        if bert.training and bert.dropout_p > 0.0:
            live = bert.dropout_state(ids.device)
            live.tick()
            drop_state = live.snapshot()
        ds = (lambda site: (drop_state, site)) if drop_state is not None else (lambda site: None)
        e = T.embed_fwd(ids, token_type, emb.word_embeddings.weight, emb.position_embeddings.weight,
                        emb.token_type_embeddings.weight, l)
        h, m0, r0 = T.layernorm_fwd(e, emb.LayerNorm.weight, emb.LayerNorm.bias, emb.LayerNorm.eps, drop=ds(0))
        saved_layers = []
        maskf = mask.to(torch.float32).contiguous()
        n_layers = len(bert.encoder.layer)
        for li, lyr in enumerate(bert.encoder.layer):
            s = lyr.attention.self
            wqkv, _ = store.fused([s.query.weight, s.key.weight, s.value.weight], 3 * d, d)
            bqkv_view = store.flat[store.offsets[id(s.query.bias)]:store.offsets[id(s.query.bias)] + 3 * d]
            qkv = ops.gemm_bf16(h, wqkv, bias=bqkv_view)
            ctxv, probs = T.attn_fwd(qkv, maskf, b, l, heads, drop=ds(1 + 3 * li))
            # pcme.py:44 reads last_hidden_state[:, 0] only: after the attention of the LAST layer every row-wise op
            # (output projection, LayerNorms, FFN) runs on the B [CLS] rows instead of the B*L tokens
            last = li == n_layers - 1
            a_in = ctxv.view(b, l * d)[:, :d] if last else ctxv
            r_in = h.view(b, l * d)[:, :d] if last else h
            if drop_state is not None:      # dense -> dropout -> + residual (HF BertSelfOutput)
                ao = T.gemm_drop(a_in, lyr.attention.output.dense.weight._w16, lyr.attention.output.dense.bias, r_in,
                                 ds(2 + 3 * li))
            else:
                ao = ops.gemm_bf16(a_in, lyr.attention.output.dense.weight._w16,
                                   bias=lyr.attention.output.dense.bias, add=r_in)
            ln1 = lyr.attention.output.LayerNorm
            h1, m1, r1 = T.layernorm_fwd(ao, ln1.weight, ln1.bias, ln1.eps)
            ff, pre = ops.gemm_bf16(h1, lyr.intermediate.dense.weight._w16, bias=lyr.intermediate.dense.bias,
                                    act=ops.ACT_GELU, want_preact=True)
            if drop_state is not None:      # HF BertOutput
                fo = T.gemm_drop(ff, lyr.output.dense.weight._w16, lyr.output.dense.bias, h1, ds(3 + 3 * li))
            else:
                fo = ops.gemm_bf16(ff, lyr.output.dense.weight._w16, bias=lyr.output.dense.bias, add=h1)
            ln2 = lyr.output.LayerNorm
            h2, m2, r2 = T.layernorm_fwd(fo, ln2.weight, ln2.bias, ln2.eps)
            if need:
                saved_layers.append((h, qkv, probs, ctxv, ao, m1, r1, h1, pre, ff, fo, m2, r2))
            h = h2
        cls = h                                             # [B, d]: the last layer already reduced to the [CLS] rows
        out = ops.gemm_bf16(cls, lin.weight._w16, bias=lin.bias, out_dtype=torch.float32)
        ctx.model = model
        ctx.saved = (ids, token_type, e, m0, r0, saved_layers, h, b, l, drop_state) if need else None
        return out

    @staticmethod
    def backward(ctx, dout):
        model = ctx.model
        bert, lin, store = model.txt_enc, model.linear, model.store()
        ids, token_type, e, m0, r0, saved_layers, h_last, b, l, drop_state = ctx.saved
        ctx.saved = None
        ds = (lambda site: (drop_state, site)) if drop_state is not None else (lambda site: None)
        d, heads = bert.hidden, bert.heads
        t = b * l
        dout16 = ops.to_bf16(dout.contiguous().float())
        cls = h_last                                        # [B, d]
        ops.gemm_bf16(dout16, cls, a_mn=True, b_mn=True, out=grad_target(lin.weight), split_k=0, accumulate=True)
        T.colsum_into(dout16, grad_target(lin.bias))
        dh = ops.gemm_bf16(dout16, lin.weight._w16, b_mn=True)          # [B, d] gradient at the last layer's [CLS] rows
        n_layers = len(bert.encoder.layer)
        for li, lyr, sv in zip(reversed(range(n_layers)), reversed(bert.encoder.layer), reversed(saved_layers)):
            h_in, qkv, probs, ctxv, ao, m1, r1, h1, pre, ff, fo, m2, r2 = sv
            last = li == n_layers - 1
            ln2, ln1 = lyr.output.LayerNorm, lyr.attention.output.LayerNorm
            # d_fo: gradient at the LayerNorm input (-> residual branch); d_fo_d: the same through the dropout of the
            # dense output (-> weight / bias / input gradients of the dense layer)
            d_fo = T.layernorm_bwd(dh, fo, ln2.weight, m2, r2, grad_target(ln2.weight), grad_target(ln2.bias),
                                   dx_colsum=grad_target(lyr.output.dense.bias), drop_out=ds(3 + 3 * li))
            d_fo, d_fo_d = d_fo if drop_state is not None else (d_fo, d_fo)
            ops.gemm_bf16(d_fo_d, ff, a_mn=True, b_mn=True, out=grad_target(lyr.output.dense.weight), split_k=0,
                          accumulate=True)
            d_pre = ops.gemm_bf16(d_fo_d, lyr.output.dense.weight._w16, b_mn=True, act=ops.ACT_DGELU, aux=pre)
            ops.gemm_bf16(d_pre, h1, a_mn=True, b_mn=True, out=grad_target(lyr.intermediate.dense.weight), split_k=0,
                          accumulate=True)
            T.colsum_into(d_pre, grad_target(lyr.intermediate.dense.bias))
            d_h1 = ops.gemm_bf16(d_pre, lyr.intermediate.dense.weight._w16, b_mn=True, add=d_fo)
            d_ao = T.layernorm_bwd(d_h1, ao, ln1.weight, m1, r1, grad_target(ln1.weight), grad_target(ln1.bias),
                                   dx_colsum=grad_target(lyr.attention.output.dense.bias), drop_out=ds(2 + 3 * li))
            d_ao, d_ao_d = d_ao if drop_state is not None else (d_ao, d_ao)
            a_in = ctxv.view(b, l * d)[:, :d] if last else ctxv
            ops.gemm_bf16(d_ao_d, a_in, a_mn=True, b_mn=True, out=grad_target(lyr.attention.output.dense.weight),
                          split_k=0, accumulate=True)
            if last:
                # scatter the [CLS]-row gradients back to token rows: every other row of the last layer is dead
                d_ctx = torch.zeros((t, d), dtype=BF16, device=dout.device)
                ops.gemm_bf16(d_ao_d, lyr.attention.output.dense.weight._w16, b_mn=True,
                              out=d_ctx.view(b, l * d)[:, :d])
                d_res = torch.zeros((t, d), dtype=BF16, device=dout.device)
                d_res.view(b, l * d)[:, :d].copy_(d_ao)
            else:
                d_ctx = ops.gemm_bf16(d_ao_d, lyr.attention.output.dense.weight._w16, b_mn=True)
                d_res = d_ao
            s = lyr.attention.self
            for p_ in (s.query.weight, s.key.weight, s.value.weight, s.query.bias, s.key.bias, s.value.bias):
                grad_target(p_)
            wqkv, gqkv = store.fused([s.query.weight, s.key.weight, s.value.weight], 3 * d, d)
            _, gbqkv = store.fused([s.query.bias, s.key.bias, s.value.bias], 3 * d)
            d_qkv = T.attn_bwd(qkv, probs, d_ctx, b, l, heads, dbias=gbqkv, drop=ds(1 + 3 * li))
            ops.gemm_bf16(d_qkv, h_in, a_mn=True, b_mn=True, out=gqkv, split_k=0, accumulate=True)
            dh = ops.gemm_bf16(d_qkv, wqkv, b_mn=True, add=d_res)
        emb = bert.embeddings
        de = T.layernorm_bwd(dh, e, emb.LayerNorm.weight, m0, r0, grad_target(emb.LayerNorm.weight),
                             grad_target(emb.LayerNorm.bias), drop_in=ds(0))
        T.embed_bwd(ids, token_type, de, l, grad_target(emb.word_embeddings.weight),
                    grad_target(emb.position_embeddings.weight), grad_target(emb.token_type_embeddings.weight))
        return (None, None, None, None) + (None,) * (len(ctx.needs_input_grad) - 4)


# ===================================================================================================== PCME
class PCME(StoreMixin, nn.Module):
    """Mirror of src/networks/models/pcme.py:15-57 with the BERT text tower (config.not_bert = False).

    forward(images, sentences, captions_word, lengths) returns the reference's 10-key dict.  `captions_word` is
    either the reference's tuple of strings (needs a BertTokenizer, only available when its vocabulary is on disk)
    or a pre-tokenised dict / tuple (input_ids [B,L] int64, attention_mask [B,L], optional token_type_ids)."""

    def __init__(self, word2idx, config, mlp_local=False):
        super().__init__()
        self.config = config
        get = (lambda k, dflt=None: config.get(k, dflt))
        self.embed_dim = get('embed_dim')
        self.n_embeddings = get('n_samples_inference', 0) or 1
        if get('not_bert', False):
            raise NotImplementedError('config.not_bert: build the model with get_model(), which returns the '
                                      'ResNet + GRU variant (creamfl_b200.clients.ClientPCME)')
        self.img_enc = EncoderImage(config, mlp_local)
        # HF BertConfig defaults (0.1); `bert_dropout: 0` gives the frozen-dropout parity protocol of SURVEY 3.2
        self.txt_enc = BertEncoder(dropout_p=get('bert_dropout', 0.1))
        self.linear = _Linear(768, self.embed_dim)
        self.tokenizer = None

    def _adjacent_groups(self):
        return self.txt_enc.qkv_groups()

    def _tokens(self, captions_word):
        if isinstance(captions_word, dict):
            ids, mask, tt = captions_word['input_ids'], captions_word['attention_mask'], captions_word.get(
                'token_type_ids')
        elif isinstance(captions_word, (tuple, list)) and len(captions_word) and torch.is_tensor(captions_word[0]):
            ids, mask = captions_word[0], captions_word[1]
            tt = captions_word[2] if len(captions_word) > 2 else None
        else:
            if self.tokenizer is None:
                from transformers import BertTokenizer
                self.tokenizer = BertTokenizer.from_pretrained('bert-base-uncased')
            enc = self.tokenizer(list(captions_word), padding=True, return_tensors='pt')
            ids, mask, tt = enc['input_ids'], enc['attention_mask'], enc['token_type_ids']
        dev = self.linear.weight.device
        ids = ids.to(dev, non_blocking=True).long().contiguous()
        mask = mask.to(dev, non_blocking=True)
        tt = torch.zeros_like(ids) if tt is None else tt.to(dev, non_blocking=True).long().contiguous()
        return ids, tt, mask

    def text_forward(self, captions_word):
        ids, tt, mask = self._tokens(captions_word)
        params = [p for p in self.txt_enc.parameters()] + [self.linear.weight, self.linear.bias]
        out = _BertFn.apply(ids, tt, mask, self, *params)
        return {'embedding': ops.l2_normalize(out)}

    def image_forward(self, images):
        self.store()
        return self.img_enc(images)

    def forward(self, images, sentences, captions_word, lengths):
        self.store()
        # the two towers share nothing until the loss: the text tower (tensor-bound GEMMs) is forked onto a side
        # stream so that it fills the SMs the image tower's HBM- / latency-bound kernels leave idle; autograd replays
        # each tower's backward on its forward stream, the towers write disjoint regions of the flat gradient buffer
        with fork_stream(images.device, getattr(self, 'overlap_towers', True)) as side:
            with side:
                caption_output = self.text_forward(captions_word)
            image_output = self.img_enc(images)
        return {
            'image_features': image_output['embedding'],
            'image_attentions': image_output.get('attention'),
            'image_residuals': image_output.get('residual'),
            'image_logsigma': image_output.get('logsigma'),
            'image_logsigma_att': image_output.get('uncertainty_attention'),
            'caption_features': caption_output['embedding'],
            'caption_attentions': caption_output.get('attention'),
            'caption_residuals': caption_output.get('residual'),
            'caption_logsigma': caption_output.get('logsigma'),
            'caption_logsigma_att': caption_output.get('uncertainty_attention'),
        }


class ImageModel(StoreMixin, nn.Module):
    """Stand-alone image tower (EncoderImage with its own parameter store)."""

    def __init__(self, config):
        super().__init__()
        self.img_enc = EncoderImage(config)

    def forward(self, images):
        self.store()
        return self.img_enc(images)['embedding']


def get_model(word2idx, config, mlp_local=False):
    """Mirror of src/networks/models/__init__.py:6-7.  config.not_bert selects the GRU text tower
    (pcme.py:28-29,37-38; with ResNet50 when it is the server, MMFL.py:82-85)."""
    get = config.get if hasattr(config, 'get') else (lambda k, d=None: getattr(config, k, d))
    if get('not_bert', False):
        if mlp_local:
            raise NotImplementedError('mlp_local heads are hard-wired to 512-d in the reference (SURVEY appendix B)')
        from .clients import ClientPCME
        vocab = len(word2idx) if word2idx else get('vocab_size', 11755)
        return ClientPCME(vocab, get('embed_dim'), cnn_type=get('cnn_type', 'resnet18'), word_dim=get('word_dim', 300))
    return PCME(word2idx, config, mlp_local)
