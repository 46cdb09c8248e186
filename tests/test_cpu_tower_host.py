"""Host-side sequencing of the image / BERT towers (creamfl_b200/towers.py, clients.py) checked WITHOUT a GPU.

The ctypes wrappers are swapped for torch emulations of the C ABI's documented semantics (tests/kernel_emulation.py);
the product code under test is everything above them: the flat ParamStore (channels-last filters, fused q/k/v
operand, padded stem filter), the per-block autograd nodes and their forward / backward kernel sequences, the
gradient accumulation targets.  Oracle: oracle/torch_towers.py (pinned against the reference's own modules by
tests/golden/towers.npz).

exact mode stores the "bf16" tensors in fp32, so a wrong operand, transpose, missing term or accumulation target
shows up as O(1) against a 1e-3 bound; the bf16 run keeps the TMA alignment assertions of the emulated GEMM armed
and is compared at the bf16 noise floor.
"""
import pytest
import torch
import torch.nn.functional as F

import kernel_emulation as KE  # tests/ is on sys.path (rootdir conftest, rootless test dir)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _pcme_pair(embed_dim=64, layers=2, seed=5, dropout=0.0):
    """dropout = 0: the frozen-dropout parity protocol; tests/test_cpu_dropout.py covers 0.1."""
    from transformers import BertConfig
    from creamfl_b200 import towers
    from oracle import torch_towers as RT
    ref = RT.RefPCME('resnet18', embed_dim, BertConfig(num_hidden_layers=layers), bert_dropout=dropout)
    RT.fill_deterministic(ref, seed=seed)
    mine = towers.PCME(None, {'embed_dim': embed_dim, 'cnn_type': 'resnet18', 'bert_dropout': dropout})
    mine.txt_enc = towers.BertEncoder(layers=layers, dropout_p=dropout, seed=1234)
    mine.load_state_dict(ref.state_dict(), strict=True)
    return ref.train(), mine.train()


def _inputs(batch=3, seq=8, size=224, seed=6, dim=64):   # the reference hard-wires the 7x7 map (image_encoder.py:55)
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, size, size, generator=g)
    ids = torch.randint(1000, 30522, (batch, seq), generator=g)
    lens = torch.tensor([seq, 5, 3][:batch])
    mask = (torch.arange(seq)[None] < lens[:, None]).long()
    ids[:, 0] = 101
    return images, ids * mask, mask, torch.randn(batch, dim, generator=g), torch.randn(batch, dim, generator=g)


def _grads_vs_fp64(mine, ref32, ref64, skip=()):
    """Per-parameter gradient error of the product path and of torch's own fp32 run, both against the fp64 oracle.
    BatchNorm backward over a 3-image batch is ill-conditioned (torch fp32 itself sits at 3e-3 .. 1e-2 through the
    ResNet trunk), so the bound is relative to that: a sequencing error is O(1), rounding differences are not."""
    p64, p32 = dict(ref64.named_parameters()), dict(ref32.named_parameters())
    bad = []
    for name, p in mine.named_parameters():
        g64 = p64[name].grad
        if g64 is None or float(g64.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, name            # e.g. the BERT pooler: dead on this path (pcme.py:44)
            continue
        if any(name.endswith(sfx) for sfx in skip):
            continue
        assert p.grad.data_ptr() == p._gview.data_ptr(), name         # accumulated into the flat gradient buffer
        e_mine, e_torch = _rel(p.grad, g64), _rel(p32[name].grad, g64)
        if e_mine > 2.0 * e_torch + 2e-4:
            bad.append((name, e_mine, e_torch))
    assert not bad, bad


def test_pcme_train_step_exact(monkeypatch):
    import copy
    KE.install(monkeypatch, exact=True)
    ref, mine = _pcme_pair()
    images, ids, mask, cot_i, cot_t = _inputs()
    ref64 = copy.deepcopy(ref).double()
    o64 = ref64(images.double(), ids, mask, torch.zeros_like(ids))
    ((o64['image_features'] * cot_i.double()).sum() + (o64['caption_features'] * cot_t.double()).sum()).backward()
    o_ref = ref(images, ids, mask, torch.zeros_like(ids))
    ((o_ref['image_features'] * cot_i).sum() + (o_ref['caption_features'] * cot_t).sum()).backward()
    st = mine.store()
    st.zero_grad()
    o = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
    assert _rel(o['image_features'], o64['image_features']) < 1e-4
    assert _rel(o['caption_features'], o64['caption_features']) < 1e-4
    ((o['image_features'] * cot_i).sum() + (o['caption_features'] * cot_t).sum()).backward()
    # the key bias has no gradient mathematically (softmax rows of dS sum to zero): both sides hold rounding noise
    _grads_vs_fp64(mine, ref, ref64, skip=('attention.self.key.bias',))
    # BatchNorm side effects of a training forward (running statistics, nn.BatchNorm2d's counter)
    ref_b, mine_b = dict(ref.named_buffers()), dict(mine.named_buffers())
    for name in ['img_enc.cnn.bn1.running_mean', 'img_enc.cnn.layer3.0.downsample.1.running_var',
                 'img_enc.cnn.layer4.1.bn2.running_mean']:
        assert _rel(mine_b[name], ref_b[name]) < 1e-4, name
    assert int(mine_b['img_enc.cnn.layer2.0.bn1.num_batches_tracked']) == 1
    # a second backward accumulates into the same buffers
    before = st.grad.clone()
    o = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
    ((o['image_features'] * cot_i).sum() + (o['caption_features'] * cot_t).sum()).backward()
    assert _rel(st.grad, 2 * before) < 1e-5      # running BN statistics do not enter a training-mode forward


def test_pcme_eval_forward_exact(monkeypatch):
    KE.install(monkeypatch, exact=True)
    ref, mine = _pcme_pair(seed=7)
    images, ids, mask, _, _ = _inputs(seed=8)
    ref.eval()
    mine.eval()
    with torch.no_grad():
        o_ref = ref(images, ids, mask, torch.zeros_like(ids))
        o = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
    assert _rel(o['image_features'], o_ref['image_features']) < 1e-4
    assert _rel(o['caption_features'], o_ref['caption_features']) < 1e-4
    # the inference path folds every block BatchNorm into its convolution (towers._EvalFold); the un-folded path
    # (BatchNorm as its own pass) gives the same features, and the fold follows the parameters / running statistics
    from creamfl_b200 import towers
    assert mine.img_enc.cnn.__dict__.get('_fold') is not None
    monkeypatch.setattr(towers.ResNet, 'fold_eval_bn', False)
    with torch.no_grad():
        o_plain = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
    assert _rel(o['image_features'], o_plain['image_features']) < 1e-5
    monkeypatch.setattr(towers.ResNet, 'fold_eval_bn', True)
    with torch.no_grad():
        blk = mine.img_enc.cnn.layer2[0]
        blk.bn1.running_var.mul_(1.7)
        blk.conv2.weight.mul_(0.6)
        blk.bn2.bias.add_(0.05)
        ref.img_enc.cnn.layer2[0].bn1.running_var.mul_(1.7)
        ref.img_enc.cnn.layer2[0].conv2.weight.mul_(0.6)
        ref.img_enc.cnn.layer2[0].bn2.bias.add_(0.05)
        o2 = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
        o2_ref = ref(images, ids, mask, torch.zeros_like(ids))
    assert _rel(o2['image_features'], o2_ref['image_features']) < 1e-4
    assert _rel(o2['image_features'], o['image_features']) > 1e-3


def test_pcme_bf16_layout_rules(monkeypatch):
    """bf16 storage: every GEMM operand the towers hand to the library satisfies the TMA rules the emulation asserts
    (16-byte aligned base and row pitch), and the result sits at the bf16 noise floor of the oracle."""
    KE.install(monkeypatch, exact=False)
    ref, mine = _pcme_pair(seed=9)
    images, ids, mask, cot_i, cot_t = _inputs(seed=10)
    with torch.no_grad():
        o_ref = ref(images, ids, mask, torch.zeros_like(ids))
    mine.store().zero_grad()
    o = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
    ((o['image_features'] * cot_i).sum() + (o['caption_features'] * cot_t).sum()).backward()
    cos_t = F.cosine_similarity(o['caption_features'].double(), o_ref['caption_features'].double(), dim=-1)
    cos_i = F.cosine_similarity(o['image_features'].double(), o_ref['image_features'].double(), dim=-1)
    assert float(cos_t.min()) > 0.999 and float(cos_i.min()) > 0.98, (cos_t, cos_i)
    w = mine.img_enc.cnn.conv1.weight
    assert w._w16.shape == (64, 152) and w._w16.dtype == torch.bfloat16 and float(w._w16[:, 147:].abs().max()) == 0.0
    assert torch.isfinite(mine.store().grad).all() and float(mine.store().grad.abs().max()) > 0


def test_image_client_supervised_step_exact(monkeypatch):
    KE.install(monkeypatch, exact=True)
    from creamfl_b200 import clients
    from oracle import torch_towers as RT
    ref = RT.RefImageClient(num_class=10, embed_dim=64)
    RT.fill_deterministic(ref, seed=21)
    with torch.no_grad():
        ref.linear.weight.mul_(0.05)
    mine = clients.resnet18_client(num_class=10, embed_dim=64, scale=128, is_train=True)
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref.train()
    mine.train()
    g = torch.Generator().manual_seed(24)     # (seed 22 puts one layer4 pre-activation at 4.5e-6: its ReLU gate flips in fp32)
    images = torch.randn(4, 3, 64, 64, generator=g)
    labels = torch.randint(0, 10, (4,), generator=g)
    import copy
    ref64 = copy.deepcopy(ref).double()
    l64, _ = RT.ref_unimodal_supervised_loss(ref64, images.double(), labels, 10)
    l64.backward()
    l_ref, _ = RT.ref_unimodal_supervised_loss(ref, images, labels, 10)
    l_ref.backward()
    mine.store().zero_grad()
    l, fvec = clients.unimodal_supervised_loss(mine, images, labels, 4.0)
    l.backward()
    assert abs(float(l) - float(l64)) < 1e-4 * abs(float(l64))
    _grads_vs_fp64(mine, ref, ref64)
    assert float(mine.class_fc_2.weight.data.min()) >= 0.0           # ReLU clamp is a side effect of forward
