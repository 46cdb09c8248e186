"""Host-side sequencing of the GRU text towers (creamfl_b200/text_towers.py) checked WITHOUT a GPU.

The ctypes wrappers are swapped for torch emulations that follow the C ABI's documented semantics
(tests/kernel_emulation.py); everything else - the ParamStore layout with padded K = 300 operands, the fused
[6H, 304] input-projection operand, the order and layouts of the GEMM calls, the gradient accumulation targets - is
the product code.  Compared against the golden vectors produced by the REFERENCE's own caption_encoder.EncoderText /
language_model.EncoderText (tests/golden/text_towers.npz).

Tolerance: operands and saved activations are bf16 (8 mantissa bits) as in the CUDA path: embeddings cos >= 0.9995,
gradients rel-L2 <= 3e-2 (a wrong transpose, view or missing term gives O(1)).
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import kernel_emulation as KE  # tests/ is on sys.path (rootdir conftest, rootless test dir)


@pytest.fixture()
def emu(monkeypatch):
    KE.install(monkeypatch)


def _load(name):
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / 'golden' / f'{name}.npz', allow_pickle=False)


def _rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_mm_text_tower_against_reference_golden(emu):
    from creamfl_b200.text_towers import TextModel
    from oracle.torch_towers import fill_deterministic
    g = _load('text_towers')
    x, lengths, coef = torch.from_numpy(g['x']), torch.from_numpy(g['lengths']), torch.from_numpy(g['coef'])
    model = TextModel(500, 300, 64)
    fill_deterministic(model.txt_enc, seed=41)
    st = model.store()
    st.sync_shadow()
    # layout facts the kernels rely on
    tw = model.txt_enc
    assert tw._wih16.shape == (6 * 32, 304) and tw._wih16.stride() == (304, 1)
    assert tw.pie_net.attention.w_1.weight._w16p.shape == (152, 304)
    assert float(tw._wih16[:, 300:].abs().max()) == 0.0
    assert torch.equal(tw._wih16[:96, :300].float(), tw.rnn.weight_ih_l0.data.to(torch.bfloat16).float())
    assert torch.equal(tw._wih16[96:, :300].float(), tw.rnn.weight_ih_l0_reverse.data.to(torch.bfloat16).float())
    assert tw._whh.data_ptr() == tw.rnn.weight_hh_l0.data_ptr()
    assert tw._whh[1].data_ptr() == tw.rnn.weight_hh_l0_reverse.data_ptr()
    assert tw._bih[96:].data_ptr() == tw.rnn.bias_ih_l0_reverse.data_ptr()
    model.train()
    emb = model(x, lengths)
    cos = F.cosine_similarity(emb.detach().double(), torch.from_numpy(g['mm_embedding']).double(), dim=-1)
    assert float(cos.min()) > 0.9995, cos
    st.zero_grad()
    (emb * coef).sum().backward()
    for name, p in model.txt_enc.named_parameters():
        ref = g['mm_grad.' + name]
        if np.abs(ref).max() == 0.0:
            assert float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad.data_ptr() == p._gview.data_ptr()
        assert _rel(p.grad, ref) < 3e-2, (name, _rel(p.grad, ref))
    # no-grad / eval forward takes the same path minus the saved activations
    with torch.no_grad():
        emb2 = model(x, lengths)
    assert torch.allclose(emb2, emb.detach(), atol=1e-6)


def test_unimodal_text_client_against_reference_golden(emu):
    from creamfl_b200.clients import TextClient, _GramFn
    from oracle.torch_towers import fill_deterministic
    g = _load('text_towers')
    x, lengths, labels = torch.from_numpy(g['x']), torch.from_numpy(g['lengths']), torch.from_numpy(g['uni_labels'])
    client = TextClient(int(g['uni_vocab']), 300, 64, num_class=4, scale=128)
    fill_deterministic(client, seed=43)
    st = client.store()
    st.sync_shadow()
    client.train()
    st.zero_grad()
    x1, x2, w1, w2 = client(x, lengths)
    assert float(w1.min()) >= 0.0 and float(w2.min()) >= 0.0            # in-place ReLU clamp (language_model.py:115-121)
    assert _rel(x1.detach(), g['uni_x1']) < 2e-2 and _rel(x2.detach(), g['uni_x2']) < 2e-2
    onehot = F.one_hot(labels, 4).float()
    loss = F.cross_entropy(x1 - 4.0 * onehot, labels) + 0.5 * F.cross_entropy(_GramFn.apply(w1), torch.arange(4))
    assert abs(float(loss) - float(g['uni_loss'])) < 2e-2 * abs(float(g['uni_loss']))
    loss.backward()
    params = dict(client.named_parameters())
    # `* 128` puts the logits in the hundreds: a 2^-9 relative rounding of the bf16 head operand moves them by O(1) and
    # the softmax with them, so the gradients carry ~5 % noise (with fp32 operands the same code matches to 1e-4)
    for name in ['rnn.weight_hh_l0', 'rnn.weight_ih_l0_reverse', 'pie_net.fc.weight', 'class_fc.weight', 'class_fc.bias']:
        assert _rel(params[name].grad, g['uni_grad.' + name]) < 1e-1, (name, _rel(params[name].grad, g['uni_grad.' + name]))
    assert _rel(params['embed.weight'].grad[:500], g['uni_grad_embed_rows']) < 1e-1
    client.is_train = False
    with torch.no_grad():
        emb = client(x, lengths)
    cos = F.cosine_similarity(emb.double(), torch.from_numpy(g['uni_embedding']).double(), dim=-1)
    assert float(cos.min()) > 0.9995


def test_text_tower_refuses_cpu_without_emulation():
    """The product has no CPU path: without the test-only emulation the store refuses CPU parameters."""
    from creamfl_b200.text_towers import TextModel
    model = TextModel(50, 300, 64)
    with pytest.raises(RuntimeError, match='CUDA only'):
        model(torch.zeros(2, 3, dtype=torch.long), torch.tensor([3, 2]))


def test_deepcopy_rebinds_to_its_own_store(emu):
    import copy
    from creamfl_b200.text_towers import TextModel
    model = TextModel(50, 300, 64)
    model.store()
    clone = copy.deepcopy(model)
    assert clone.txt_enc._bound is None
    x, lengths = torch.randint(0, 50, (3, 5)), torch.tensor([5, 3, 1])
    a, b = model(x, lengths), clone(x, lengths)
    assert torch.equal(a, b)
    assert clone.txt_enc._whh.data_ptr() != model.txt_enc._whh.data_ptr()
    assert clone.txt_enc._whh.data_ptr() == clone.txt_enc.rnn.weight_hh_l0.data_ptr()


def test_mm_text_tower_exact_mode(monkeypatch):
    """Same sequencing with the "bf16" tensors stored in fp32: every parameter gradient of the GRU text tower agrees
    with the fp64 oracle to 1e-4 (a wrong operand, transpose or accumulation target would be O(1))."""
    import copy
    KE.install(monkeypatch, exact=True)
    from creamfl_b200.text_towers import TextModel
    from oracle.torch_towers import RefGRUEncoderText, fill_deterministic
    g = torch.Generator().manual_seed(31)
    vocab, b, l = 200, 7, 11
    lengths = torch.tensor([11, 11, 8, 5, 3, 2, 1])
    x = torch.randint(1, vocab, (b, l), generator=g)
    coef = torch.randn(b, 128, generator=g)
    ref = RefGRUEncoderText(vocab, 300, 128)
    fill_deterministic(ref, seed=32)
    model = TextModel(vocab, 300, 128)
    model.txt_enc.load_state_dict(ref.state_dict(), strict=True)
    ref64 = copy.deepcopy(ref).double()
    e64 = ref64(x, lengths)['embedding']
    (e64 * coef.double()).sum().backward()
    st = model.store()
    st.zero_grad()
    model.train()
    emb = model(x, lengths)
    assert _rel(emb.detach(), e64.detach()) < 1e-5
    (emb * coef).sum().backward()
    p64 = dict(ref64.named_parameters())
    for name, p in model.txt_enc.named_parameters():
        r = p64[name].grad
        if r is None or float(r.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) == 0.0, name
        else:
            assert _rel(p.grad, r) < 1e-4, (name, _rel(p.grad, r))
