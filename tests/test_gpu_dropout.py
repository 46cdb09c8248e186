"""GPU parity of the fused BERT dropout (include/creamfl_b200.h "BERT dropout"; reference: HF BertConfig defaults
p = 0.1 under model.train(), src/networks/models/pcme.py:31, src/algorithms/retrieval_trainer.py:187).

The kernels never store a mask; `creamfl_dropout_mask` exports the mask they regenerate.  Tests:
  * the exported mask equals the numpy Philox4x32-10 restatement (tests/kernel_emulation.py, pinned to the Random123
    known-answer vectors by tests/test_cpu_dropout.py) - bit-exact;
  * each fused kernel (GEMM epilogue, LayerNorm fwd/bwd, attention fwd/bwd) against torch arithmetic using the
    exported mask: bf16 outputs rel-L2 4e-3, fp32 column sums 1e-3;
  * the whole BERT tower with dropout on against HF BertModel fed the SAME masks (oracle.torch_towers.frozen_dropout):
    embeddings cos >= 0.9995 per row, gradients calibrated against torch's own bf16 autocast like test_gpu_towers;
  * a captured CUDA graph of the server step draws a new mask at every replay; eval mode applies none."""
import numpy as np
import pytest
import torch

import kernel_emulation as KE

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


@pytest.fixture(scope='module')
def T():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from creamfl_b200 import tower_ops
    return tower_ops


def rel(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-300))


def _state(T, seed=77, p=0.1, ticks=3):
    st = T.DropoutState(seed, p, torch.device('cuda'))
    for _ in range(ticks):
        st.tick()
    return st


@pytest.mark.parametrize('n,site,p', [(1000, 0, 0.1), (12 * 32 * 32 * 5 + 3, 7, 0.1), (4096 * 768, 36, 0.1), (777, 2, 0.5)])
def test_exported_mask_is_philox(T, n, site, p):
    st = _state(T, seed=2 ** 41 + 5, p=p, ticks=4)
    got = st.keep_mask(site, n).cpu().numpy()
    assert int(st.rng[1]) == 4
    want = KE.keep_mask_np(2 ** 41 + 5, 4, site, n, p)
    assert np.array_equal(got, want)


@pytest.mark.parametrize('m,n,k,strided', [(4096, 768, 768, False), (4096, 768, 3072, False), (128, 768, 768, True),
                                           (100, 768, 768, False)])
def test_gemm_epilogue_dropout(T, m, n, k, strided):
    g = torch.Generator().manual_seed(1)
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(BF16).cuda()
    bias = torch.randn(n, generator=g).cuda()
    if strided:      # the last layer's [CLS] rows: row pitch L * d
        big_a = torch.randn(m, 4 * k, generator=g).to(BF16).cuda()
        big_r = torch.randn(m, 4 * n, generator=g).to(BF16).cuda()
        a, r = big_a[:, :k], big_r[:, :n]
    else:
        a, r = torch.randn(m, k, generator=g).to(BF16).cuda(), torch.randn(m, n, generator=g).to(BF16).cuda()
    st = _state(T)
    out = T.gemm_drop(a, w, bias, r, (st, 5))
    keep = st.keep_mask(5, m * n).view(m, n).double()
    want = (a.double() @ w.double().t() + bias.double()) * keep / 0.9 + r.double()
    assert rel(out, want) < 4e-3
    # dropped positions hold exactly the residual
    dropped = keep == 0
    assert torch.equal(out[dropped], r[dropped])


def test_layernorm_dropout_fwd_bwd(T):
    g = torch.Generator().manual_seed(2)
    r_, d = 4096, 768
    x = torch.randn(r_, d, generator=g).to(BF16).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(d, generator=g)).cuda(), (0.1 * torch.randn(d, generator=g)).cuda()
    dy = torch.randn(r_, d, generator=g).to(BF16).cuda()
    st = _state(T)
    y, mean, rstd = T.layernorm_fwd(x, gamma, beta, 1e-12, drop=(st, 0))
    k0 = st.keep_mask(0, r_ * d).view(r_, d).double()
    xd = x.double().requires_grad_(True)
    yref = torch.nn.functional.layer_norm(xd, (d,), gamma.double(), beta.double(), 1e-12) * k0 / 0.9
    assert rel(y, yref) < 4e-3
    # backward with the input mask (embeddings) and the output mask (dense under the LayerNorm) at once
    dgamma, dbeta, colsum = (torch.zeros(d, device='cuda') for _ in range(3))
    dx, dxd = T.layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, dx_colsum=colsum, drop_in=(st, 0),
                              drop_out=(st, 9))
    gam = gamma.double().requires_grad_(True)
    bet = beta.double().requires_grad_(True)
    yln = torch.nn.functional.layer_norm(xd, (d,), gam, bet, 1e-12)
    (yln * (dy.double() * k0 / 0.9)).sum().backward()
    k9 = st.keep_mask(9, r_ * d).view(r_, d).double()
    assert rel(dx, xd.grad) < 4e-3
    assert rel(dxd, xd.grad * k9 / 0.9) < 4e-3
    assert rel(dgamma, gam.grad) < 1e-3 and rel(dbeta, bet.grad) < 1e-3
    assert rel(colsum, (xd.grad * k9 / 0.9).sum(0)) < 2e-3


@pytest.mark.parametrize('b,l', [(16, 32), (5, 27), (3, 64)])
def test_attention_dropout_fwd_bwd(T, b, l):
    heads = 12
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(b * l, 3 * heads * 64, generator=g).to(BF16).cuda()
    lens = torch.randint(max(1, l // 2), l + 1, (b,), generator=g)
    lens[0] = l
    mask = (torch.arange(l)[None] < lens[:, None]).float().cuda()
    dctx = torch.randn(b * l, heads * 64, generator=g).to(BF16).cuda()
    st = _state(T)
    ctx, probs = T.attn_fwd(qkv, mask, b, l, heads, drop=(st, 4))
    keep = st.keep_mask(4, b * heads * l * l).view(b, heads, l, l).double()
    q, k, v = qkv.double().view(b, l, 3, heads, 64).permute(2, 0, 3, 1, 4)
    q, k, v = (t.contiguous().requires_grad_(True) for t in (q, k, v))
    s = q @ k.transpose(-1, -2) * 0.125 + torch.where(mask > 0.5, 0.0, -3.0e38)[:, None, None, :].double()
    p = torch.softmax(s, dim=-1)
    want = ((p * keep / 0.9) @ v).permute(0, 2, 1, 3).reshape(b * l, heads * 64)
    assert rel(probs, p) < 4e-3                      # the saved probabilities are the un-dropped ones
    assert rel(ctx, want) < 5e-3
    dbias = torch.zeros(3 * heads * 64, device='cuda')
    dqkv = T.attn_bwd(qkv, probs, dctx, b, l, heads, dbias=dbias, drop=(st, 4))
    (want * dctx.double()).sum().backward()
    ref = torch.stack([q.grad, k.grad, v.grad], 0).permute(1, 3, 0, 2, 4).reshape(b * l, 3 * heads * 64)
    assert rel(dqkv, ref) < 8e-3


def _provider(state_seed, p, step, batch, layers):
    def provider(call, shape):
        site = call
        if site in (2 + 3 * (layers - 1), 3 + 3 * (layers - 1)):        # last layer: [CLS] rows only in the product
            full = torch.ones(shape)
            small = torch.from_numpy(KE.keep_mask_np(state_seed, step, site, batch * shape[-1], p))
            full[:, 0, :] = small.view(batch, shape[-1]).float()
            return full
        n = int(np.prod(shape))
        return torch.from_numpy(KE.keep_mask_np(state_seed, step, site, n, p)).view(*shape).float()
    return provider


def test_bert_tower_with_dropout_matches_hf_on_same_masks(T):
    import copy
    from transformers import BertConfig
    from creamfl_b200 import towers
    from oracle import torch_towers as RT
    layers, batch, seq = 12, 8, 32
    ref = RT.RefPCME('resnet18', 256, BertConfig(num_hidden_layers=layers), bert_dropout=0.1)
    RT.fill_deterministic(ref, seed=31)
    ref = ref.cuda().train()
    mine = towers.PCME(None, {'embed_dim': 256, 'cnn_type': 'resnet18', 'bert_dropout': 0.1})
    mine.txt_enc = towers.BertEncoder(layers=layers, dropout_p=0.1, seed=4242)
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda().train()
    g = torch.Generator().manual_seed(32)
    ids = torch.randint(1000, 30522, (batch, seq), generator=g)
    lens = torch.randint(6, seq + 1, (batch,), generator=g)
    lens[0] = seq
    mask = (torch.arange(seq)[None] < lens[:, None]).long()
    ids[:, 0] = 101
    ids, mask = (ids * mask).cuda(), mask.cuda()
    cot = torch.randn(batch, 256, generator=g).cuda()
    prov = _provider(4242, 0.1, 1, batch, layers)

    def run_ref(model):
        with RT.frozen_dropout(prov) as fd:
            hid = model.txt_enc(input_ids=ids, attention_mask=mask, token_type_ids=torch.zeros_like(ids))
            assert fd.calls == 1 + 3 * layers
        return RT.l2_normalize(model.linear(hid['last_hidden_state'][:, 0, :]))
    e_ref = run_ref(ref)
    (e_ref * cot).sum().backward()
    amp = copy.deepcopy(ref)
    amp.zero_grad()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        e_amp = run_ref(amp)
    (e_amp.float() * cot).sum().backward()
    mine.zero_grad()
    e = mine.text_forward({'input_ids': ids, 'attention_mask': mask})['embedding']
    (e * cot).sum().backward()
    torch.cuda.synchronize()
    for i in range(batch):
        assert cos(e[i], e_ref[i]) >= 0.9995, (i, cos(e[i], e_ref[i]))
    # sanity: the dropout-free forward is a measurably different function
    mine.eval()
    with torch.no_grad():
        e0 = mine.text_forward({'input_ids': ids, 'attention_mask': mask})['embedding']
    assert min(cos(e0[i], e_ref[i]) for i in range(batch)) < 0.999
    ref_p, mine_p, amp_p = dict(ref.named_parameters()), dict(mine.named_parameters()), dict(amp.named_parameters())
    names = ['linear.weight', 'linear.bias', 'txt_enc.encoder.layer.11.output.dense.weight',
             'txt_enc.encoder.layer.11.output.dense.bias', 'txt_enc.encoder.layer.11.attention.output.dense.weight',
             'txt_enc.encoder.layer.11.attention.self.value.weight', 'txt_enc.encoder.layer.6.attention.self.value.bias',
             'txt_enc.encoder.layer.6.intermediate.dense.weight', 'txt_enc.encoder.layer.6.output.dense.bias',
             'txt_enc.encoder.layer.6.attention.output.dense.bias', 'txt_enc.encoder.layer.3.attention.output.LayerNorm.weight',
             'txt_enc.encoder.layer.0.attention.self.key.weight', 'txt_enc.encoder.layer.0.attention.self.query.weight',
             'txt_enc.encoder.layer.0.output.LayerNorm.bias', 'txt_enc.embeddings.LayerNorm.weight',
             'txt_enc.embeddings.position_embeddings.weight', 'txt_enc.embeddings.word_embeddings.weight']
    bad = []
    for n in names:
        c, c_amp = cos(mine_p[n].grad, ref_p[n].grad), cos(amp_p[n].grad, ref_p[n].grad)
        ratio = float(mine_p[n].grad.double().norm() / ref_p[n].grad.double().norm())
        if not (c >= c_amp - 0.02 and abs(ratio - 1) <= 0.10):
            bad.append((n, round(c, 4), round(c_amp, 4), round(ratio, 4)))
    assert not bad, bad


def test_graph_replay_draws_new_masks_and_eval_is_deterministic(T):
    from creamfl_b200 import engine
    server = engine.ServerEngine(64, 'resnet18', use_graphs=True, bert_dropout=0.1)
    g = torch.Generator().manual_seed(40)
    images = torch.randn(4, 3, 224, 224, generator=g).cuda()
    ids = torch.randint(1000, 30522, (4, 16), generator=g).cuda()
    mask = torch.ones(4, 16, dtype=torch.long).cuda()
    tok = {'input_ids': ids, 'attention_mask': mask}
    rng = server.model.txt_enc.dropout_state(server.device).rng
    for k in range(3):
        server.train_step(images, tok)
        assert int(rng[1]) == k + 1            # warm-up run of the capture was rolled back; one tick per step
    a, _ = server.extract(images, tok)
    a = a.clone()
    b, _ = server.extract(images, tok)
    assert torch.equal(a, b) and int(rng[1]) == 3
