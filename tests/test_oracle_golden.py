"""CPU tests: the oracle (oracle/creamfl_oracle.py) against golden vectors produced by the reference's own code
(tests/golden/make_golden.py).  These pin the oracle; the GPU parity tests then pin the CUDA path to the oracle."""
import numpy as np
import pytest
import torch

from oracle import creamfl_oracle as O
from conftest import conw_inputs

T = lambda a: torch.from_numpy(np.asarray(a))


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_pcme_loss_and_grads(golden, tag):
    g = golden('pcme')
    img = T(g[f'{tag}_img']).requires_grad_(True)
    txt = T(g[f'{tag}_txt']).requires_grad_(True)
    shift = torch.tensor(float(g[f'{tag}_shift']), dtype=torch.float64, requires_grad=True)
    scale = torch.tensor(float(g[f'{tag}_scale']), dtype=torch.float64, requires_grad=True)
    loss, info = O.pcme_loss(img, txt, shift, scale)
    loss.backward()
    assert loss.item() == pytest.approx(float(g[f'{tag}_loss']), rel=1e-12)
    assert info['i2t_pos_loss'] == pytest.approx(float(g[f'{tag}_i2t_pos']), rel=1e-12)
    assert info['i2t_neg_loss'] == pytest.approx(float(g[f'{tag}_i2t_neg']), rel=1e-12)
    assert info['t2i_loss'] == pytest.approx(float(g[f'{tag}_t2i_loss']), rel=1e-12)
    np.testing.assert_allclose(img.grad.numpy(), g[f'{tag}_d_img'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(txt.grad.numpy(), g[f'{tag}_d_txt'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(shift.grad.numpy(), g[f'{tag}_d_shift'][0], rtol=1e-9)
    np.testing.assert_allclose(scale.grad.numpy(), g[f'{tag}_d_scale'][0], rtol=1e-9)


def test_pcme_closed_form_identity(golden):
    """loss = 2 * sum softplus(-2 m (-s d + b))."""
    g = golden('pcme')
    img, txt = T(g['a_img']), T(g['a_txt'])
    d = O.pcme_pair_distance(img, txt)
    m = 2 * torch.eye(len(img), dtype=torch.float64) - 1
    l = -float(g['a_scale']) * d + float(g['a_shift'])
    closed = 2 * torch.nn.functional.softplus(-2 * m * l).sum()
    assert closed.item() == pytest.approx(float(g['a_loss']), rel=1e-12)


def test_pcme_rejects_ragged():
    with pytest.raises(RuntimeError):
        O.pcme_loss(torch.zeros(3, 4), torch.zeros(2, 4), torch.tensor(1.0), torch.tensor(1.0))


@pytest.mark.parametrize('tag', ['both', 'intra', 'inter', 'both_scaled'])
def test_mm_client_contrast(golden, tag):
    g = golden('mm_contrast')
    rows = g['rows']
    oi = T(g['cur_img'][rows]).requires_grad_(True)
    ot = T(g['cur_txt'][rows]).requires_grad_(True)
    old_i, old_t = T(g['old_img'][rows]), T(g['old_txt'][rows])
    g_img, g_txt = T(g['g_img']), T(g['g_txt'])
    lookup = {int(b): a for a, b in enumerate(g['distill_index'])}
    d_idx = [lookup[int(i)] for i in g['batch_index']]
    assert d_idx == list(rows)
    parts = O.mm_client_contrast_loss(oi, ot, old_i, old_t, g_img, g_txt, d_idx, 0.5, tag == 'both_scaled')
    if tag == 'intra':
        loss = parts['intra']          # MMClientTrainer.py:264 - no interintra_weight on this branch
    elif tag == 'inter':
        loss = parts['inter']          # MMClientTrainer.py:308
    else:
        loss = parts['loss']
    loss.backward()
    assert loss.item() == pytest.approx(float(g[f'{tag}_loss']), rel=1e-12)
    np.testing.assert_allclose(oi.grad.numpy(), g[f'{tag}_d_img'], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(ot.grad.numpy(), g[f'{tag}_d_txt'], rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize('dset', ['Cifar100', 'AG_NEWS'])
@pytest.mark.parametrize('tag', ['both', 'intra', 'inter', 'both_scaled'])
def test_unimodal_contrast(golden, dset, tag):
    g, s = golden('uni_contrast'), golden('mm_contrast')
    # uni_contrast used seed 11: regenerate the same tensors through the shared recipe
    gen = torch.Generator().manual_seed(11)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    n_pub, d, b = 96, 32, 8
    g_img = unit(torch.randn(n_pub, d, generator=gen, dtype=torch.float64))
    g_txt = unit(0.7 * g_img + 0.5 * torch.randn(n_pub, d, generator=gen, dtype=torch.float64))
    cur_img = unit(g_img + 0.4 * torch.randn(n_pub, d, generator=gen, dtype=torch.float64))
    cur_txt = unit(g_txt + 0.4 * torch.randn(n_pub, d, generator=gen, dtype=torch.float64))
    old_img = unit(cur_img + 0.2 * torch.randn(n_pub, d, generator=gen, dtype=torch.float64))
    old_txt = unit(cur_txt + 0.2 * torch.randn(n_pub, d, generator=gen, dtype=torch.float64))
    _ = torch.randperm(10 * n_pub, generator=gen)
    rows = torch.randperm(n_pub, generator=gen)[:b].tolist()
    is_img = dset == 'Cifar100'
    feat = (cur_img if is_img else cur_txt)[rows].clone().requires_grad_(True)
    old = (old_img if is_img else old_txt)[rows]
    parts = O.unimodal_contrast_loss(feat, old, g_img if is_img else g_txt, g_txt if is_img else g_img, rows,
                                     0.5, tag == 'both_scaled')
    loss = {'intra': parts['intra'], 'inter': parts['inter']}.get(tag, parts['loss'])
    loss.backward()
    assert loss.item() == pytest.approx(float(g[f'{dset}_{tag}_loss']), rel=1e-12)
    np.testing.assert_allclose(feat.grad.numpy(), g[f'{dset}_{tag}_d_feat'], rtol=1e-9, atol=1e-14)


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_recall_matches_reference_evaluator(golden, tag):
    g = golden('recall')
    img, cap = T(g[f'{tag}_img']), T(g[f'{tag}_cap'])
    il, cl = g[f'{tag}_img_lab'], g[f'{tag}_cap_lab']
    for name, (q, gal, ql, gl) in {'i2t': (img, cap, il, cl), 't2i': (cap, img, cl, il)}.items():
        sorted_ranks = O.recall_ranks_sorted(q, gal, ql, gl)
        count_ranks = O.recall_ranks_count(q, gal, ql, gl)
        np.testing.assert_array_equal(sorted_ranks, count_ranks)
        sc = O.recall_scores(count_ranks)
        for k in ('recall_1', 'recall_5', 'recall_10', 'rsum', 'medr', 'meanr'):
            assert sc[k] == pytest.approx(float(g[f'{tag}_{name}_{k}']), rel=1e-12), (name, k)


@pytest.mark.parametrize('tag', ['synth', 'cifar', 'agnews'])
def test_hetero_partition_bit_exact(golden, tag):
    g = golden('partition')
    n, k, nets, seed = (int(v) for v in g[f'{tag}_args'])
    name = {'synth': 'synth', 'cifar': 'cifar100', 'agnews': 'AG_NEWS'}[tag]
    part = O.hetero_partition(name, n, nets, float(g[f'{tag}_alpha']), np.arange(n) % k, seed=seed)
    assert [len(part[j]) for j in range(nets)] == list(g[f'{tag}_sizes'])
    np.testing.assert_array_equal(np.asarray([part[j][:8] for j in range(nets)]), g[f'{tag}_head'])
    assert O.partition_digest(part) == str(g[f'{tag}_sha256'])
    flat = np.sort(np.concatenate([np.asarray(part[j]) for j in range(nets)]))
    np.testing.assert_array_equal(flat, np.arange(n))  # disjoint and complete


def test_survey_partition_kat():
    """SURVEY.md section 4 item 5 known answer."""
    part = O.hetero_partition('synth', 50000, 4, 0.1, np.arange(50000) % 100, seed=2021)
    assert [len(part[j]) for j in range(4)] == [12562, 12621, 12162, 12655]
    assert part[0][:5] == [15037, 46424, 20641, 12533, 4243]
    assert O.partition_digest(part).startswith('c5b16d61cd6565da')


def test_flickr_shard_partition_bit_exact(golden):
    g = golden('partition')
    part = O.shard_partition(145000, 15, 150, seed=2021)
    assert [len(part[j]) for j in range(15)] == list(g['f30k_sizes'])
    np.testing.assert_array_equal(np.asarray([part[j][:8] for j in range(15)]), g['f30k_head'])
    assert O.partition_digest(part) == str(g['f30k_sha256'])
    assert list(g['f30k_sizes']) == [9660] * 14 + [9760]


def test_conw_closed_form_small():
    """s_c[n] = <V_c[n],G[n]> - logsumexp_j <V_c[n],G[j]>."""
    g_img, g_txt, i_vecs, _ = conw_inputs(3, 512, 64, 3)
    vecs = [T(v).double() for v in i_vecs]
    G = T(g_txt).double()
    agg, w = O.conw_aggregate(vecs, G, chunk=100)
    s = torch.stack([(v * G).sum(1) - torch.logsumexp(v @ G.T, dim=1) for v in vecs])
    np.testing.assert_allclose(w.numpy(), torch.softmax(s, 0).numpy(), rtol=1e-9)
    np.testing.assert_allclose(agg.numpy(), sum(v * torch.softmax(s, 0)[c][:, None] for c, v in enumerate(vecs)).numpy(),
                               rtol=1e-9)


def test_conw_full_size_against_reference(golden):
    """The reference's aggregation() only runs at N = 50000 (hard-coded); compare the oracle on the same inputs."""
    g = golden('conw')
    n, d, c = int(g['n']), int(g['d']), int(g['c'])
    g_img, g_txt, i_vecs, t_vecs = conw_inputs(int(g['seed']), n, d, c)
    agg_i, _ = O.conw_aggregate([T(v) for v in i_vecs], T(g_txt))
    agg_t, _ = O.conw_aggregate([T(v) for v in t_vecs], T(g_img))
    rows = g['rows']
    np.testing.assert_allclose(agg_i.numpy()[rows], g['img_rows'], rtol=2e-5, atol=2e-7)
    np.testing.assert_allclose(agg_t.numpy()[rows], g['txt_rows'], rtol=2e-5, atol=2e-7)
    assert agg_i.double().abs().sum().item() == pytest.approx(float(g['img_abs_sum']), rel=1e-6)
    assert agg_t.double().abs().sum().item() == pytest.approx(float(g['txt_abs_sum']), rel=1e-6)


def test_tower_restatement_matches_reference_modules(golden):
    """oracle/torch_towers.py against the reference's own PCME / EncoderImage / PIENet / resnet18_client (executed by
    tests/golden/make_golden.py::case_towers on the same deterministic weights): fp32, rel 1e-5."""
    from transformers import BertConfig
    from oracle import torch_towers as RT
    g = golden('towers')
    model = RT.RefPCME('resnet18', 64, BertConfig(num_hidden_layers=2))
    RT.fill_deterministic(model, seed=31)
    images, ids = T(g['images']), T(g['ids'])
    mask = torch.ones_like(ids)
    mask[1, int(g['len1']):] = 0
    model.eval()
    with torch.no_grad():
        o = model(images, ids * mask, mask, torch.zeros_like(ids))
    np.testing.assert_allclose(o['image_features'].numpy(), g['eval_image_features'], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(o['caption_features'].numpy(), g['eval_caption_features'], rtol=1e-4, atol=1e-6)
    model.img_enc.train()
    with torch.no_grad():
        tr = model.img_enc(images)['embedding']
    np.testing.assert_allclose(tr.numpy(), g['train_image_embedding'], rtol=1e-4, atol=1e-6)
    assert list(g['none_keys']) == sorted(['image_attentions', 'image_residuals', 'image_logsigma', 'image_logsigma_att',
                                           'caption_attentions', 'caption_residuals', 'caption_logsigma',
                                           'caption_logsigma_att'])
    client = RT.RefImageClient(num_class=10, embed_dim=64)
    RT.fill_deterministic(client, seed=33)
    with torch.no_grad():
        client.linear.weight.mul_(0.05)
    small = T(g['client_images'])
    client.train()
    x1, _, w1, _ = client(small)
    np.testing.assert_allclose(x1.detach().numpy(), g['client_x1'], rtol=1e-4, atol=1e-5)
    assert float(w1.min()) == float(g['client_w_min']) == 0.0
    client.phase, client.is_train = 'extract_conv_feature', False
    with torch.no_grad():
        emb = client(small)
    np.testing.assert_allclose(emb.numpy(), g['client_embedding'], rtol=1e-4, atol=1e-6)


def test_text_tower_restatement_matches_reference_modules(golden):
    """oracle/torch_towers.py RefGRUEncoderText / RefTextClient against the reference's own
    caption_encoder.EncoderText and language_model.EncoderText (tests/golden/make_golden.py::case_text_towers, same
    deterministic weights, ragged length-sorted captions): forward values and parameter gradients, fp32, rel 1e-5."""
    from oracle import torch_towers as RT
    g = golden('text_towers')
    x, lengths, coef = T(g['x']), T(g['lengths']), T(g['coef'])
    enc = RT.RefGRUEncoderText(500, 300, 64)
    RT.fill_deterministic(enc, seed=41)
    enc.train()
    emb = enc(x, lengths)['embedding']
    np.testing.assert_allclose(emb.detach().numpy(), g['mm_embedding'], rtol=1e-5, atol=1e-6)
    (emb * coef).sum().backward()
    for name, p in enc.named_parameters():
        ref = g['mm_grad.' + name]
        got = p.grad.numpy() if p.grad is not None else np.zeros_like(ref)
        np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-7, err_msg=name)
    # the reverse direction contributes one step only: its recurrent matrix receives no gradient
    assert np.abs(g['mm_grad.rnn.weight_hh_l0_reverse']).max() == 0.0
    client = RT.RefTextClient(int(g['uni_vocab']), 300, 64, num_class=4, scale=128)
    RT.fill_deterministic(client, seed=43)
    client.train()
    labels = T(g['uni_labels'])
    loss, _ = RT.ref_text_supervised_loss(client, x, lengths, labels, 4)
    x1, x2, w1, w2 = client(x, lengths)          # second forward: weights already clamped, same values
    np.testing.assert_allclose(x1.detach().numpy(), g['uni_x1'], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(x2.detach().numpy(), g['uni_x2'], rtol=1e-4, atol=1e-4)
    assert abs(float(loss) - float(g['uni_loss'])) <= 1e-5 * abs(float(g['uni_loss']))
    assert min(float(w1.min()), float(w2.min())) == float(g['uni_w_min']) == 0.0
    loss.backward()
    params = dict(client.named_parameters())
    for name in ['rnn.weight_hh_l0', 'rnn.weight_ih_l0_reverse', 'pie_net.fc.weight', 'class_fc.weight', 'class_fc.bias']:
        np.testing.assert_allclose(params[name].grad.numpy(), g['uni_grad.' + name], rtol=2e-4, atol=1e-6, err_msg=name)
    np.testing.assert_allclose(params['embed.weight'].grad[:500].numpy(), g['uni_grad_embed_rows'], rtol=2e-4, atol=1e-7)
    client.is_train = False
    with torch.no_grad():
        np.testing.assert_allclose(client(x, lengths).numpy(), g['uni_embedding'], rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------- AdamP pinning
# The adamp==0.3.0 package is absent (reference requirements.txt:1), so `adamp_step` cannot be pinned against it as a
# whole.  What can be pinned: (1) without a projection AdamP IS Adam - checked against torch.optim.Adam in fp64;
# (2) the projection branch against a known-answer case worked out by hand from Algorithm 1 of the paper
# (Heo et al., ICLR 2021): cosine test `max_rows |<g, w>| / (|g| |w|) < delta / sqrt(dim)`, projected update
# `u - w_hat <w_hat, u>`, weight-decay ratio wd_ratio.
def test_adamp_without_projection_is_torch_adam():
    from oracle import creamfl_oracle as O
    g = torch.Generator().manual_seed(3)
    shapes = [(7,), (5, 3), (4, 2, 3, 3)]
    ps = [torch.randn(s, generator=g, dtype=torch.float64) for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    adam = torch.optim.Adam(ref, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, foreach=False)
    ms, vs = [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    for step in range(1, 6):
        # gradients with a large component along the weights: the cosine test never fires
        grads = [0.5 * p + 0.3 * torch.randn(p.shape, generator=g, dtype=torch.float64) for p in ps]
        for r, gr in zip(ref, grads):
            r.grad = gr.clone()
        adam.step()
        fired = O.adamp_step(ps, grads, ms, vs, step, 2e-4)
        assert fired == [0, 0, 0]
        for p, r in zip(ps, ref):
            # same arithmetic up to the association of (m / denom) * step_size: a few fp64 ulps
            assert float((p - r.detach()).abs().max()) <= 4 * torch.finfo(torch.float64).eps * float(p.abs().max())


def test_adamp_projection_known_answer():
    from oracle import creamfl_oracle as O
    lr, wd, eps, delta, wd_ratio = 0.01, 0.5, 1e-8, 0.1, 0.1
    w = torch.tensor([[1.0, 0.0], [0.0, 2.0]], dtype=torch.float64)
    g = torch.tensor([[0.05, 1.0], [-3.0, 0.1]], dtype=torch.float64)
    # row cosines: 0.05 / sqrt(1.0025) = 0.04994 and 0.2 / (2 * sqrt(9.01)) = 0.03331, both < delta / sqrt(2) = 0.0707
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    p = w.clone()
    fired = O.adamp_step([p], [g], [m], [v], 1, lr, (0.9, 0.999), eps, wd, delta, wd_ratio)
    assert fired == [1]                                    # the channel-wise test fires first
    # step 1 by hand: m = 0.1 g, v = 0.001 g^2, denom = |g| + eps, lr / bias_correction1 = lr / 0.1, so the Adam
    # direction is u = g / (|g| + eps) element-wise (scaled by 0.1, undone by the bias correction)
    exp = torch.empty_like(w)
    for r in range(2):
        u = [g[r, c].item() / (abs(g[r, c].item()) + eps) for c in range(2)]
        norm = (w[r, 0].item() ** 2 + w[r, 1].item() ** 2) ** 0.5 + eps
        wh = [w[r, c].item() / norm for c in range(2)]
        dot = wh[0] * u[0] + wh[1] * u[1]
        for c in range(2):
            exp[r, c] = w[r, c].item() * (1 - lr * wd * wd_ratio) - lr * (u[c] - wh[c] * dot)
    assert float((p - exp).abs().max()) < 1e-15
    # the radial component of the update is gone (that is the point of AdamP): <w_hat, p_new - w (1 - lr wd ratio)> = 0
    upd = p - w * (1 - lr * wd * wd_ratio)
    assert float((upd * w).sum(1).abs().max()) < 1e-9
    # row 0 one notch above the channel threshold (cosine 0.0797 > 0.0707): the channel test fails, the layer-wise test
    # |<g, w>| / (|g| |w|) = 0.28 / (3.1649 * 2.2361) = 0.0396 < delta / sqrt(4) = 0.05 fires instead
    g2 = torch.tensor([[0.08, 1.0], [-3.0, 0.1]], dtype=torch.float64)
    p2, m2, v2 = w.clone(), torch.zeros_like(w), torch.zeros_like(w)
    assert O.adamp_step([p2], [g2], [m2], [v2], 1, lr, (0.9, 0.999), eps, wd, delta, wd_ratio) == [2]
    upd2 = p2 - w * (1 - lr * wd * wd_ratio)
    assert abs(float((upd2 * w).sum())) < 1e-9                    # radial component w.r.t. the whole tensor removed
    # well above both thresholds nothing fires: plain Adam step with the full weight decay
    g3 = torch.tensor([[1.0, 1.0], [-3.0, 2.0]], dtype=torch.float64)
    p3, m3, v3 = w.clone(), torch.zeros_like(w), torch.zeros_like(w)
    assert O.adamp_step([p3], [g3], [m3], [v3], 1, lr, (0.9, 0.999), eps, wd, delta, wd_ratio) == [0]
    exp3 = w * (1 - lr * wd) - lr * g3 / (g3.abs() + eps)
    assert float((p3 - exp3).abs().max()) < 1e-15


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_match_prob(golden, tag):
    """criterion.match_prob of the reference (probemb.py:210-219), query-broadcast / same-N / gallery-of-one / 2-D."""
    g = golden('match_prob')
    shift = torch.tensor([float(g[f'{tag}_shift'])], dtype=torch.float64)
    scale = torch.tensor([float(g[f'{tag}_scale'])], dtype=torch.float64)
    prob = O.pcme_match_prob(T(g[f'{tag}_q']), T(g[f'{tag}_g']), shift, scale)
    np.testing.assert_allclose(prob.numpy(), g[f'{tag}_prob'], rtol=1e-12, atol=0)


def test_match_prob_rejects_non_broadcastable():
    with pytest.raises(RuntimeError):
        O.pcme_match_prob(torch.zeros(3, 2, 4), torch.zeros(2, 2, 4), torch.ones(1), torch.ones(1))
