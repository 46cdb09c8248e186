"""GPU parity tests (run on the B200 box with `-m gpu`): every loss / aggregation entry point of the C ABI,
reached through creamfl_b200.ops (ctypes -> libcreamfl_b200.so), against the CPU oracle on the same seeded inputs
and against the golden vectors the reference's own code produced (tests/golden/*.npz).

Tolerances (stated per test): integer outputs exact; fp32 CUDA-core kernels rel 1e-4 on losses / rel-L2 1e-4 on
gradients; tensor-core kernels are fed bf16 operands, the oracle is fed the same bf16-rounded values in fp64, so
what remains is fp32 accumulation order and ex2.approx (rel 2e-4 on losses, rel-L2 2e-3 on gradients whose
probabilities are bf16-rounded before the dQ contraction).
"""
import numpy as np
import pytest
import torch

from conftest import conw_inputs

pytestmark = pytest.mark.gpu

T = lambda a: torch.from_numpy(np.asarray(a))


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from creamfl_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def O():
    from oracle import creamfl_oracle
    return creamfl_oracle


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float64)


# --------------------------------------------------------------------------------------------------- PCME
@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_pcme_golden(ops, golden, tag):
    """fp32 kernel vs the reference's MCSoftContrastiveLoss output (fp64 golden): loss rel 1e-5, grads rel-L2 1e-4."""
    g = golden('pcme')
    img = T(g[f'{tag}_img']).float().cuda().requires_grad_(True)
    txt = T(g[f'{tag}_txt']).float().cuda().requires_grad_(True)
    shift = torch.tensor(float(g[f'{tag}_shift']), device='cuda', requires_grad=True)
    scale = torch.tensor(float(g[f'{tag}_scale']), device='cuda', requires_grad=True)
    loss, parts = ops.pcme_loss(img, txt, shift, scale)
    loss.backward()
    assert loss.item() == pytest.approx(float(g[f'{tag}_loss']), rel=1e-5)
    assert parts[1].item() == pytest.approx(float(g[f'{tag}_i2t_pos']), rel=1e-4, abs=1e-6)
    assert parts[2].item() == pytest.approx(float(g[f'{tag}_i2t_neg']), rel=1e-4, abs=1e-6)
    assert rel_l2(img.grad, g[f'{tag}_d_img']) < 1e-4
    assert rel_l2(txt.grad, g[f'{tag}_d_txt']) < 1e-4
    assert shift.grad.item() == pytest.approx(float(g[f'{tag}_d_shift'][0]), rel=1e-4)
    assert scale.grad.item() == pytest.approx(float(g[f'{tag}_d_scale'][0]), rel=1e-4)


@pytest.mark.parametrize('n,d', [(128, 256), (16, 256), (1, 64), (77, 40), (512, 256)])
def test_pcme_vs_oracle(ops, O, n, d):
    gen = torch.Generator().manual_seed(100 + n)
    img = unit(torch.randn(n, d, generator=gen, dtype=torch.float64))
    txt = unit(0.6 * img + 0.6 * torch.randn(n, d, generator=gen, dtype=torch.float64) / d ** 0.5)
    io, to = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
    so = torch.tensor(15.0, dtype=torch.float64, requires_grad=True)
    no = torch.tensor(15.0, dtype=torch.float64, requires_grad=True)
    lo, _ = O.pcme_loss(io, to, so, no)
    lo.backward()
    ig, tg = img.float().cuda().requires_grad_(True), txt.float().cuda().requires_grad_(True)
    sg = torch.tensor(15.0, device='cuda', requires_grad=True)
    ng = torch.tensor(15.0, device='cuda', requires_grad=True)
    lg, _ = ops.pcme_loss(ig, tg, sg, ng)
    (lg * 0.5).backward()
    assert lg.item() == pytest.approx(lo.item(), rel=2e-5)
    assert rel_l2(ig.grad * 2, io.grad) < 2e-4
    assert rel_l2(tg.grad * 2, to.grad) < 2e-4
    assert sg.grad.item() * 2 == pytest.approx(so.grad.item(), rel=2e-4)
    assert ng.grad.item() * 2 == pytest.approx(no.grad.item(), rel=2e-4)


def test_pcme_rejects_ragged(ops):
    with pytest.raises(RuntimeError):
        ops.pcme_loss(torch.zeros(3, 8, device='cuda'), torch.zeros(2, 8, device='cuda'),
                      torch.tensor(1.0, device='cuda'), torch.tensor(1.0, device='cuda'))


def test_ops_reject_cpu_tensors(ops):
    with pytest.raises(RuntimeError):
        ops.l2_normalize(torch.zeros(4, 8))


# --------------------------------------------------------------------------------------------------- InfoNCE
@pytest.mark.parametrize('b,n,d', [(128, 50000, 256), (8, 96, 64), (1, 64, 64), (80, 1000, 128), (130, 4100, 192),
                                   (512, 50000, 256), (256, 777, 256)])
def test_infonce_vs_oracle(ops, O, b, n, d):
    """bf16 operands on both sides; loss rel 2e-4, dQ rel-L2 2e-3 (P is rounded to bf16 before P.G)."""
    gen = torch.Generator().manual_seed(7 * b + n)
    bank = unit(torch.randn(n, d, generator=gen))
    labels = torch.randint(0, n, (b,), generator=gen)
    q = unit(bank[labels] + 0.5 * torch.randn(b, d, generator=gen) / d ** 0.5)
    qo = bf16_round(q).requires_grad_(True)
    lo = O.inter_infonce(qo, bf16_round(bank), labels)
    lo.backward()
    qg = q.cuda().requires_grad_(True)
    lg = ops.infonce_loss(qg, ops.to_bf16(bank.cuda()), labels.cuda(), 2.0)
    (lg * 3.0).backward()
    assert lg.item() == pytest.approx(lo.item(), rel=2e-4)
    assert rel_l2(qg.grad / 3.0, qo.grad) < 2e-3


def test_mm_contrast_golden(ops, golden):
    """The reference's MMClientTrainer.train_epoch (both flags) on the golden batch: loss rel 2e-3 / grads rel-L2
    1e-2 - the golden is fp64 on unrounded inputs, the CUDA path rounds Q and the bank to bf16 (D = 32 is padded
    to 64 with zero columns, which changes nothing)."""
    g = golden('mm_contrast')
    rows = torch.from_numpy(g['rows']).long()
    pad = lambda x: torch.nn.functional.pad(T(x).float(), (0, 32))
    g_img, g_txt = pad(g['g_img']).cuda(), pad(g['g_txt']).cuda()
    oi = pad(g['cur_img'])[rows].cuda().requires_grad_(True)
    ot = pad(g['cur_txt'])[rows].cuda().requires_grad_(True)
    old_i, old_t = pad(g['old_img'])[rows].cuda(), pad(g['old_txt'])[rows].cuda()
    idx = rows.cuda()
    b = len(rows)
    intra = ops.moon_intra_loss(oi, old_i, g_img, idx, 2.0, 2 * b) + ops.moon_intra_loss(ot, old_t, g_txt, idx, 2.0,
                                                                                          2 * b)
    inter = ops.infonce_loss(oi, ops.to_bf16(g_txt), idx, 2.0) + ops.infonce_loss(ot, ops.to_bf16(g_img), idx, 2.0)
    loss = (intra + inter) * 0.5
    loss.backward()
    assert loss.item() == pytest.approx(float(g['both_loss']), rel=2e-3)
    assert rel_l2(oi.grad[:, :32], g['both_d_img']) < 1e-2
    assert rel_l2(ot.grad[:, :32], g['both_d_txt']) < 1e-2


# --------------------------------------------------------------------------------------------------- MOON / MSE / L2
@pytest.mark.parametrize('r,d', [(128, 256), (256, 256), (1, 32), (33, 100)])
def test_moon_vs_oracle(ops, O, r, d):
    gen = torch.Generator().manual_seed(r)
    bank = unit(torch.randn(500, d, generator=gen, dtype=torch.float64))
    idx = torch.randint(0, 500, (r,), generator=gen)
    z = unit(torch.randn(r, d, generator=gen, dtype=torch.float64))
    zold = unit(z + 0.3 * torch.randn(r, d, generator=gen, dtype=torch.float64))
    zo = z.clone().requires_grad_(True)
    lo = O.moon_intra(zo, zold, bank[idx], denom=2 * r)
    lo.backward()
    zg = z.float().cuda().requires_grad_(True)
    lg = ops.moon_intra_loss(zg, zold.float().cuda(), bank.float().cuda(), idx.cuda(), 2.0, 2 * r)
    lg.backward()
    assert lg.item() == pytest.approx(lo.item(), rel=1e-5)
    assert rel_l2(zg.grad, zo.grad) < 1e-5


@pytest.mark.parametrize('r,d', [(128, 256), (3, 8), (257, 256)])
def test_mse_gather_vs_oracle(ops, O, r, d):
    gen = torch.Generator().manual_seed(r + 1)
    agg = torch.randn(1000, d, generator=gen, dtype=torch.float64)
    idx = torch.randint(0, 1000, (r,), generator=gen)
    x = torch.randn(r, d, generator=gen, dtype=torch.float64)
    xo = x.clone().requires_grad_(True)
    lo = O.distill_mse(xo, agg, idx)
    lo.backward()
    xg = x.float().cuda().requires_grad_(True)
    lg = ops.mse_gather_loss(xg, agg.float().cuda(), idx.cuda())
    lg.backward()
    assert lg.item() == pytest.approx(lo.item(), rel=1e-5)
    assert rel_l2(xg.grad, xo.grad) < 1e-5


@pytest.mark.parametrize('r,d', [(128, 256), (5, 7), (1000, 512)])
def test_l2norm_vs_oracle(ops, O, r, d):
    gen = torch.Generator().manual_seed(r + 2)
    x = torch.randn(r, d, generator=gen, dtype=torch.float64) * 3
    w = torch.randn(r, d, generator=gen, dtype=torch.float64)
    xo = x.clone().requires_grad_(True)
    (O.l2_normalize(xo) * w).sum().backward()
    xg = x.float().cuda().requires_grad_(True)
    y = ops.l2_normalize(xg)
    (y * w.float().cuda()).sum().backward()
    assert rel_l2(y, O.l2_normalize(x)) < 1e-6
    assert rel_l2(xg.grad, xo.grad) < 1e-5


# --------------------------------------------------------------------------------------------------- con_w
@pytest.mark.parametrize('n,d,c', [(512, 64, 3), (4096, 256, 4), (1000, 128, 2), (64, 64, 1), (333, 192, 5)])
def test_conw_vs_oracle(ops, O, n, d, c):
    """scores abs 2e-4 (bf16-rounded operands on both sides), weights abs 1e-4, aggregate rel-L2 1e-4."""
    g_img, g_txt, i_vecs, _ = conw_inputs(n + d, n, d, c)
    G = T(g_txt)
    vo = [bf16_round(T(v)) for v in i_vecs]
    so = torch.stack([O.conw_scores(v, bf16_round(G)) for v in vo])
    wo = torch.softmax(so, 0)
    agg_o = sum(T(v).double() * wo[k][:, None] for k, v in enumerate(i_vecs))
    vg = [T(v).cuda() for v in i_vecs]
    gb = ops.to_bf16(G.cuda())
    sg = torch.stack([ops.conw_score(ops.to_bf16(v), gb) for v in vg])
    agg_g, wg = ops.conw_reduce(vg, sg, want_weights=True)
    assert (sg.cpu().double() - so).abs().max().item() < 2e-4
    assert (wg.cpu().double() - wo).abs().max().item() < 1e-4
    assert rel_l2(agg_g, agg_o) < 1e-4


def test_conw_full_size_d256_vs_oracle(ops, O):
    """N = 50000 (the size the reference hard-codes, MMFL.py:302) at the embedding width of the benchmark, D = 256,
    two clients, against the oracle on the same bf16-rounded operands (fp32 on the host: bf16 products are exact in
    fp32, the 256-term sums and the 50000-term log-sum-exp carry ~1e-6): scores abs 3e-4, weights abs 2e-4,
    aggregate rel-L2 2e-4."""
    n, d, c = 50000, 256, 2
    g_img, g_txt, i_vecs, _ = conw_inputs(77, n, d, c)
    G = bf16_round(T(g_txt)).float()
    vo = [bf16_round(T(v)).float() for v in i_vecs]
    so = torch.stack([O.conw_scores(v, G) for v in vo]).double()
    wo = torch.softmax(so, 0)
    agg_o = sum(T(v).double() * wo[k][:, None] for k, v in enumerate(i_vecs))
    vg = [T(v).cuda() for v in i_vecs]
    gb = ops.to_bf16(T(g_txt).cuda())
    sg = torch.stack([ops.conw_score(ops.to_bf16(v), gb) for v in vg])
    agg_g, wg = ops.conw_reduce(vg, sg, want_weights=True)
    assert (sg.cpu().double() - so).abs().max().item() < 3e-4
    assert (wg.cpu().double() - wo).abs().max().item() < 2e-4
    assert rel_l2(agg_g, agg_o) < 2e-4


def test_conw_full_size_golden(ops, golden):
    """N = 50000 (the size the reference hard-codes, MMFL.py:302), D = 64, three clients (the golden's own n, d, c;
    D = 256 is covered against the oracle above): rows of the reference's own aggregation() output.  bf16 operands vs the reference's fp32: weights shift by <= 2e-3, so the
    aggregated rows agree to abs 2e-4 (unit-norm rows, entries ~0.06)."""
    g = golden('conw')
    n, d, c = int(g['n']), int(g['d']), int(g['c'])
    g_img, g_txt, i_vecs, t_vecs = conw_inputs(int(g['seed']), n, d, c)
    agg_i = ops.conw_aggregate([T(v).cuda() for v in i_vecs], T(g_txt).cuda())
    agg_t = ops.conw_aggregate([T(v).cuda() for v in t_vecs], T(g_img).cuda())
    rows = torch.from_numpy(g['rows']).long()
    assert (agg_i[rows.cuda()].cpu().double() - T(g['img_rows']).double()).abs().max().item() < 2e-4
    assert (agg_t[rows.cuda()].cpu().double() - T(g['txt_rows']).double()).abs().max().item() < 2e-4
    assert agg_i.double().abs().sum().item() == pytest.approx(float(g['img_abs_sum']), rel=1e-4)
    assert agg_t.double().abs().sum().item() == pytest.approx(float(g['txt_abs_sum']), rel=1e-4)


def test_conw_single_client_is_identity(ops):
    """Size-independent property: with one client the softmax over clients is 1, the aggregate is the input."""
    g_img, g_txt, i_vecs, _ = conw_inputs(5, 50000, 256, 1)
    v = T(i_vecs[0]).cuda()
    out = ops.conw_aggregate([v], T(g_txt).cuda())
    assert torch.equal(out, v)


def test_conw_identical_clients_average(ops):
    """Property at full size: identical clients get identical scores -> weights exactly 1/C."""
    g_img, g_txt, i_vecs, _ = conw_inputs(6, 50000, 256, 1)
    v = T(i_vecs[0]).cuda()
    out, w = ops.conw_reduce([v, v, v, v], torch.stack([ops.conw_score(ops.to_bf16(v), ops.to_bf16(T(g_txt).cuda()))] * 4),
                             want_weights=True)
    assert torch.equal(w, torch.full_like(w, 0.25))
    assert torch.allclose(out, v, rtol=0, atol=1e-7)


# --------------------------------------------------------------------------------------------------- Recall@K
@pytest.mark.parametrize('tag', ['a', 'b'])
def test_recall_golden(ops, O, golden, tag):
    """fp32 similarities vs the reference's fp64 evaluator: scores identical on the golden sets (no near ties)."""
    g = golden('recall')
    img, cap = T(g[f'{tag}_img']).float().cuda(), T(g[f'{tag}_cap']).float().cuda()
    il, cl = T(g[f'{tag}_img_lab']).cuda(), T(g[f'{tag}_cap_lab']).cuda()
    for name, (q, gal, ql, gl) in {'i2t': (img, cap, il, cl), 't2i': (cap, img, cl, il)}.items():
        ranks = ops.recall_ranks(q, gal, ql, gl).cpu().numpy().astype(np.float64)
        sc = O.recall_scores(ranks)
        for k in ('recall_1', 'recall_5', 'recall_10', 'medr'):
            assert sc[k] == pytest.approx(float(g[f'{tag}_{name}_{k}']), abs=0.11), (name, k)
        assert sc['meanr'] == pytest.approx(float(g[f'{tag}_{name}_meanr']), rel=2e-3)


def test_recall_coco1k_shape_vs_oracle(ops, O):
    """COCO-1K fold shape (1000 images x 5000 captions): ranks equal to the fp64 oracle except where fp32
    rounding flips a near tie (<= 0.2 % of queries, each by a small amount); Recall@K within 0.2 pt."""
    gen = torch.Generator().manual_seed(42)
    img = unit(torch.randn(1000, 256, generator=gen))
    cap = unit(img.repeat_interleave(5, 0) + 0.9 * torch.randn(5000, 256, generator=gen) / 16)
    il, cl = torch.arange(1000), torch.arange(1000).repeat_interleave(5)
    for q, gal, ql, gl in ((img, cap, il, cl), (cap, img, cl, il)):
        want = O.recall_ranks_count(q, gal, ql.numpy(), gl.numpy())
        got = ops.recall_ranks(q.cuda(), gal.cuda(), ql.cuda(), gl.cuda()).cpu().numpy().astype(np.float64)
        assert (want != got).mean() <= 0.002
        a, b = O.recall_scores(want), O.recall_scores(got)
        for k in ('recall_1', 'recall_5', 'recall_10'):
            assert abs(a[k] - b[k]) <= 0.2


def test_recall_query_without_positive_ranks_last(ops):
    """A query whose label has no match in the gallery gets rank Ng (worse than every hit), not a silent rank 0 - the
    reference's evaluator raises there (eval_coco.py:311-317 indexes an empty np.where)."""
    g = torch.Generator().manual_seed(3)
    q, gal = unit(torch.randn(70, 64, generator=g)).cuda(), unit(torch.randn(150, 64, generator=g)).cuda()
    ql, gl = torch.arange(70), torch.arange(150) % 60           # labels 60..69 never occur in the gallery
    ranks = ops.recall_ranks(q, gal, ql.cuda(), gl.cuda()).cpu()
    assert bool((ranks[60:] == 150).all()) and bool((ranks[:60] < 150).all())

