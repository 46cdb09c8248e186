"""The product's criterion.match_prob / the evaluator's 'matching_prob' mode (host tensor ops, evaluation only)
against what the reference's MCSoftContrastiveLoss.match_prob returned (tests/golden/match_prob.npz)."""
import numpy as np
import pytest
import torch

from creamfl_b200.criterions import MCSoftContrastiveLoss

T = lambda a: torch.from_numpy(np.asarray(a))


def _criterion(shift, scale):
    crit = MCSoftContrastiveLoss({'init_shift': shift, 'init_negative_scale': scale, 'num_samples': 1})
    return crit.to('cpu').double()


@pytest.mark.parametrize('tag', ['a', 'b', 'c', 'd'])
def test_match_prob_matches_reference(golden, tag):
    g = golden('match_prob')
    crit = _criterion(float(g[f'{tag}_shift']), float(g[f'{tag}_scale']))
    with torch.no_grad():
        prob = crit.match_prob(T(g[f'{tag}_q']), T(g[f'{tag}_g']), None, None)
    # sigmoid(2 l) against the reference's e^l / (e^l + e^-l): same value, a few ulp apart
    np.testing.assert_allclose(prob.numpy(), g[f'{tag}_prob'], rtol=1e-10, atol=1e-300)


def test_match_prob_rejects_non_broadcastable():
    crit = _criterion(1.0, 1.0)
    with pytest.raises(RuntimeError):
        crit.match_prob(torch.zeros(3, 2, 4), torch.zeros(2, 2, 4))


def test_match_prob_large_logits_stay_finite():
    """|l| > 88 overflows the reference's fp32 quotient to nan; sigmoid(2 l) saturates instead."""
    crit = _criterion(200.0, 1.0).float()
    with torch.no_grad():
        p = crit.match_prob(torch.zeros(2, 1, 4), torch.zeros(2, 1, 4))
    assert torch.isfinite(p).all() and float(p.min()) == 1.0


def test_evaluator_matching_prob_mode_ranks_like_matmul_on_unit_features(golden):
    """For one unit-norm embedding per item the matching probability is a decreasing function of the distance, i.e.
    an increasing function of the dot product: the 'matching_prob' evaluator must reproduce the scores the reference's
    evaluator produced with 'matmul' on the same features (tests/golden/recall.npz), as long as no two probabilities
    collide - fp64 and a mild scale keep them apart."""
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    for p in (str(root), str(root / 'src')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from src.algorithms.eval_coco import COCOEvaluator
    g = golden('recall')
    img, cap = T(g['a_img']).double(), T(g['a_cap']).double()
    assert torch.allclose(img.norm(dim=1), torch.ones(len(img), dtype=torch.float64), atol=1e-5)
    il, cl = T(g['a_img_lab']), T(g['a_cap_lab'])
    ev = COCOEvaluator(eval_method='matching_prob')
    ev.set_criterion(_criterion(1.0, 2.0))
    for name, sc in (('i2t', ev.evaluate_recall(img, cap, il, cl)), ('t2i', ev.evaluate_recall(cap, img, cl, il))):
        for k in ('recall_1', 'recall_5', 'recall_10', 'medr', 'meanr'):
            assert sc[k] == pytest.approx(float(g[f'a_{name}_{k}']), rel=1e-9), (name, k)
    with pytest.raises(ValueError):
        COCOEvaluator(eval_method='cosine')


def test_losses_factory_softmax(monkeypatch):
    """src/losses.create('softmax') (reference src/losses/__init__.py:11-38, ClientTrainer.py:280): mean cross entropy
    like nn.CrossEntropyLoss(); the kernel wrapper is swapped for the test-only emulation on the CPU."""
    import sys
    from pathlib import Path
    import kernel_emulation as KE
    root = Path(__file__).resolve().parent.parent
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    from creamfl_b200 import ops
    from src import losses
    monkeypatch.setattr(ops, 'cross_entropy', KE.cross_entropy)
    crit = losses.create('softmax')
    g = torch.Generator().manual_seed(3)
    x = torch.randn(9, 5, generator=g, requires_grad=True)
    y = torch.randint(0, 5, (9,), generator=g)
    loss = crit(x, y)
    assert loss.item() == pytest.approx(torch.nn.CrossEntropyLoss()(x, y).item(), rel=1e-6)
    assert losses.names() == ['softmax']
    with pytest.raises(KeyError):
        losses.create('triplet')
    with pytest.raises(KeyError):
        losses.create('nope')
