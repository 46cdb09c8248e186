"""GPU parity tests of the fused optimizer step (creamfl_optimizer_step) against the oracle restatement of AdamP /
clip_grad_norm_ / SGD-momentum in fp64.  Tolerance: parameters after 3 steps rel-L2 <= 2e-6 (fp32 arithmetic),
the projection decision per tensor identical."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from creamfl_b200 import optim
    from oracle import creamfl_oracle as O
    return optim, O


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def make_params(gen):
    """A mix of tensors: scale-invariant style (grad orthogonal to weight -> channel projection), generic 2-D / 4-D,
    1-D, scalars."""
    shapes = [(64, 3, 7, 7), (128, 64, 3, 3), (256, 768), (30, 300), (768,), (1,), (1,), (5, 9000)]
    ps = [torch.randn(s, generator=gen) * 0.1 for s in shapes]
    return ps


def make_grads(ps, gen, k):
    gs = []
    for i, p in enumerate(ps):
        g = torch.randn(p.shape, generator=gen) * (0.5 + k)
        if p.dim() > 1 and i in (1, 2):       # remove the per-row radial component: cos = 0 -> projection must fire
            pv, gv = p.reshape(p.shape[0], -1), g.reshape(p.shape[0], -1)
            gv -= pv * ((gv * pv).sum(1, keepdim=True) / (pv * pv).sum(1, keepdim=True))
        if i == 3:                            # remove only the whole-tensor radial component: layer view fires
            g -= p * ((g * p).sum() / (p * p).sum())
            g.reshape(30, -1)[:] += 0.05 * p.reshape(30, -1) * torch.sign(torch.randn(30, 1, generator=gen))
            g -= p * ((g * p).sum() / (p * p).sum())
        gs.append(g)
    return gs


@pytest.mark.parametrize('mode,wd,max_norm', [('adamp', 0.0, 2.0), ('adamp', 0.01, 0.0), ('adam', 0.0, 2.0),
                                              ('sgd', 5e-5, 0.0)])
def test_fused_step_vs_oracle(env, mode, wd, max_norm):
    optim, O = env
    gen = torch.Generator().manual_seed(3)
    ps = make_params(gen)
    dev_ps = [torch.nn.Parameter(p.clone().cuda()) for p in ps]
    lr = 2e-4 if mode != 'sgd' else 1e-2
    opt = optim.FusedOptimizer(dev_ps, lr=lr, weight_decay=wd, max_norm=max_norm, mode=mode, no_clip=dev_ps[5:7])
    ref_p = [p.double().clone() for p in ps]
    m = [torch.zeros_like(p) for p in ref_p]
    v = [torch.zeros_like(p) for p in ref_p]
    for k in range(3):
        # projection geometry must hold for the CURRENT parameters: rebuild the grads from the oracle's parameters
        gs = make_grads([p.float() for p in ref_p], gen, k)
        opt.zero_grad()
        for p, g in zip(dev_ps, gs):
            p.grad.copy_(g.cuda()) if p.grad is not None else setattr(p, 'grad', g.cuda())
        opt.step()
        g64 = [g.double().clone() for g in gs]
        if max_norm > 0:
            clip_idx = [i for i in range(len(ps)) if i not in (5, 6)]
            norm = O.clip_grad_norm([g64[i] for i in clip_idx], max_norm)
            assert opt.grad_norm.item() == pytest.approx(norm, rel=1e-5)
        if mode == 'sgd':
            O.sgd_momentum_step(ref_p, g64, m, k + 1, lr, 0.9, wd)
        else:
            fired = O.adamp_step(ref_p, g64, m, v, k + 1, lr, weight_decay=wd,
                                 delta=0.1 if mode == 'adamp' else -1.0)
            if mode == 'adamp':
                assert opt._flag.cpu().tolist() == fired
                assert fired[1] == 1 and fired[2] == 1 and fired[4] == 0
        for i, (p, r) in enumerate(zip(dev_ps, ref_p)):
            assert rel_l2(p.data, r) < 2e-6, (k, i)
    torch.cuda.synchronize()
    assert opt._total.item() == 0.0


def test_optimizer_on_param_store_refreshes_shadow(env):
    optim, O = env
    from creamfl_b200 import towers
    model = towers.ImageModel({'embed_dim': 256, 'cnn_type': 'resnet18'}).cuda().train()
    st = model.store()
    opt = optim.FusedOptimizer(model.parameters(), lr=1e-3, max_norm=2.0).attach_stores(model)
    images = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(1)).cuda()
    opt.zero_grad()
    before = st.flat.clone()
    model(images).sum().backward()
    opt.step()
    torch.cuda.synchronize()
    assert not torch.equal(before, st.flat)
    # every shadow equals the bf16 rounding of its fp32 master, including the padded stem filter
    for p in model.parameters():
        if p.dim() == 4:
            o, i, r, s = p.shape
            want = p.data.permute(0, 2, 3, 1).reshape(o, -1).to(torch.bfloat16)
            assert torch.equal(p._w16[:, :want.shape[1]], want)
        else:
            assert torch.equal(p._w16, p.data.to(torch.bfloat16))
    w = model.img_enc.cnn.conv1.weight
    assert w._w16.shape == (64, 152) and torch.count_nonzero(w._w16[:, 147:]) == 0
    opt.zero_grad()
    assert torch.count_nonzero(st.grad) == 0
    # lr schedulers of the reference drive it unchanged (optimizers.py:53-55)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=30)
    sched.step()
    opt.prepare()
    assert opt._hyper[0].item() == pytest.approx(opt.param_groups[0]['lr'])
