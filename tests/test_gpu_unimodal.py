"""GPU parity tests of the unimodal image client (resnet18_client mirror, supervised step of ClientTrainer.tra, and the
contrast-phase embedding) against the torch restatement in oracle/torch_towers.py; and of the CE / pool primitives.

Tolerances: CE kernel vs fp64 torch rel 1e-5; client losses rel 2e-2 (bf16 trunk), head gradients calibrated against
torch bf16 autocast as in tests/test_gpu_towers.py."""
import copy

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from creamfl_b200 import clients, ops, optim
    from oracle import torch_towers
    return clients, ops, optim, torch_towers


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


@pytest.mark.parametrize('r,c,margin,with_labels', [(512, 100, 4.0, True), (100, 100, 0.0, False), (7, 4, 4.0, True),
                                                    (33, 1000, 0.0, True)])
def test_cross_entropy_kernel(env, r, c, margin, with_labels):
    _, ops, _, _ = env
    g = torch.Generator().manual_seed(r)
    x = torch.randn(r, c, generator=g, dtype=torch.float64) * 3
    labels = torch.randint(0, c, (r,), generator=g) if with_labels else torch.arange(r)
    xo = x.clone().requires_grad_(True)
    lo = F.cross_entropy(xo - margin * F.one_hot(labels, c), labels)
    lo.backward()
    xg = x.float().cuda().requires_grad_(True)
    lg = ops.cross_entropy(xg, labels.cuda() if with_labels else None, margin)
    (lg * 2).backward()
    assert lg.item() == pytest.approx(lo.item(), rel=1e-5)
    assert ((xg.grad.cpu().double() / 2 - xo.grad).norm() / xo.grad.norm()).item() < 1e-5


def test_image_client_supervised_step_and_embedding(env):
    clients, ops, optim, RT = env
    ref = RT.RefImageClient(num_class=100, embed_dim=256)
    RT.fill_deterministic(ref, seed=21)
    with torch.no_grad():
        for name, p in ref.named_parameters():
            if name.endswith('bn2.weight'):
                p.mul_(0.2)
        ref.linear.weight.mul_(0.05)          # keep `x * 128` logits in a sane range, as a trained client would
        ref.class_fc_2.weight.mul_(0.3)
    ref = ref.cuda().train()
    mine = clients.resnet18_client(num_class=100, embed_dim=256, scale=128, is_train=True)
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda().train()
    g = torch.Generator().manual_seed(22)
    images = torch.randn(16, 3, 64, 64, generator=g).cuda()
    labels = torch.randint(0, 100, (16,), generator=g).cuda()
    amp = copy.deepcopy(ref)
    l_ref, f_ref = RT.ref_unimodal_supervised_loss(ref, images, labels, 100)
    l_ref.backward()
    with torch.autocast('cuda', dtype=torch.bfloat16):
        l_amp, _ = RT.ref_unimodal_supervised_loss(amp, images, labels, 100)
    l_amp.float().backward()
    mine.zero_grad()
    l, fvec = clients.unimodal_supervised_loss(mine, images, labels, 4.0)
    l.backward()
    torch.cuda.synchronize()
    assert l.item() == pytest.approx(l_ref.item(), rel=2e-2)
    assert cos(fvec, f_ref + 4.0 * F.one_hot(labels, 100)) > 0.999          # ours returns the un-shifted logits
    # the ReLU clamp is a side effect of forward on both sides (resnet_client.py:193-197)
    assert (mine.class_fc_2.weight.data >= 0).all() and torch.equal(mine.class_fc_2.weight._w16,
                                                                    mine.class_fc_2.weight.data.to(torch.bfloat16))
    rp, mp, ap = dict(ref.named_parameters()), dict(mine.named_parameters()), dict(amp.named_parameters())
    for n in ['class_fc_2.weight', 'class_fc_2.bias', 'linear.weight', 'linear.bias', 'layer4.1.conv2.weight',
              'layer2.0.downsample.0.weight', 'conv1.weight']:
        c, c_amp = cos(mp[n].grad, rp[n].grad), cos(ap[n].grad, rp[n].grad)
        ratio = (mp[n].grad.norm() / rp[n].grad.norm()).item()
        assert c >= c_amp - 0.03 and abs(ratio - 1) < 0.1, (n, c, c_amp, ratio)
    assert mp['class_fc_22.weight'].grad.abs().sum() == 0                  # second head is returned but unused
    # SGD(lr 1e-4, momentum 0.9, wd 5e-5) of ClientTrainer.py:287-288 through the fused optimizer
    opt = optim.FusedOptimizer(mine.parameters(), lr=1e-4, momentum=0.9, weight_decay=5e-5, mode='sgd').attach_stores(mine)
    before = mine.linear.weight.data.clone()
    opt.step()
    want = before - 1e-4 * (mp['linear.weight'].grad + 5e-5 * before)
    assert torch.allclose(mine.linear.weight.data, want, rtol=1e-5, atol=1e-9)
    # contrast phase: L2-normalised embedding (ClientTrainer.py:372-381)
    ref.phase = mine.phase = 'extract_conv_feature'
    ref.is_train = mine.is_train = False
    ref.load_state_dict(mine.state_dict())
    with torch.no_grad():
        e_ref, e = ref(images), mine(images)
    for i in range(16):
        assert cos(e[i], e_ref[i]) > 0.999
