"""GPU parity tests of the encoder towers (creamfl_b200/towers.py) against the torch restatement of the reference's
towers (oracle/torch_towers.py: torchvision ResNet + HF BertModel + the reference's glue) on identical weights and
inputs.  The restatement runs in fp32 on the same GPU with TF32 off.

Tolerances (bf16 activations / bf16 tensor-core operands against an fp32 reference): embeddings
cosine >= 0.999 per row; parameter gradients: cosine >= 0.98 and norm ratio within 6 % for every parameter tensor
checked (deep-stack bf16 rounding noise accumulates through 100+ layers with batch-statistics BatchNorm).
"""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from creamfl_b200 import towers
    from oracle import torch_towers
    return towers, torch_towers


def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


def check_grads(pairs, cos_min=0.98, ratio_tol=0.06):
    bad = []
    for name, mine, ref in pairs:
        c = cos(mine, ref)
        ratio = (mine.double().norm() / ref.double().norm().clamp_min(1e-300)).item()
        if not (c >= cos_min and abs(ratio - 1) <= ratio_tol):
            bad.append((name, round(c, 4), round(ratio, 4)))
    assert not bad, bad


@pytest.mark.parametrize('arch,batch', [('resnet18', 8), ('resnet101', 4)])
def test_image_tower_train_step(env, arch, batch):
    towers, RT = env
    ref = RT.RefEncoderImage(arch, 256)
    RT.fill_deterministic(ref, seed=1)
    ref = ref.cuda().train()
    mine = towers.ImageModel({'embed_dim': 256, 'cnn_type': arch})
    missing = mine.img_enc.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda().train()
    g = torch.Generator().manual_seed(2)
    images = torch.randn(batch, 3, 224, 224, generator=g).cuda()
    cot = torch.randn(batch, 256, generator=g).cuda()
    # the reference sees bf16-rounded images and weights too, so the comparison isolates the arithmetic
    e_ref = ref(images)['embedding']
    (e_ref * cot).sum().backward()
    mine.zero_grad()
    e = mine(images)
    (e * cot).sum().backward()
    torch.cuda.synchronize()
    for i in range(batch):
        assert cos(e[i], e_ref[i]) >= 0.999, (i, cos(e[i], e_ref[i]))
    ref_p = dict(ref.named_parameters())
    names = ['fc.weight', 'fc.bias', 'pie_net.attention.w_1.weight', 'pie_net.attention.w_2.weight', 'pie_net.fc.weight',
             'pie_net.layer_norm.weight', 'cnn.layer4.1.conv2.weight', 'cnn.layer4.0.downsample.0.weight',
             'cnn.layer3.0.conv1.weight', 'cnn.layer2.0.conv2.weight', 'cnn.layer2.1.bn1.weight',
             'cnn.layer1.0.conv1.weight', 'cnn.layer1.0.bn1.bias', 'cnn.bn1.weight', 'cnn.conv1.weight']
    mine_p = dict(mine.img_enc.named_parameters())
    check_grads([(n, mine_p[n].grad, ref_p[n].grad) for n in names])
    # running statistics advance identically (momentum 0.1)
    assert cos(mine.img_enc.cnn.layer3[0].bn2.running_var, ref.cnn.layer3[0].bn2.running_var) > 0.9999
    assert cos(mine.img_enc.cnn.bn1.running_mean, ref.cnn.bn1.running_mean) > 0.9999


def test_image_tower_eval_and_deepcopy(env):
    towers, RT = env
    ref = RT.RefEncoderImage('resnet18', 256)
    RT.fill_deterministic(ref, seed=3)
    ref = ref.cuda().eval()
    mine = towers.ImageModel({'embed_dim': 256, 'cnn_type': 'resnet18'})
    mine.img_enc.load_state_dict(ref.state_dict())
    mine = mine.cuda().eval()
    images = torch.randn(4, 3, 224, 224, generator=torch.Generator().manual_seed(4)).cuda()
    with torch.no_grad():
        e_ref = ref(images)['embedding']
        e = mine(images)
        old = copy.deepcopy(mine)               # MMClientTrainer.py:92 deep-copies the model every round
        e_old = old(images)
    for i in range(4):
        assert cos(e[i], e_ref[i]) >= 0.999
    assert torch.equal(e, e_old)
    assert old.store() is not mine.store()


@pytest.mark.parametrize('batch,seq', [(8, 16), (4, 32)])
def test_pcme_train_step(env, batch, seq):
    towers, RT = env
    ref = RT.RefPCME('resnet18', 256)
    RT.fill_deterministic(ref, seed=5)
    ref = ref.cuda().train()
    mine = towers.PCME(None, {'embed_dim': 256, 'cnn_type': 'resnet18'})
    mine.load_state_dict(ref.state_dict(), strict=True)
    mine = mine.cuda().train()
    g = torch.Generator().manual_seed(6)
    images = torch.randn(batch, 3, 224, 224, generator=g).cuda()
    ids = torch.randint(1000, 30522, (batch, seq), generator=g)
    lens = torch.randint(4, seq + 1, (batch,), generator=g)
    lens[0] = seq
    mask = (torch.arange(seq)[None] < lens[:, None]).long()
    ids[:, 0] = 101
    ids = (ids * mask).cuda()
    mask = mask.cuda()
    cot_i, cot_t = torch.randn(batch, 256, generator=g).cuda(), torch.randn(batch, 256, generator=g).cuda()
    o_ref = ref(images, ids, mask, torch.zeros_like(ids))
    ((o_ref['image_features'] * cot_i).sum() + (o_ref['caption_features'] * cot_t).sum()).backward()
    mine.zero_grad()
    o = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
    assert set(o.keys()) == {'image_features', 'image_attentions', 'image_residuals', 'image_logsigma',
                             'image_logsigma_att', 'caption_features', 'caption_attentions', 'caption_residuals',
                             'caption_logsigma', 'caption_logsigma_att'}
    ((o['image_features'] * cot_i).sum() + (o['caption_features'] * cot_t).sum()).backward()
    torch.cuda.synchronize()
    for i in range(batch):
        assert cos(o['caption_features'][i], o_ref['caption_features'][i]) >= 0.9995
        assert cos(o['image_features'][i], o_ref['image_features'][i]) >= 0.999
    ref_p, mine_p = dict(ref.named_parameters()), dict(mine.named_parameters())
    names = ['linear.weight', 'linear.bias', 'txt_enc.encoder.layer.11.output.dense.weight',
             'txt_enc.encoder.layer.11.attention.self.query.weight', 'txt_enc.encoder.layer.6.attention.self.value.bias',
             'txt_enc.encoder.layer.6.intermediate.dense.weight', 'txt_enc.encoder.layer.6.intermediate.dense.bias',
             'txt_enc.encoder.layer.3.attention.output.LayerNorm.weight', 'txt_enc.encoder.layer.0.attention.self.key.weight',
             'txt_enc.encoder.layer.0.output.LayerNorm.bias', 'txt_enc.embeddings.LayerNorm.weight',
             'txt_enc.embeddings.position_embeddings.weight', 'txt_enc.embeddings.word_embeddings.weight',
             'txt_enc.embeddings.token_type_embeddings.weight', 'img_enc.fc.weight']
    check_grads([(n, mine_p[n].grad, ref_p[n].grad) for n in names])
    # the pooler is dead on this path (pcme.py:44): no gradient on either side
    assert ref_p['txt_enc.pooler.dense.weight'].grad is None or ref_p['txt_enc.pooler.dense.weight'].grad.abs().sum() == 0
    assert mine_p['txt_enc.pooler.dense.weight'].grad.abs().sum() == 0


def test_grad_accumulates_and_zero_grad_none(env):
    """Two backward passes accumulate; optimizer.zero_grad(set_to_none=True) semantics are honoured."""
    towers, RT = env
    mine = towers.ImageModel({'embed_dim': 256, 'cnn_type': 'resnet18'}).cuda().train()
    images = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(7)).cuda()
    mine.zero_grad()
    mine(images).sum().backward()
    w = mine.img_enc.cnn.layer2[0].conv1.weight
    g1 = w.grad.clone()
    for bn in [m for m in mine.modules() if isinstance(m, towers.BN)]:
        bn.momentum = 0.0                      # keep running stats fixed; batch stats are what the forward uses anyway
    mine(images).sum().backward()
    assert cos(w.grad, 2 * g1) > 0.9999 and abs(w.grad.norm().item() / (2 * g1.norm().item()) - 1) < 1e-3
    for p in mine.parameters():
        p.grad = None
    mine(images).sum().backward()
    assert cos(w.grad, g1) > 0.9999 and abs(w.grad.norm().item() / g1.norm().item() - 1) < 1e-3
