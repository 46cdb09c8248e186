"""End-to-end parity of one server train step against the torch restatement of the reference step (RefPCME forward ->
MCSoftContrastiveLoss -> backward -> clip_grad_norm_(2) -> AdamP) on identical weights and inputs, at the shape of
BASELINE.json configs[0] (16 pairs) AND configs[1] (batch 128, BERT length 32 - the shape bench.py times).
Reference: src/algorithms/retrieval_trainer.py:192-214.  Dropout is frozen (tests/test_gpu_dropout.py
covers the dropout path with exported masks).

Tolerances: loss rel 2e-2 (bf16 towers); clipped-gradient norm rel 5e-2; parameter updates of the head tensors
(well-conditioned gradients) cosine >= 0.98 against the oracle AdamP update; backbone gradients (ResNet101 layer 1-4
convolutions, BERT mid-stack) calibrated against torch's own bf16-autocast run of the same reference in the same test:
cos(ours, fp32) >= cos(autocast, fp32) - 0.02 and norm ratio within 10 %; every parameter stays finite and moves by
O(lr) per element."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


HEADS = ['img_enc.fc.weight', 'img_enc.pie_net.fc.weight', 'linear.weight', 'linear.bias',
         'img_enc.pie_net.layer_norm.weight', 'txt_enc.encoder.layer.11.output.dense.weight']
BACKBONE = ['img_enc.cnn.layer4.2.conv3.weight', 'img_enc.cnn.layer4.0.conv2.weight', 'img_enc.cnn.layer3.22.conv2.weight',
            'img_enc.cnn.layer3.10.conv2.weight', 'img_enc.cnn.layer3.10.conv1.weight', 'img_enc.cnn.layer3.0.downsample.0.weight',
            'img_enc.cnn.layer2.3.conv3.weight', 'img_enc.cnn.layer1.0.conv1.weight',
            'txt_enc.encoder.layer.5.intermediate.dense.weight', 'txt_enc.encoder.layer.0.attention.self.query.weight']


@pytest.mark.parametrize('B', [16, 128])
def test_server_train_step_matches_reference_step(B):
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    from creamfl_b200 import engine
    from oracle import creamfl_oracle as O, torch_towers as RT
    L, lr = 32, 2e-4
    ref = RT.RefPCME('resnet101', 256)
    RT.fill_deterministic(ref, seed=41)
    with torch.no_grad():
        for name, p in ref.named_parameters():
            if name.endswith('bn3.weight'):
                p.mul_(0.2)
    ref = ref.cuda().train()
    server = engine.ServerEngine(256, 'resnet101', lr=lr, grad_clip=2.0, bert_dropout=0.0)
    server.model.load_state_dict(ref.state_dict(), strict=True)
    server.model.sync_shadow()
    g = torch.Generator().manual_seed(42)
    images = torch.randn(B, 3, 224, 224, generator=g).cuda()
    lens = torch.sort(torch.randint(8, L + 1, (B,), generator=g), descending=True).values
    mask = (torch.arange(L)[None] < lens[:, None]).long()
    ids = torch.randint(1000, 30522, (B, L), generator=g)
    ids[:, 0] = 101
    ids, mask = (ids * mask).cuda(), mask.cuda()
    # ---- torch bf16-autocast run of the same reference (calibration of the backbone bar)
    amp = copy.deepcopy(ref)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        o_amp = amp(images, ids, mask, torch.zeros_like(ids))
    sh_a = torch.tensor(15.0, device='cuda', requires_grad=True)
    sc_a = torch.tensor(15.0, device='cuda', requires_grad=True)
    O.pcme_loss(o_amp['image_features'].float(), o_amp['caption_features'].float(), sh_a, sc_a)[0].backward()
    amp_grads = {n: p.grad.detach().clone() for n, p in amp.named_parameters() if n in BACKBONE}
    del amp, o_amp
    # ---- reference step (fp32 torch)
    shift = torch.tensor(15.0, device='cuda', requires_grad=True)
    scale = torch.tensor(15.0, device='cuda', requires_grad=True)
    out = ref(images, ids, mask, torch.zeros_like(ids))
    loss_ref, _ = O.pcme_loss(out['image_features'], out['caption_features'], shift, scale)
    loss_ref.backward()
    ref_params = dict(ref.named_parameters())
    model_grads = [p.grad for p in ref.parameters() if p.grad is not None]
    norm_ref = O.clip_grad_norm(model_grads, 2.0)            # scales the reference's gradients in place
    before = {n: ref_params[n].detach().clone() for n in HEADS}
    ps = [ref_params[n].detach().double().clone() for n in HEADS]
    gs = [ref_params[n].grad.double() for n in HEADS]
    O.adamp_step(ps, gs, [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps], 1, lr)
    # ---- CUDA step
    mine_params = dict(server.model.named_parameters())
    all_before = server.model.store().flat.clone()
    loss = server.train_step(images, {'input_ids': ids, 'attention_mask': mask})
    torch.cuda.synchronize()
    assert loss.item() == pytest.approx(loss_ref.item(), rel=2e-2)
    assert server.optimizer.grad_norm.item() == pytest.approx(norm_ref, rel=5e-2)
    for n, p_new in zip(HEADS, ps):
        d_ref = p_new.float().cuda() - before[n]
        d_mine = mine_params[n].detach() - before[n]
        assert cos(d_mine, d_ref) >= 0.98, (n, cos(d_mine, d_ref))
    # backbone: the gradients the step left in the flat buffer (zero_grad runs at the START of a step)
    bad = []
    for n in BACKBONE:
        g_mine, g_ref, g_amp = mine_params[n].grad, ref_params[n].grad, amp_grads[n]
        c, c_amp = cos(g_mine, g_ref), cos(g_amp, g_ref)
        # the reference's gradients were clipped in place (coefficient 2 / norm_ref); ours are kept un-clipped
        ratio = float(g_mine.double().norm() / g_ref.double().norm()) * min(1.0, 2.0 / norm_ref)
        if not (c >= c_amp - 0.02 and abs(ratio - 1) <= 0.10):
            bad.append((n, round(c, 4), round(c_amp, 4), round(ratio, 4)))
    assert not bad, bad
    moved = (server.model.store().flat - all_before).abs()
    assert torch.isfinite(server.model.store().flat).all()
    # first Adam step moves every element by ~lr; the AdamP projection can add a component along the weight
    assert moved.max().item() <= 50 * lr
    # criterion parameters are optimised too (retrieval_trainer.py:62-63) but not clipped
    assert server.criterion.shift.item() != 15.0 and abs(server.criterion.shift.item() - 15.0) <= lr + 2e-6  # 1 ulp at 15
