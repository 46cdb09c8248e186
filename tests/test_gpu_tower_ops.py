"""GPU parity tests of the encoder-tower primitives (convolution, BatchNorm, pooling, LayerNorm, BERT attention and
embeddings, PIE pooling) against fp64 torch on the CPU fed the same bf16-rounded operands.

The reference reaches these operations through torchvision / HF transformers: the oracle for them is
torch's own functional ops in fp64.  Tolerances: bf16 outputs rel-L2 <= 4e-3 (one rounding of an fp32
accumulator), fp32 outputs (weight / parameter gradients) rel-L2 <= 2e-4 unless noted.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def T():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from creamfl_b200 import tower_ops
    return tower_ops


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16)


CONVS = [  # n, h, w, cin, cout, r, stride, pad
    (2, 56, 56, 64, 64, 3, 1, 1), (4, 28, 28, 128, 128, 3, 1, 1), (32, 14, 14, 256, 256, 3, 1, 1),
    (3, 14, 14, 256, 256, 3, 1, 1), (130, 7, 7, 512, 512, 3, 1, 1), (2, 56, 56, 64, 256, 1, 1, 0),
    (2, 56, 56, 256, 64, 1, 1, 0), (2, 56, 56, 128, 128, 3, 2, 1), (2, 56, 56, 256, 512, 1, 2, 0),
    (5, 14, 14, 1024, 2048, 1, 2, 0), (2, 14, 14, 512, 512, 3, 2, 1), (2, 12, 20, 64, 128, 3, 1, 1),
]


@pytest.mark.parametrize('n,h,w,cin,cout,r,stride,pad', CONVS)
def test_conv_fprop_dgrad_wgrad(T, n, h, w, cin, cout, r, stride, pad):
    x = rnd(n, h, w, cin, seed=1)
    wt = rnd(cout, r, r, cin, seed=2, scale=(r * r * cin) ** -0.5)
    x64 = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    w64 = wt.double().permute(0, 3, 1, 2).requires_grad_(True)
    y64 = F.conv2d(x64, w64, stride=stride, padding=pad)
    dy = rnd(*y64.permute(0, 2, 3, 1).shape, seed=3)
    y64.backward(dy.double().permute(0, 3, 1, 2))
    w2d = wt.reshape(cout, -1).cuda()
    y = T.conv_fprop(x.cuda(), w2d, r, r, stride, pad)
    assert y.shape == dy.shape
    assert rel_l2(y, y64.permute(0, 2, 3, 1)) < 4e-3
    dx = T.conv_dgrad(dy.cuda(), w2d, x.shape, r, r, stride, pad)
    assert rel_l2(dx, x64.grad.permute(0, 2, 3, 1)) < 4e-3
    add = rnd(*x.shape, seed=4)
    dx2 = T.conv_dgrad(dy.cuda(), w2d, x.shape, r, r, stride, pad, add=add.cuda())
    assert rel_l2(dx2, x64.grad.permute(0, 2, 3, 1) + add.double()) < 4e-3
    dw = torch.full((cout, r * r * cin), 0.5, device='cuda')       # accumulate semantics
    T.conv_wgrad(dy.cuda(), x.cuda(), dw, r, r, stride, pad)
    assert rel_l2(dw - 0.5, w64.grad.permute(0, 2, 3, 1).reshape(cout, -1)) < 2e-4


@pytest.mark.parametrize('n,h,w,cin,cout,r,stride,pad', [(128, 14, 14, 1024, 256, 1, 1, 0), (16, 56, 56, 64, 256, 1, 1, 0),
                                                         (4, 28, 28, 128, 128, 3, 1, 1), (8, 56, 56, 256, 512, 1, 2, 0),
                                                         (3, 14, 14, 256, 1024, 1, 1, 0), (2, 56, 56, 64, 64, 1, 1, 0)])
def test_conv_fused_bn_statistics(T, n, h, w, cin, cout, r, stride, pad):
    """BatchNorm statistics produced by the convolution (GEMM epilogue or stand-alone pass, chosen by the library)
    equal the per-channel sum / sum of squares of the bf16 output it wrote: rel 1e-5 (fp32 partials, fp64 totals)."""
    x = rnd(n, h, w, cin, seed=31)
    wt = rnd(cout, r * r * cin, seed=32, scale=(r * r * cin) ** -0.5)
    sums = torch.zeros(2 * cout, dtype=torch.float64, device='cuda')
    y = T.conv_fprop(x.cuda(), wt.cuda(), r, r, stride, pad, bn_sums=sums)
    y2 = T.conv_fprop(x.cuda(), wt.cuda(), r, r, stride, pad)
    assert torch.equal(y, y2)
    yd = y.double().reshape(-1, cout)
    assert rel_l2(sums[:cout], yd.sum(0)) < 1e-5
    assert rel_l2(sums[cout:], (yd * yd).sum(0)) < 1e-5


def test_conv_is_linear_full_size(T):
    """Size-independent property at the ResNet101 layer3 size (B = 128): conv(x1 + x2) == conv(x1) + conv(x2) exactly
    when all partial sums are dyadic rationals."""
    g = torch.Generator().manual_seed(5)
    q = lambda *s: (torch.randint(-4, 5, s, generator=g).float() / 4).to(torch.bfloat16).cuda()
    x1, x2, wt = q(128, 14, 14, 256), q(128, 14, 14, 256), q(256, 3 * 3 * 256)
    y1, y2 = T.conv_fprop(x1, wt, 3, 3, 1, 1), T.conv_fprop(x2, wt, 3, 3, 1, 1)
    y12 = T.conv_fprop(x1 + x2, wt, 3, 3, 1, 1)
    # outputs are rounded to bf16: compare with one-ulp slack on the sum of two rounded values
    assert rel_l2(y12, y1.float() + y2.float()) < 4e-3


def test_stem_im2col(T):
    from creamfl_b200 import ops
    img = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(6))
    wt = rnd(64, 7, 7, 3, seed=7, scale=0.1)
    col = T.im2col_images(img.cuda(), 7, 7, 2, 3, 152)
    w2d = torch.zeros(64, 152, dtype=torch.bfloat16)
    w2d[:, :147] = wt.reshape(64, 147)
    y = ops.gemm_bf16(col, w2d.cuda())
    ref = F.conv2d(img.to(torch.bfloat16).double(), wt.double().permute(0, 3, 1, 2), stride=2, padding=3)
    assert rel_l2(y.reshape(3, 32, 32, 64), ref.permute(0, 2, 3, 1)) < 4e-3
    assert torch.count_nonzero(col[:, 147:]) == 0


@pytest.mark.parametrize('shape', [(4, 14, 14, 256), (2, 56, 56, 64), (3, 7, 7, 2048), (128, 14, 14, 1024)])
@pytest.mark.parametrize('relu,with_res', [(True, False), (True, True), (False, False)])
def test_batchnorm_train(T, shape, relu, with_res):
    c = shape[-1]
    x = rnd(*shape, seed=8, scale=2.0) + 0.5
    res = rnd(*shape, seed=9) if with_res else None
    g = torch.Generator().manual_seed(10)
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    rm, rv = torch.zeros(c), torch.ones(c)
    x64 = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm64, rv64 = rm.double(), rv.double()
    y64 = F.batch_norm(x64, rm64, rv64, g64, b64, training=True, momentum=0.1, eps=1e-5)
    if with_res:
        y64 = y64 + res.double().permute(0, 3, 1, 2)
    if relu:
        y64 = torch.relu(y64)
    sc = T.BNScratch(c, 'cuda')
    rmg, rvg = rm.cuda(), rv.cuda()
    nbt = torch.tensor(41, dtype=torch.long, device='cuda')
    y, mean, rstd = T.bn_train_fwd(x.cuda(), gamma.cuda(), beta.cuda(), rmg, rvg, sc, 1e-5, 0.1,
                                   res=res.cuda() if with_res else None, relu=relu, num_batches_tracked=nbt)
    assert nbt.item() == 42                      # nn.BatchNorm2d's counter rides in the finalize kernel
    assert rel_l2(y, y64.permute(0, 2, 3, 1)) < 4e-3
    assert rel_l2(rmg, rm64) < 1e-5 and rel_l2(rvg, rv64) < 1e-5
    dy = rnd(*shape, seed=11)
    # use the CUDA output as the ReLU mask on both sides (bf16 rounding can move values across zero)
    mask = (y.cpu() > 0).double().permute(0, 3, 1, 2) if relu else None
    pre = F.batch_norm(x64, None, None, g64, b64, training=True, eps=1e-5)
    (pre * (dy.double().permute(0, 3, 1, 2) * (mask if relu else 1.0))).sum().backward()
    dg, db = torch.full((c,), 0.25, device='cuda'), torch.full((c,), 0.25, device='cuda')
    dx, gout = T.bn_train_bwd(dy.cuda(), y if relu else None, x.cuda(), gamma.cuda(), mean, rstd, sc, dg, db,
                              want_g=True)
    if relu and not with_res:      # gate recomputed from x instead of read from y: same result up to sign ties at 0
        dg2, db2 = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
        dx2, _ = T.bn_train_bwd(dy.cuda(), None, x.cuda(), gamma.cuda(), mean, rstd, sc, dg2, db2, beta=beta.cuda(),
                                relu_from_x=True)
        assert rel_l2(dx2, dx) < 2e-3 and rel_l2(dg2, dg - 0.25) < 2e-3
    assert rel_l2(dx, x64.grad.permute(0, 2, 3, 1)) < 6e-3
    assert rel_l2(dg - 0.25, g64.grad) < 1e-3 and rel_l2(db - 0.25, b64.grad) < 1e-3
    assert rel_l2(gout, dy.double() * (mask.permute(0, 2, 3, 1) if relu else 1.0)) < 1e-6
    assert torch.count_nonzero(sc.sums) == 0
    if relu and with_res:
        # gate handed over as one bit per element (creamfl_bn_train_fwd_mask / _bwd_mask): same forward output, the bits
        # equal (y > 0) except where the fp32 pre-activation is positive but rounds to a bf16 zero, same backward
        rm2, rv2 = rm.cuda(), rv.cuda()
        y2, mean2, rstd2, bits = T.bn_train_fwd(x.cuda(), gamma.cuda(), beta.cuda(), rm2, rv2, sc, 1e-5, 0.1,
                                                res=res.cuda(), relu=True, want_mask=True)
        assert torch.equal(y2, y) and torch.equal(mean2, mean)
        unpacked = ((bits[:, None] >> torch.arange(8, dtype=torch.uint8, device='cuda')) & 1).reshape(y.shape).bool()
        assert (unpacked != (y > 0)).float().mean().item() < 1e-5
        dg3, db3 = torch.full((c,), 0.25, device='cuda'), torch.full((c,), 0.25, device='cuda')
        dx3, g3 = T.bn_train_bwd(dy.cuda(), None, x.cuda(), gamma.cuda(), mean, rstd, sc, dg3, db3, want_g=True, mask=bits)
        assert rel_l2(dx3, dx) < 1e-4 and rel_l2(g3, gout) < 1e-4 and rel_l2(dg3, dg) < 1e-4 and rel_l2(db3, db) < 1e-4
        assert torch.count_nonzero(sc.sums) == 0


def test_batchnorm_eval(T):
    shape, c = (4, 14, 14, 256), 256
    x = rnd(*shape, seed=12)
    g = torch.Generator().manual_seed(13)
    gamma, beta, rm, rv = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g), torch.randn(c, generator=g), \
        torch.rand(c, generator=g) + 0.5
    ref = torch.relu(F.batch_norm(x.double().permute(0, 3, 1, 2), rm.double(), rv.double(), gamma.double(),
                                  beta.double(), training=False, eps=1e-5))
    y = T.bn_eval_fwd(x.cuda(), gamma.cuda(), beta.cuda(), rm.cuda(), rv.cuda(), T.BNScratch(c, 'cuda'), 1e-5)
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < 4e-3


@pytest.mark.parametrize('shape', [(2, 112, 112, 64), (3, 13, 9, 16)])
def test_maxpool(T, shape):
    x = rnd(*shape, seed=14)
    x64 = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    y64 = F.max_pool2d(x64, 3, 2, 1)
    y, idx = T.maxpool_fwd(x.cuda())
    assert torch.equal(y.cpu().double(), y64.permute(0, 2, 3, 1))
    # distinct values almost surely -> unique argmax -> gradients must agree exactly up to bf16 rounding of sums
    dy = rnd(*y.shape, seed=15)
    y64.backward(dy.double().permute(0, 3, 1, 2))
    dx = T.maxpool_bwd(dy.cuda(), idx, x.shape)
    assert rel_l2(dx, x64.grad.permute(0, 2, 3, 1)) < 4e-3


@pytest.mark.parametrize('r,d,dtype', [(4096, 768, torch.bfloat16), (128, 256, torch.float32), (7, 1024, torch.bfloat16),
                                       (100, 8, torch.float32)])
@pytest.mark.parametrize('with_res', [False, True])
def test_layernorm(T, r, d, dtype, with_res):
    g = torch.Generator().manual_seed(16)
    x = torch.randn(r, d, generator=g).to(dtype)
    res = torch.randn(r, d, generator=g).to(dtype) if with_res else None
    gamma, beta = torch.rand(d, generator=g) + 0.5, torch.randn(d, generator=g) * 0.1
    dy = torch.randn(r, d, generator=g).to(dtype)
    xin = (x.double() + (res.double() if with_res else 0)).requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    y64 = F.layer_norm(xin, (d,), g64, b64, 1e-12)
    y64.backward(dy.double())
    y, mean, rstd = T.layernorm_fwd(x.cuda(), gamma.cuda(), beta.cuda(), 1e-12, res=res.cuda() if with_res else None)
    tol = 4e-3 if dtype == torch.bfloat16 else 1e-5
    assert rel_l2(y, y64) < tol
    dg, db = torch.ones(d, device='cuda'), torch.ones(d, device='cuda')
    dxs = torch.full((d,), 2.0, device='cuda')
    dx = T.layernorm_bwd(dy.cuda(), x.cuda(), gamma.cuda(), mean, rstd, dg, db, res=res.cuda() if with_res else None,
                         dx_colsum=dxs)
    assert rel_l2(dxs - 2, xin.grad.sum(0)) < (2e-3 if dtype == torch.bfloat16 else 1e-4)
    assert rel_l2(dx, xin.grad) < (6e-3 if dtype == torch.bfloat16 else 1e-4)
    assert rel_l2(dg - 1, g64.grad) < 1e-4 and rel_l2(db - 1, b64.grad) < 1e-4


def test_colsum_add_actbwd(T):
    x = rnd(4096, 2304, seed=17)
    out = torch.ones(2304, device='cuda')
    T.colsum_into(x.cuda(), out)
    assert rel_l2(out - 1, x.double().sum(0)) < 1e-5
    a, b = rnd(1000, 64, seed=18), rnd(1000, 64, seed=19)
    assert rel_l2(T.add_bf16(a.cuda(), b.cuda()), a.double() + b.double()) < 4e-3
    g = torch.Generator().manual_seed(20)
    dy, y = torch.randn(128, 256, generator=g), torch.rand(128, 256, generator=g)
    assert rel_l2(T.act_bwd(dy.cuda(), y.cuda(), 6), dy.double() * y.double() * (1 - y.double())) < 4e-3
    assert rel_l2(T.act_bwd(dy.cuda(), y.cuda(), 3), dy.double() * (1 - y.double() ** 2)) < 4e-3


@pytest.mark.parametrize('b,l,h', [(128, 32, 12), (3, 17, 2), (2, 64, 12), (5, 1, 1)])
def test_attention(T, b, l, h):
    qkv = rnd(b * l, 3 * h * 64, seed=21, scale=0.7)
    g = torch.Generator().manual_seed(22)
    lens = torch.randint(1, l + 1, (b,), generator=g)
    lens[0] = l
    mask = (torch.arange(l)[None, :] < lens[:, None]).float()
    q64 = qkv.double().requires_grad_(True)
    t = q64.reshape(b, l, 3, h, 64).permute(2, 0, 3, 1, 4)
    scores = t[0] @ t[1].transpose(-1, -2) / 8.0 + (1 - mask.double())[:, None, None, :] * torch.finfo(torch.float32).min
    p = torch.softmax(scores, -1)
    ctx64 = (p @ t[2]).permute(0, 2, 1, 3).reshape(b * l, h * 64)
    dctx = rnd(b * l, h * 64, seed=23)
    ctx64.backward(dctx.double())
    ctx, probs = T.attn_fwd(qkv.cuda(), mask.cuda(), b, l, h)
    assert rel_l2(ctx, ctx64) < 6e-3
    assert rel_l2(probs, p) < 4e-3
    dbias = torch.ones(3 * h * 64, device='cuda')
    dqkv = T.attn_bwd(qkv.cuda(), probs, dctx.cuda(), b, l, h, dbias=dbias)
    assert rel_l2(dqkv, q64.grad) < 1e-2
    assert rel_l2(dbias - 1, q64.grad.sum(0)) < 1e-2


def test_embeddings(T):
    g = torch.Generator().manual_seed(24)
    v, l, d, b = 1000, 32, 768, 16
    word, pos, typ = torch.randn(v, d, generator=g), torch.randn(512, d, generator=g), torch.randn(2, d, generator=g)
    ids = torch.randint(0, v, (b, l), generator=g)
    tt = torch.randint(0, 2, (b, l), generator=g)
    ref = word[ids] + pos[:l][None] + typ[tt]
    out = T.embed_fwd(ids.cuda(), tt.cuda(), word.cuda(), pos.cuda(), typ.cuda(), l)
    assert rel_l2(out.reshape(b, l, d), ref) < 4e-3
    dh = rnd(b * l, d, seed=25)
    dw, dp, dt = torch.zeros(v, d, device='cuda'), torch.zeros(512, d, device='cuda'), torch.zeros(2, d, device='cuda')
    T.embed_bwd(ids.cuda(), tt.cuda(), dh.cuda(), l, dw, dp, dt)
    rw = torch.zeros(v, d, dtype=torch.float64).index_add_(0, ids.reshape(-1), dh.double())
    rp = dh.double().reshape(b, l, d).sum(0)
    assert rel_l2(dw, rw) < 1e-5 and rel_l2(dp[:l], rp) < 1e-5 and torch.count_nonzero(dp[l:]) == 0
    assert rel_l2(dt, torch.zeros(2, d, dtype=torch.float64).index_add_(0, tt.reshape(-1), dh.double())) < 1e-5


@pytest.mark.parametrize('b,p,c,hd', [(128, 49, 2048, 1024), (3, 49, 512, 256), (2, 30, 304, 152)])
def test_pie_pool(T, b, p, c, hd):
    x = rnd(b, p, c, seed=26)
    h = torch.tanh(rnd(b, p, hd, seed=27).float()).to(torch.bfloat16)
    w2 = torch.randn(hd, generator=torch.Generator().manual_seed(28)) * 0.2
    x64, h64, w64 = x.double().requires_grad_(True), h.double().requires_grad_(True), w2.double().requires_grad_(True)
    a64 = torch.softmax(h64 @ w64, dim=1)
    r64 = (a64[:, :, None] * x64).sum(1)
    m64 = x64.mean(1)
    attn, r, pooled = T.pie_pool_fwd(x.cuda(), h.cuda(), w2.cuda())
    assert rel_l2(attn, a64) < 1e-4 and rel_l2(r, r64) < 4e-3 and rel_l2(pooled, m64) < 4e-3
    d_r, d_m = rnd(b, c, seed=29), rnd(b, c, seed=30)
    ((r64 * d_r.double()).sum() + (m64 * d_m.double()).sum()).backward()
    dw2 = torch.zeros(hd, device='cuda')
    dx, dpre = T.pie_pool_bwd(x.cuda(), h.cuda(), w2.cuda(), attn, d_r.cuda(), d_m.cuda(), dw2)
    assert rel_l2(dx, x64.grad) < 6e-3
    assert rel_l2(dpre, h64.grad * (1 - h.double() ** 2)) < 6e-3
    assert rel_l2(dw2, w64.grad) < 1e-3


@pytest.mark.parametrize('n,h,w', [(4, 224, 224), (3, 64, 64), (2, 256, 256), (150, 32, 48), (2, 225, 132)])
def test_fused_stem_matches_conv7x7(T, n, h, w):
    """csrc/stem_tc.cu (patch operand assembled in shared memory) against fp64 conv2d / its weight gradient on the same
    bf16-rounded operands, and against the materialised im2col + GEMM path it replaces."""
    from creamfl_b200 import ops
    assert T.stem_supported(h, w)
    g = torch.Generator().manual_seed(5)
    images = torch.randn(n, 3, h, w, generator=g)
    wt = (torch.randn(64, 3, 7, 7, generator=g) * (2.0 / 147) ** 0.5).to(torch.bfloat16)
    w16 = torch.zeros(64, 152, dtype=torch.bfloat16)
    w16[:, :147] = wt.permute(0, 2, 3, 1).reshape(64, 147)              # (r, s, c) column order, zero tail to the pitch
    y = T.stem_fprop(images.cuda(), w16.cuda())
    ref = F.conv2d(images.to(torch.bfloat16).double(), wt.double(), stride=2, padding=3).permute(0, 2, 3, 1)
    assert y.shape == ref.shape and rel_l2(y, ref) < 4e-3
    col = T.im2col_images(images.cuda(), 7, 7, 2, 3, 152)
    y_old = ops.gemm_bf16(col, w16.cuda()).view(y.shape)
    assert rel_l2(y, y_old) < 1e-6 or torch.equal(y, y_old)             # same operands, same fp32 accumulation
    dy = rnd(*y.shape, seed=6)
    dw = torch.full((64, 147), 0.5, device='cuda')
    T.stem_wgrad(images.cuda(), dy.cuda(), dw)
    xd = images.to(torch.bfloat16).double()
    wd = wt.double().requires_grad_(True)
    (F.conv2d(xd, wd, stride=2, padding=3) * dy.double().permute(0, 3, 1, 2)).sum().backward()
    want = wd.grad.permute(0, 2, 3, 1).reshape(64, 147)
    assert rel_l2(dw - 0.5, want) < 2e-4


def test_fused_stem_refuses_wide_images(T):
    assert not T.stem_supported(300, 300) and not T.stem_supported(64, 130)     # rows travel as 16-byte chunks
    with pytest.raises(RuntimeError):
        T.stem_fprop(torch.randn(1, 3, 300, 300).cuda(), torch.zeros(64, 152, dtype=torch.bfloat16).cuda())


@pytest.mark.parametrize('shape,training', [((4, 112, 112, 64), True), ((3, 32, 32, 64), True), ((2, 14, 10, 64), True),
                                            ((4, 112, 112, 64), False)])
def test_fused_stem_tail_matches_unfused_sequence(T, shape, training):
    """creamfl_bn_train_stats / _eval_affine + creamfl_maxpool_affine_fwd and creamfl_bn_pool_bwd against the separate
    passes they replace (BatchNorm apply + ReLU -> maxpool; maxpool backward -> BatchNorm backward with the gate from
    x).  Forward: the un-fused path rounds the normalised map to bf16 before pooling, the fused one pools fp32 values:
    equal up to one bf16 rounding (rel-L2 4e-3), winning taps equal except near-ties (< 0.5 %).  Backward: the
    un-fused path rounds the up-sampled gradient to bf16, rel-L2 6e-3; parameter gradients 2e-3."""
    from creamfl_b200 import towers
    n, h, w, c = shape
    x = (rnd(*shape, seed=51, scale=1.5) + 0.25).cuda()
    bn_a, bn_b = towers.BN(c).cuda(), towers.BN(c).cuda()
    g = torch.Generator().manual_seed(52)
    for bn in (bn_a, bn_b):
        bn.weight.data.copy_(torch.rand(c, generator=torch.Generator().manual_seed(53)) + 0.5)
        bn.bias.data.copy_(torch.randn(c, generator=torch.Generator().manual_seed(54)) * 0.2)
        bn.running_mean.copy_(torch.randn(c, generator=torch.Generator().manual_seed(55)) * 0.1)
        bn.running_var.copy_(torch.rand(c, generator=torch.Generator().manual_seed(56)) + 0.5)
        bn.train(training)
    # un-fused reference sequence
    if training:
        a, mean, rstd = T.bn_train_fwd(x, bn_a.weight, bn_a.bias, bn_a.running_mean, bn_a.running_var, bn_a.scratch(),
                                       bn_a.eps, bn_a.momentum, relu=True, num_batches_tracked=bn_a.num_batches_tracked)
    else:
        a = T.bn_eval_fwd(x, bn_a.weight, bn_a.bias, bn_a.running_mean, bn_a.running_var, bn_a.scratch(), bn_a.eps, relu=True)
    y_ref, idx_ref = T.maxpool_fwd(a, want_idx=True)
    y, idx, saved = T.stem_tail_fwd(x, bn_b, training, want_idx=True)
    assert rel_l2(y, y_ref) < 4e-3
    assert (idx != idx_ref).float().mean().item() < 5e-3
    if not training:
        assert saved is None
        return
    assert torch.equal(saved[0], mean) and torch.equal(saved[1], rstd)
    assert torch.equal(bn_a.running_var, bn_b.running_var) and int(bn_b.num_batches_tracked) == 1
    dy = rnd(*y.shape, seed=57).cuda()
    dg_ref, db_ref = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    da = T.maxpool_bwd(dy, idx, a.shape)                 # same winning taps on both sides
    do_ref, _ = T.bn_train_bwd(da, None, x, bn_a.weight, mean, rstd, bn_a.scratch(), dg_ref, db_ref, beta=bn_a.bias,
                               relu_from_x=True)
    dg, db = torch.zeros(c, device='cuda'), torch.zeros(c, device='cuda')
    do = T.stem_tail_bwd(dy, idx, x, bn_b, saved, dg, db)
    assert rel_l2(do, do_ref) < 6e-3
    assert rel_l2(dg, dg_ref) < 2e-3 and rel_l2(db, db_ref) < 2e-3
    assert torch.count_nonzero(bn_b.scratch().sums) == 0
