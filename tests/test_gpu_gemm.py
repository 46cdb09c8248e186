"""GPU parity tests of the tcgen05 GEMM (creamfl_gemm_bf16) against an fp64 matmul of the same bf16 operands.

Tolerance: operands are exactly representable on both sides, the kernel accumulates in fp32 ->
rel-L2 <= 1e-5 for fp32 output, <= 4e-3 for bf16 output (one bf16 rounding, 2^-9 relative).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from creamfl_b200 import ops as _ops
    return _ops


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def mk(rows, cols, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) / cols ** 0.25).to(torch.bfloat16)


SHAPES = [(128, 128, 64), (4096, 2304, 768), (4096, 768, 3072), (6272, 1024, 2048), (200, 100, 72), (1, 8, 8),
          (129, 257, 136), (25088, 64, 256), (128, 256, 768), (300, 48, 1000)]


@pytest.mark.parametrize('m,n,k', SHAPES)
@pytest.mark.parametrize('a_mn,b_mn', [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_layouts(ops, m, n, k, a_mn, b_mn):
    if (a_mn and m % 8) or (b_mn and n % 8):
        pytest.skip('MN-major operand needs a 16-byte row pitch')
    a = mk(k, m, 1) if a_mn else mk(m, k, 1)
    b = mk(k, n, 2) if b_mn else mk(n, k, 2)
    ref = (a.double().t() if a_mn else a.double()) @ (b.double() if b_mn else b.double().t())
    out = ops.gemm_bf16(a.cuda(), b.cuda(), a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    assert rel_l2(out, ref) < 1e-5
    out16 = ops.gemm_bf16(a.cuda(), b.cuda(), a_mn=a_mn, b_mn=b_mn, out_dtype=torch.bfloat16)
    assert rel_l2(out16, ref) < 4e-3


@pytest.mark.parametrize('m,n,k,split', [(256, 256, 50000, 37), (768, 3072, 4096, 4), (64, 147, 12544, 16),
                                         (128, 256, 50048, 148)])
def test_gemm_split_k(ops, m, n, k, split):
    a, b = mk(k, m, 3), mk(k, n, 4)          # the wgrad form: both operands MN-major
    ref = a.double().t() @ b.double()
    if n % 8:
        pytest.skip('pitch')
    out = ops.gemm_bf16(a.cuda(), b.cuda(), a_mn=True, b_mn=True, split_k=split)
    assert out.dtype == torch.float32
    assert rel_l2(out, ref) < 1e-5


@pytest.mark.parametrize('act', ['none', 'gelu', 'relu', 'tanh', 'sigmoid'])
def test_gemm_epilogues(ops, act):
    m, n, k = 392, 328, 264
    a, b = mk(m, k, 5), mk(n, k, 6)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(7))
    add = mk(m, n, 8)
    pre = 0.5 * (a.double() @ b.double().t()) + bias.double() + add.double()
    fn = {'none': lambda x: x, 'gelu': lambda x: torch.nn.functional.gelu(x), 'relu': torch.relu,
          'tanh': torch.tanh, 'sigmoid': torch.sigmoid}[act]
    code = {'none': ops.ACT_NONE, 'gelu': ops.ACT_GELU, 'relu': ops.ACT_RELU, 'tanh': ops.ACT_TANH,
            'sigmoid': ops.ACT_SIGMOID}[act]
    out, pre_g = ops.gemm_bf16(a.cuda(), b.cuda(), bias=bias.cuda(), add=add.cuda(), act=code, alpha=0.5,
                               out_dtype=torch.float32, want_preact=True)
    assert rel_l2(out, fn(pre)) < (1e-3 if act == 'tanh' else 2e-5)      # tanh.approx.f32: 2^-11 relative
    assert rel_l2(pre_g, pre) < 4e-3
    # bf16 output: leaves through the staged TMA store
    out16 = ops.gemm_bf16(a.cuda(), b.cuda(), bias=bias.cuda(), add=add.cuda(), act=code, alpha=0.5)
    assert out16.dtype == torch.bfloat16 and rel_l2(out16, fn(pre)) < 4e-3


@pytest.mark.parametrize('m,n,k', [(392, 160, 264), (129, 64, 72), (1000, 768, 768), (4096, 2304, 768), (200, 256, 1024),
                                   (25088, 1024, 256)])
@pytest.mark.parametrize('b_mn', [False, True])
@pytest.mark.parametrize('bias,add', [(True, False), (False, True), (True, True), (False, False)])
def test_gemm_specialised_store_epilogue(ops, m, n, k, b_mn, bias, add):
    """The launches that qualify for the specialised bf16-store epilogue (alpha = 1, N % 32 == 0, optional bias, bf16
    residual through TMA; gemm_tc_lean.cu) against fp64: ragged M, every tile width, single CTA (M < 256) and CTA pair."""
    a = mk(m, k, 21)
    b = mk(k, n, 22) if b_mn else mk(n, k, 22)
    bv = torch.randn(n, generator=torch.Generator().manual_seed(23)) if bias else None
    av = mk(m, n, 24) if add else None
    ref = a.double() @ (b.double() if b_mn else b.double().t())
    if bias:
        ref = ref + bv.double()
    if add:
        ref = ref + av.double()
    out = ops.gemm_bf16(a.cuda(), b.cuda(), b_mn=b_mn, bias=bv.cuda() if bias else None, add=av.cuda() if add else None)
    assert out.dtype == torch.bfloat16 and rel_l2(out, ref) < 4e-3


def test_gelu_bf16_epilogue_accuracy(ops):
    """bf16 GELU / dGELU epilogues use the sigmoid-form normal CDF (max |error| 3.1e-5): against erf GELU in fp64 the
    result stays at the bf16 rounding level (rel-L2 4e-3), tails included (pre-activations up to |x| ~ 12)."""
    m, n, k = 512, 3072, 768
    g = torch.Generator().manual_seed(31)
    a = (torch.randn(m, k, generator=g) * 0.35).to(torch.bfloat16)
    b = (torch.randn(n, k, generator=g) * 0.35).to(torch.bfloat16)
    pre = a.double() @ b.double().t()
    assert pre.abs().max() > 8
    out, pre_g = ops.gemm_bf16(a.cuda(), b.cuda(), act=ops.ACT_GELU, want_preact=True)
    assert rel_l2(out, torch.nn.functional.gelu(pre)) < 4e-3
    ref = torch.nn.functional.gelu(pre)
    # element-wise: half a bf16 ulp (2^-8 relative at the bottom of a binade) + the approximation's 3.1e-5 / fp32 noise
    assert bool(((out.double().cpu() - ref).abs() <= 0.004 * ref.abs() + 2e-4).all())
    x = pre_g.double().cpu().requires_grad_(True)
    torch.nn.functional.gelu(x).sum().backward()
    dy = mk(m, k, 32)
    w = mk(k, n, 33)
    d = ops.gemm_bf16(dy.cuda(), w.cuda(), b_mn=True, act=ops.ACT_DGELU, aux=pre_g)
    assert rel_l2(d, (dy.double() @ w.double()) * x.grad) < 4e-3


@pytest.mark.parametrize('act', ['dgelu', 'drelu'])
def test_gemm_backward_epilogues(ops, act):
    m, n, k = 256, 3072, 768
    a, b = mk(m, k, 9), mk(k, n, 10)
    aux = mk(m, n, 11)
    acc = a.double() @ b.double()
    x = aux.double().requires_grad_(True)
    (torch.nn.functional.gelu(x) if act == 'dgelu' else torch.relu(x)).sum().backward()
    ref = acc * x.grad
    out = ops.gemm_bf16(a.cuda(), b.cuda(), b_mn=True, aux=aux.cuda(), out_dtype=torch.float32,
                        act=ops.ACT_DGELU if act == 'dgelu' else ops.ACT_DRELU)
    assert rel_l2(out, ref) < 2e-5


def test_gemm_linearity_full_size(ops):
    """Size-independent property at the BERT FFN size: G(a1 + a2, b) = G(a1, b) + G(a2, b) when the sums are
    exact in bf16 (operands restricted to a few mantissa bits)."""
    g = torch.Generator().manual_seed(12)
    q = lambda r, c: (torch.randint(-8, 9, (r, c), generator=g).float() / 8).to(torch.bfloat16)
    a1, a2, b = q(4096, 768), q(4096, 768), q(3072, 768)
    o1 = ops.gemm_bf16(a1.cuda(), b.cuda(), out_dtype=torch.float32)
    o2 = ops.gemm_bf16(a2.cuda(), b.cuda(), out_dtype=torch.float32)
    o12 = ops.gemm_bf16((a1 + a2).cuda(), b.cuda(), out_dtype=torch.float32)
    assert torch.equal(o12, o1 + o2)     # every partial sum is a small dyadic rational: exact in fp32


def test_gemm_rejects_bad_arguments(ops):
    a = torch.zeros(16, 12, dtype=torch.bfloat16, device='cuda')    # pitch 24 B: not TMA-addressable
    with pytest.raises(RuntimeError):
        ops.gemm_bf16(a, a)
    with pytest.raises(TypeError):
        ops.gemm_bf16(a.float(), a.float())
