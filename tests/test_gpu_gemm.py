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
