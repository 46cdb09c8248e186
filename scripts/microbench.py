"""Kernel-level timings (CUDA events, L2 flushed between iterations) of the C-ABI entry points.  Development
aid; bench.py is the contract benchmark."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


res = {}
which = sys.argv[1:] or ['gemm', 'conw', 'infonce', 'reduce']
if 'gemm' in which:
    for (m, n, k) in [(4096, 2304, 768), (4096, 3072, 768), (4096, 768, 3072), (8192, 8192, 8192), (25088, 256, 1024),
                      (25088, 1024, 256), (401408, 64, 256), (6272, 2048, 512)]:
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        b = torch.randn(n, k, device=dev).to(torch.bfloat16)
        ms = timeit(lambda: ops.gemm_bf16(a, b))
        ms_t = timeit(lambda: torch.matmul(a, b.t()))
        res[f'gemm_{m}x{n}x{k}'] = {'ms': ms, 'tflops': 2 * m * n * k / ms / 1e9, 'torch_ms': ms_t,
                                    'torch_tflops': 2 * m * n * k / ms_t / 1e9}
if 'epi' in which:
    # epilogue-heavy shapes of the server step
    m, n, k = 4096, 3072, 768
    a = torch.randn(m, k, device=dev).to(torch.bfloat16); b = torch.randn(n, k, device=dev).to(torch.bfloat16)
    bias = torch.randn(n, device=dev)
    ms = timeit(lambda: ops.gemm_bf16(a, b, bias=bias, act=ops.ACT_GELU, want_preact=True))
    res['ffn1_gelu_preact'] = {'ms': ms, 'tflops': 2 * m * n * k / ms / 1e9}
    m, n, k = 401408, 256, 64
    a = torch.randn(m, k, device=dev).to(torch.bfloat16); b = torch.randn(k, n, device=dev).to(torch.bfloat16)
    add = torch.randn(m, n, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.gemm_bf16(a, b, b_mn=True, add=add))
    res['conv1x1_dgrad_add_401408x256x64'] = {'ms': ms, 'GBps': (m * k + 2 * m * n) * 2 / ms / 1e6}
    m, n, k = 100352, 512, 128
    a = torch.randn(m, k, device=dev).to(torch.bfloat16); b = torch.randn(n, k, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.gemm_bf16(a, b))
    res['conv1x1_fprop_100352x512x128'] = {'ms': ms, 'GBps': (m * k + m * n) * 2 / ms / 1e6}
    m, n, k = 256, 64, 401408      # wgrad layer1 conv1
    a = torch.randn(k, m, device=dev).to(torch.bfloat16); b = torch.randn(k, n, device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.gemm_bf16(a, b, a_mn=True, b_mn=True, split_k=0, accumulate=True, out=torch.zeros(m, n, device=dev)))
    res['conv1x1_wgrad_256x64x401408'] = {'ms': ms, 'GBps': (m * k + n * k) * 2 / ms / 1e6}
if 'bn' in which:
    from creamfl_b200 import tower_ops as T
    x = torch.randn(128, 56, 56, 256, device=dev).to(torch.bfloat16)
    g = torch.ones(256, device=dev); bta = torch.zeros(256, device=dev); rm = torch.zeros(256, device=dev); rv = torch.ones(256, device=dev)
    sc = T.BNScratch(256, dev)
    ms = timeit(lambda: T.bn_train_fwd(x, g, bta, rm, rv, sc, 1e-5, 0.1))
    nb = x.numel() * 2
    res['bn_train_fwd_128x56x56x256'] = {'ms': ms, 'GBps_algorithmic(3 passes)': 3 * nb / ms / 1e6}
    y, mean, rstd = T.bn_train_fwd(x, g, bta, rm, rv, sc, 1e-5, 0.1)
    dg = torch.zeros(256, device=dev); db = torch.zeros(256, device=dev)
    ms = timeit(lambda: T.bn_train_bwd(x, y, x, g, mean, rstd, sc, dg, db, want_g=True))
    res['bn_train_bwd_128x56x56x256'] = {'ms': ms, 'GBps_algorithmic(8 passes)': 8 * nb / ms / 1e6}
if 'conw' in which:
    for n in (16384, 50000):
        v = unit(torch.randn(n, 256, device=dev)).to(torch.bfloat16)
        g = unit(torch.randn(n, 256, device=dev)).to(torch.bfloat16)
        ms = timeit(lambda: ops.conw_score(v, g), iters=5)
        res[f'conw_score_{n}'] = {'ms': ms, 'tflops': 2.0 * n * n * 256 / ms / 1e9}
if 'infonce' in which:
    for b in (128, 512):
        n = 50000
        bank = unit(torch.randn(n, 256, device=dev)).to(torch.bfloat16)
        q = unit(torch.randn(b, 256, device=dev)).requires_grad_(True)
        lab = torch.randint(0, n, (b,), device=dev)
        ms_f = timeit(lambda: ops.infonce_loss(q, bank, lab))

        def fb():
            q.grad = None
            ops.infonce_loss(q, bank, lab).backward()
        ms_fb = timeit(fb)
        res[f'infonce_B{b}'] = {'fwd_ms': ms_f, 'fwd_bwd_ms': ms_fb, 'fwd_GBps': n * 256 * 2 / ms_f / 1e6}
if 'reduce' in which:
    n, c = 50000, 8
    vs = [unit(torch.randn(n, 256, device=dev)) for _ in range(c)]
    sc = torch.randn(c, n, device=dev)
    ms = timeit(lambda: ops.conw_reduce(vs, sc))
    res['conw_reduce_C8'] = {'ms': ms, 'GBps': (c * n * 256 * 4 + c * n * 4 + n * 256 * 4) / ms / 1e6}
print(json.dumps(res, indent=1))
Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/microbench.json').write_text(json.dumps(res, indent=1))
