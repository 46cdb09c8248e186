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
