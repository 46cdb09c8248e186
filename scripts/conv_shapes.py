"""Graph-launched timing of the 3x3 implicit-GEMM convolution (fprop / dgrad / wgrad) on the ResNet101 / ResNet18
shapes at batch 128, rotating operand sets.  Development aid -> gpurun_out/conv_shapes.json."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import tower_ops as T  # noqa: E402

dev = torch.device('cuda:0')
REP, SETS = 12, 4
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)


def timeit(fn):
    for i in range(SETS):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            fn(i % SETS)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / REP * 1e3)
    return sorted(ts)[2]


res = {}
for (n, hw, c) in [(128, 14, 256), (128, 28, 128), (128, 7, 512), (128, 56, 64)]:
    xs = [bf(n, hw, hw, c) for _ in range(SETS)]
    dys = [bf(n, hw, hw, c) for _ in range(SETS)]
    w = bf(c, 9 * c) / (9 * c) ** 0.5
    dw = torch.zeros(c, 9 * c, device=dev)
    flops = 2.0 * n * hw * hw * c * c * 9
    r = {}
    for name, fn in (('fprop', lambda i: T.conv_fprop(xs[i], w, 3, 3, 1, 1)),
                     ('dgrad', lambda i: T.conv_dgrad(dys[i], w, xs[i].shape, 3, 3, 1, 1)),
                     ('wgrad', lambda i: T.conv_wgrad(dys[i], xs[i], dw, 3, 3, 1, 1))):
        us = timeit(fn)
        r[name] = {'us': round(us, 1), 'tflops': round(flops / us / 1e6, 1), 'frac_tensor': round(flops / us / 1e6 / 1435.6, 3)}
    res[f'{n}x{hw}x{hw}x{c}'] = r
    print(f'{n}x{hw}x{hw}x{c}', r, flush=True)
Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/conv_shapes.json').write_text(json.dumps(res, indent=1))
