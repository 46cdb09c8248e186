"""Timing of the fused stem kernels against the im2col + GEMM path they replace (batch 128, 224x224), graph-launched
back to back on rotating inputs.  Development aid."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import ops, tower_ops as T  # noqa: E402

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
REP, SETS = 8, 4
imgs = [torch.randn(B, 3, 224, 224, device=dev) for _ in range(SETS)]
w16 = torch.zeros(64, 152, dtype=torch.bfloat16, device=dev)
w16[:, :147] = torch.randn(64, 147, device=dev).to(torch.bfloat16) * 0.1
dy = [torch.randn(B, 112, 112, 64, device=dev).to(torch.bfloat16) for _ in range(SETS)]
dw = torch.zeros(64, 147, device=dev)
cols = [T.im2col_images(imgs[i], 7, 7, 2, 3, 152) for i in range(2)]


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
    return round(sorted(ts)[2], 1)


res = {
    'stem_fprop_fused_us': timeit(lambda i: T.stem_fprop(imgs[i], w16)),
    'stem_wgrad_fused_us': timeit(lambda i: T.stem_wgrad(imgs[i], dy[i], dw)),
    'im2col_us': timeit(lambda i: T.im2col_images(imgs[i], 7, 7, 2, 3, 152)),
    'gemm_fwd_us': timeit(lambda i: ops.gemm_bf16(cols[i % 2], w16)),
    'gemm_wgrad_us': timeit(lambda i: ops.gemm_bf16(dy[i].view(-1, 64), cols[i % 2], a_mn=True, b_mn=True, out=dw,
                                                    split_k=0, accumulate=True, n_cols=147)),
}
print(json.dumps(res))
Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/stem_bench.json').write_text(json.dumps(res))
