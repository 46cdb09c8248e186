"""Split-K sweep of the weight-gradient GEMM (both operands MN-major, fp32 atomic accumulation) on the shapes the
server step issues most often; calibrates the split planner of gemm_tc.cu (waves of units x (k-blocks + epilogue)).
Development aid; prints one JSON line per shape."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import ops  # noqa: E402

dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CASES = [  # (M = Cout, N = Cin, K = pixels / tokens, splits)
    (1024, 256, 25088, [0, 6, 9, 12, 18, 19, 24, 36, 37, 54]),
    (256, 1024, 25088, [0, 9, 18, 19, 36, 37]),
    (2048, 512, 6272, [0, 2, 3, 4, 5, 8, 9]),
    (512, 128, 100352, [0, 18, 37, 72, 74, 148]),
    (768, 768, 4096, [0, 2, 4, 8, 16]),
    (3072, 768, 4096, [0, 1, 2, 3, 4]),
    (2304, 768, 4096, [0, 1, 2, 3, 5]),
]
for m, n, k, splits in CASES:
    a = torch.randn(k, m, device=dev).to(torch.bfloat16)
    b = torch.randn(k, n, device=dev).to(torch.bfloat16)
    out = torch.zeros(m, n, device=dev)
    res = {}
    for s in splits:
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm_bf16(a, b, a_mn=True, b_mn=True, out=out, split_k=s, accumulate=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res[s] = round(sorted(ts)[len(ts) // 2], 1)
    print(json.dumps({'M': m, 'N': n, 'K': k, 'us_by_split': res, 'gflop': round(2 * m * n * k / 1e9, 2)}))
