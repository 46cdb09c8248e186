"""Times the BatchNorm statistics / backward kernels on the ResNet101 layer shapes at batch 128 (L2 flushed between
launches).  Development aid for the reduction-grid heuristic; prints one JSON object."""
import json
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import tower_ops as T  # noqa: E402

dev = torch.device('cuda', 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
SHAPES = [(25088, 1024, True), (25088, 256, False), (100352, 512, True), (100352, 128, False), (401408, 256, True),
          (401408, 64, False), (6272, 2048, True), (6272, 512, False)]
out = {'ppt_small': os.environ.get('CFL_BN_PPT_SMALL'), 'ppt_large': os.environ.get('CFL_BN_PPT_LARGE')}
tot = 0.0
for p, c, residual in SHAPES:
    x = torch.randn(p, 1, 1, c, device=dev).to(torch.bfloat16)
    dy = torch.randn(p, 1, 1, c, device=dev).to(torch.bfloat16)
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    sc = T.BNScratch(c, dev)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    res = torch.randn_like(x) if residual else None
    y, mean, rstd = T.bn_train_fwd(x, gamma, beta, rm, rv, sc, 1e-5, 0.1, res=res, relu=True)
    ts_f, ts_b = [], []
    for _ in range(7):
        flush.zero_()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        T.bn_train_fwd(x, gamma, beta, rm, rv, sc, 1e-5, 0.1, res=res, relu=True)
        e1.record()
        if residual:
            T.bn_train_bwd(dy, y, x, gamma, mean, rstd, sc, dg, db, want_g=True)
        else:
            T.bn_train_bwd(dy, None, x, gamma, mean, rstd, sc, dg, db, beta=beta, relu_from_x=True)
        e2.record()
        torch.cuda.synchronize()
        ts_f.append(e0.elapsed_time(e1) * 1e3)
        ts_b.append(e1.elapsed_time(e2) * 1e3)
    f, b = sorted(ts_f)[3], sorted(ts_b)[3]
    out[f'{p}x{c}'] = {'fwd_us': round(f, 1), 'bwd_us': round(b, 1)}
    tot += f + b
out['total_us'] = round(tot, 1)
print(json.dumps(out))
