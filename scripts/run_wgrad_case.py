"""Drives the weight-gradient GEMM on the two ResNet101 layer-3 1x1 shapes (for an `ncu -k regex:gemm_tc_kernel`
capture).  Development aid."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import ops  # noqa: E402

dev = 'cuda'
for m, n, k in ((1024, 256, 25088), (256, 1024, 25088)):
    a = torch.randn(k, m, device=dev).to(torch.bfloat16)
    b = torch.randn(k, n, device=dev).to(torch.bfloat16)
    out = torch.zeros(m, n, device=dev)
    for _ in range(3):
        ops.gemm_bf16(a, b, a_mn=True, b_mn=True, out=out, split_k=0, accumulate=True)
torch.cuda.synchronize()
print('ok')
