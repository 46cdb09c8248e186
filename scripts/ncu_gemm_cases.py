"""The GEMM shapes of the server step that sit furthest from their roofline, two launches each (ncu target):
  0/1  BERT FFN1 forward, GELU epilogue + pre-activation output  4096 x 3072 x 768
  2/3  ResNet101 layer-3 1x1 256->1024 forward with BatchNorm statistics   25088 x 1024 x 256
  4/5  layer-3 1x1 dgrad 1024->256 with residual add                25088 x 1024 x 256 (B MN-major)
  6/7  layer-3 1x1 wgrad (split-K, fp32 reds)                      1024 x 256 x 25088
  8/9  3x3 256->256 @14x14 implicit GEMM forward with statistics
ncu --set full --import-source on -k regex:'gemm_tc|conv_tc' python scripts/ncu_gemm_cases.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import ops, tower_ops as T  # noqa: E402

dev = torch.device('cuda:0')
bf = lambda *s: torch.randn(*s, device=dev).to(torch.bfloat16)
a, w, bias = bf(4096, 768), bf(3072, 768), torch.randn(3072, device=dev)
x3, w3 = bf(128, 14, 14, 256), bf(1024, 256)
dy, add = bf(128, 14, 14, 256), bf(128, 14, 14, 1024)
x33, w33 = bf(128, 14, 14, 256), bf(256, 9 * 256) / 48
gw = torch.zeros(1024, 256, device=dev)
sums = torch.zeros(2 * 1024, dtype=torch.float64, device=dev)
sums2 = torch.zeros(2 * 256, dtype=torch.float64, device=dev)
for _ in range(2):
    ops.gemm_bf16(a, w, bias=bias, act=ops.ACT_GELU, want_preact=True)
for _ in range(2):
    T.conv_fprop(x3, w3, 1, 1, 1, 0, bn_sums=sums)
for _ in range(2):
    T.conv_dgrad(dy, w3.t().contiguous(), (128, 14, 14, 1024), 1, 1, 1, 0, add=add)
for _ in range(2):
    T.conv_wgrad(add, x3, gw, 1, 1, 1, 0)
for _ in range(2):
    T.conv_fprop(x33, w33, 3, 3, 1, 1, bn_sums=sums2)
torch.cuda.synchronize()
print('ok')
