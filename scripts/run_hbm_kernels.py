"""Drives the HBM-bound kernels of the path once each at their largest shapes (for `ncu -k regex:...` captures):
BatchNorm forward / backward on a layer-1 map, the fused optimizer over 155 M parameters, the con_w weighted reduce at
8 clients, the masked sequence pooling.  Development aid."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import ops, tower_ops as T  # noqa: E402
from creamfl_b200.optim import FusedOptimizer  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(0)
# BatchNorm on [128, 56, 56, 256] (205 MB per tensor: layer1 bn3 of ResNet101 at batch 128)
n, h, w, c = 128, 56, 56, 256
x = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
res = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
dy = torch.randn(n, h, w, c, device=dev).to(torch.bfloat16)
gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
sc = T.BNScratch(c, dev)
dgamma, dbeta = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
for _ in range(2):
    y, mean, rstd = T.bn_train_fwd(x, gamma, beta, rm, rv, sc, 1e-5, 0.1, res=res, relu=True)
    dx, gout = T.bn_train_bwd(dy, y, x, gamma, mean, rstd, sc, dgamma, dbeta, want_g=True)
# fused optimizer: 155 M parameters in 600 tensors
params = [torch.nn.Parameter(torch.randn(512, 505, device=dev)) for _ in range(600)]
for p in params:
    p.grad = torch.randn_like(p) * 1e-3
opt = FusedOptimizer(params, lr=2e-4, max_norm=2.0, mode='adamp')
for _ in range(2):
    opt.step()
# con_w weighted reduce at C = 8 clients (461 MB)
vecs = [torch.nn.functional.normalize(torch.randn(50000, 256, device=dev), dim=1) for _ in range(8)]
scores = torch.randn(8, 50000, device=dev)
for _ in range(2):
    out = ops.conw_reduce(vecs, scores)
torch.cuda.synchronize()
print('ok', float(out.abs().sum()))
