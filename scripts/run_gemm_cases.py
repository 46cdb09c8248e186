import sys, torch
sys.path.insert(0, '.')
from creamfl_b200 import ops
dev = 'cuda'
m, n, k = 100352, 512, 128
a = torch.randn(m, k, device=dev).to(torch.bfloat16); b = torch.randn(n, k, device=dev).to(torch.bfloat16)
m2, n2, k2 = 401408, 256, 64
a2 = torch.randn(m2, k2, device=dev).to(torch.bfloat16); b2 = torch.randn(k2, n2, device=dev).to(torch.bfloat16)
add = torch.randn(m2, n2, device=dev).to(torch.bfloat16)
for _ in range(3):
    ops.gemm_bf16(a, b)
    ops.gemm_bf16(a2, b2, b_mn=True, add=add)
torch.cuda.synchronize()
