import sys, torch
sys.path.insert(0, '.')
from creamfl_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
g = torch.Generator().manual_seed(0)
u = lambda x: x / x.norm(dim=-1, keepdim=True)
v = u(torch.randn(n, 256, generator=g)).cuda().to(torch.bfloat16)
w = u(torch.randn(n, 256, generator=g)).cuda().to(torch.bfloat16)
for _ in range(3):
    s = ops.conw_score(v, w)
torch.cuda.synchronize()
print(s[:4])
