import sys, torch
sys.path.insert(0, '.')
from creamfl_b200 import tower_ops as T
g = torch.Generator().manual_seed(0)
x = torch.randn(128, 14, 14, 256, generator=g).to(torch.bfloat16).cuda()
w = (torch.randn(256, 9 * 256, generator=g) / 48).to(torch.bfloat16).cuda()
for _ in range(3):
    y = T.conv_fprop(x, w, 3, 3, 1, 1)
torch.cuda.synchronize()
print(float(y.float().abs().mean()))
