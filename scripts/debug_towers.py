import sys, torch
sys.path.insert(0, '.')
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
from creamfl_b200 import towers
from oracle import torch_towers as RT

def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()

arch, batch = sys.argv[1], int(sys.argv[2])
ref = RT.RefEncoderImage(arch, 256); RT.fill_deterministic(ref, 1)
with torch.no_grad():
    for name, p in ref.named_parameters():
        if name.endswith('bn3.weight') or (arch == 'resnet18' and name.endswith('bn2.weight')):
            p.mul_(0.2)
ref = ref.cuda().train()
mine = towers.ImageModel({'embed_dim': 256, 'cnn_type': arch}); mine.img_enc.load_state_dict(ref.state_dict()); mine = mine.cuda().train()
g = torch.Generator().manual_seed(2)
images = torch.randn(batch, 3, 224, 224, generator=g).cuda()
cot = torch.randn(batch, 256, generator=g).cuda()
import copy
ref2 = copy.deepcopy(ref)
e_ref = ref(images)['embedding']; (e_ref * cot).sum().backward()
with torch.autocast('cuda', dtype=torch.bfloat16):
    e2 = ref2(images)['embedding']
(e2.float() * cot).sum().backward()
mine.zero_grad(); e = mine(images); (e * cot).sum().backward()
torch.cuda.synchronize()
print('emb cos mine/ref', min(cos(e[i], e_ref[i]) for i in range(batch)), 'autocast/ref', min(cos(e2[i], e_ref[i]) for i in range(batch)))
rp, r2p, mp = dict(ref.named_parameters()), dict(ref2.named_parameters()), dict(mine.img_enc.named_parameters())
for n in rp:
    if n.endswith('weight') and ('conv' in n or 'downsample.0' in n or n.startswith('fc') or 'pie' in n) or n.endswith('bn1.bias'):
        print(f'{n:45s} mine/ref {cos(mp[n].grad, rp[n].grad):.4f} ratio {(mp[n].grad.norm()/rp[n].grad.norm()).item():.3f} | autocast/ref {cos(r2p[n].grad, rp[n].grad):.4f} | mine/autocast {cos(mp[n].grad, r2p[n].grad):.4f}')
