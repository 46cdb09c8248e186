"""Per-shape breakdown (CUDA events around every C-ABI call, creamfl_b200/calltimer.py) of one eager server train step
and one eager multimodal-client contrast step at batch 128.  Writes gpurun_out/family_shapes.json.  Development aid."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import engine  # noqa: E402
from creamfl_b200.calltimer import CallTimer  # noqa: E402
import bench  # noqa: E402

dev = torch.device('cuda:0')
B = 128
pub = {k: v.to(dev) for k, v in bench.make_public(1, B, 1, pin=False).items()}
out = {}


def table(t, fams):
    res = {'families': {k: {kk: round(vv, 3) if isinstance(vv, float) else vv for kk, vv in v.items()}
                        for k, v in sorted(t.families().items(), key=lambda kv: -kv[1]['ms'])}}
    for f in fams:
        res[f] = {k: {'ms': round(v['ms'], 4), 'calls': v['calls'], 'us_per_call': round(1e3 * v['ms'] / v['calls'], 1),
                      'tflops': round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 1), 'GBps': round(v['bytes'] / max(v['ms'], 1e-9) / 1e6, 1)}
                  for k, v in sorted(t.shapes(f).items(), key=lambda kv: -kv[1]['ms'])}
    return res


FAMS = ['bn_bwd', 'bn_fwd', 'conv_tc_fwd', 'conv_tc_dgrad', 'conv_tc_wgrad', 'gemm_tc_fwd', 'gemm_tc_dgrad', 'gemm_tc_wgrad',
        'gemm_tc_im2col_fwd', 'gemm_tc_im2col_dgrad', 'gemm_tc_im2col_wgrad']
torch.manual_seed(0)
server = engine.ServerEngine(256, 'resnet101', device=dev)
server.model.overlap_towers = False
txt = {'ids': pub['ids'][0], 'mask': pub['mask'][0]}
for _ in range(3):
    server._train_step(pub['images'][0], txt)
torch.cuda.synchronize()
with CallTimer() as t:
    t.stall(150.0)
    server._train_step(pub['images'][0], txt)
    out['server_train_step'] = table(t, FAMS)
del server
torch.cuda.empty_cache()

client = engine.MMClient(256, device=dev)
client.overlap_old_model = False
client.model.overlap_towers = False
g_img, g_txt = bench.make_banks(dev, 3)
client.begin_round()
args = (pub['images'][0], pub['caps'][0], pub['cap_lens'][0], pub['d_idx'][0], g_img, g_txt)
for _ in range(3):
    client.contrast_step(*args)
torch.cuda.synchronize()
with CallTimer() as t:
    t.stall(150.0)
    client.contrast_step(*args)
    out['client_contrast_step'] = table(t, FAMS)
Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/family_shapes.json').write_text(json.dumps(out, indent=1))
for phase, res in out.items():
    print('==', phase)
    for k, v in res['families'].items():
        print(f"  {k:14s} {v['ms']:8.3f} ms  calls {v['calls']}")
