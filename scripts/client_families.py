"""Where the time of a multimodal client's contrast step / generate call goes (ResNet18 + GRU at batch 128, InfoNCE
against N_pub = 50000): one eager single-stream step with CUDA events around every C-ABI call (calltimer.py), per
family and per shape, next to the graph-replayed time of the same step.  Development aid -> gpurun_out/client_families.json."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from creamfl_b200 import engine  # noqa: E402
from creamfl_b200.calltimer import CallTimer  # noqa: E402

dev = torch.device('cuda:0')
B, D = 128, 256
torch.manual_seed(0)
pub = {k: v[0].to(dev) for k, v in bench.make_public(1, B, 7, pin=False).items()}
g_img, g_txt = bench.make_banks(dev, 5)


def table(t, total_key='sum_ms'):
    fam = t.families()
    total = sum(f['ms'] for f in fam.values())
    out = {total_key: round(total, 3), 'families': {}}
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]['ms']):
        row = {'ms': round(f['ms'], 3), 'share': round(f['ms'] / total, 3), 'calls': f['calls']}
        if f['flops']:
            row['tflops'] = round(f['flops'] / f['ms'] / 1e9, 1)
        if f['bytes']:
            row['gbps'] = round(f['bytes'] / f['ms'] / 1e6, 1)
        shapes = {s: {'ms': round(d['ms'], 3), 'calls': d['calls']} for s, d in
                  sorted(t.shapes(k).items(), key=lambda kv: -kv[1]['ms'])[:8] if s}
        if shapes:
            row['shapes'] = shapes
        out['families'][k] = row
    names = {}
    for name, (f, _fl, _by, _tag), a, b in t.calls:
        if f == 'other':
            d = names.setdefault(name, [0.0, 0])
            d[0] += t._ms(a, b)
            d[1] += 1
    out['other_by_entry'] = {k: [round(v[0], 3), v[1]] for k, v in sorted(names.items(), key=lambda kv: -kv[1][0])}
    return out


def replay_ms(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


res = {}
for graphs in (False, True):
    torch.manual_seed(1)
    cl = engine.MMClient(D, device=dev, use_graphs=graphs)
    cl.begin_round()
    contrast = lambda: cl.contrast_step(pub['images'], pub['caps'], pub['cap_lens'], pub['d_idx'], g_img, g_txt)
    generate = lambda: cl.generate(pub['images'], pub['caps'], pub['cap_lens'])
    key = 'graph' if graphs else 'eager'
    res[key] = {'contrast_ms': round(replay_ms(contrast), 3), 'generate_ms': round(replay_ms(generate), 3)}
    if not graphs:
        cl.overlap_old_model = False
        cl.model.overlap_towers = False
        for name, fn in (('contrast', contrast), ('generate', generate)):
            fn()
            torch.cuda.synchronize()
            with CallTimer() as t:
                t.stall(60.0)
                fn()
                res[name] = table(t)
    print(key, res[key], flush=True)
for name in ('contrast', 'generate'):
    print(name, json.dumps(res[name], indent=1), flush=True)
Path(ROOT / 'gpurun_out').mkdir(exist_ok=True)
(ROOT / 'gpurun_out' / 'client_families.json').write_text(json.dumps(res, indent=1))
