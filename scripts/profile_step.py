"""In-situ kernel timing of the hot-path phases with torch.profiler (CUPTI): wall time vs summed kernel time,
top kernels.  Development aid."""
import sys, json, collections, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from creamfl_b200 import engine, ops
from torch.profiler import profile, ProfilerActivity

dev = torch.device('cuda', 0)
S, B = 1, 128
server = engine.ServerEngine(256, 'resnet101', device=dev)
client = engine.MMClient(256, device=dev)
host = bench.make_host_batches(S, B, 1)
cur = {k: (v.to(dev) if k != 'cap_lens' else v) for k, v in host.items()}
g_img, g_txt, c_img, c_txt = bench.make_banks(dev, 99)
g_img16, g_txt16 = ops.to_bf16(g_img), ops.to_bf16(g_txt)
tok = {'input_ids': cur['ids'][0], 'attention_mask': cur['mask'][0]}
client.begin_round()
phases = {
    'server_train': lambda: server.train_step(cur['images'][0], tok),
    'server_extract': lambda: server.extract(cur['images'][0], tok),
    'client_contrast': lambda: client.contrast_step(cur['images'][0], cur['caps'][0], cur['cap_lens'], cur['d_idx'][0], g_img, g_txt, g_img16, g_txt16),
    'client_private': lambda: client.private_step(cur['priv_images'], cur['priv_caps'], cur['cap_lens']),
    'client_generate': lambda: client.generate(cur['images'][0], cur['caps'][0], cur['cap_lens']),
    'conw': lambda: (engine.exchange_and_aggregate(c_img, g_txt16), engine.exchange_and_aggregate(c_txt, g_img16)),
    'server_distill': lambda: server.distill_step(cur['images'][0], tok, cur['d_idx'][0], c_img, c_txt),
}
which = sys.argv[1:] or list(phases)
report = {}
for name in which:
    fn = phases[name]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 3 * 1e3
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    total = 0.0
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = ev.name.split('(')[0][:60]
            agg[k][0] += 1
            agg[k][1] += ev.device_time_total / 1e3 if hasattr(ev, 'device_time_total') else ev.cuda_time_total / 1e3
    total = sum(v for _, v in agg.values())
    n = sum(c for c, _ in agg.values())
    print(f'== {name}: wall {wall:.2f} ms | kernels {n} | summed kernel time {total:.2f} ms')
    top = sorted(agg.items(), key=lambda x: -x[1][1])[:24]
    for k, (c, v) in top:
        print(f'   {v:8.3f} ms {100 * v / total:5.1f}% {c:5d}  {k}')
    report[name] = {'wall_ms': wall, 'kernels': n, 'kernel_ms': total, 'top': [(k, c, v) for k, (c, v) in top]}
Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/profile_step.json').write_text(json.dumps(report, indent=1))
