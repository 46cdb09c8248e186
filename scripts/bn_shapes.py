"""Graph-launched timing of the BatchNorm entry points on the shapes of the server (ResNet101) and client (ResNet18)
steps at batch 128: forward (statistics ready, as after a fused-statistics convolution) and backward in its three
flavours.  Reports effective GB/s on the bytes each call actually moves.  Development aid."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import tower_ops as T  # noqa: E402

dev = torch.device('cuda:0')
REP, SETS = 12, 4
BF = torch.bfloat16


def timeit(fn):
    for i in range(SETS):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            fn(i % SETS)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / REP * 1e3)
    return sorted(ts)[2]


res = {}
for (p, c) in [(25088, 1024), (25088, 256), (401408, 256), (100352, 512), (401408, 64), (100352, 128), (6272, 2048),
               (6272, 512), (1605632, 64)]:
    shape = (128, p // 128, 1, c)
    xs = [torch.randn(shape, device=dev).to(BF) for _ in range(SETS)]
    rs = [torch.randn(shape, device=dev).to(BF) for _ in range(SETS)]
    dys = [torch.randn(shape, device=dev).to(BF) for _ in range(SETS)]
    gamma, beta = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    sc = T.BNScratch(c, dev)
    nb = p * c * 2

    def fwd(i, res_=False):
        return T.bn_train_fwd(xs[i], gamma, beta, rm, rv, sc, 1e-5, 0.1, res=rs[i] if res_ else None, relu=True,
                              stats_ready=False)
    y, mean, rstd = fwd(0)
    dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    t_f = timeit(lambda i: fwd(i))
    t_fr = timeit(lambda i: fwd(i, True))
    t_bx = timeit(lambda i: T.bn_train_bwd(dys[i], None, xs[i], gamma, mean, rstd, sc, dg, db, beta=beta, relu_from_x=True))
    t_by = timeit(lambda i: T.bn_train_bwd(dys[i], rs[i], xs[i], gamma, mean, rstd, sc, dg, db, want_g=True))
    r = {'fwd_stats+apply_us': round(t_f, 1), 'fwd_GBps(3 passes)': round(3 * nb / t_f / 1e3),
         'fwd_res_us': round(t_fr, 1), 'fwd_res_GBps(4 passes)': round(4 * nb / t_fr / 1e3),
         'bwd_gate_from_x_us': round(t_bx, 1), 'bwd_x_GBps(5 passes)': round(5 * nb / t_bx / 1e3),
         'bwd_mask_from_y_g_us': round(t_by, 1), 'bwd_y_GBps(8 passes)': round(8 * nb / t_by / 1e3)}
    res[f'{p}x{c}'] = r
    print(f'{p}x{c}', r, flush=True)
Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/bn_shapes.json').write_text(json.dumps(res, indent=1))
