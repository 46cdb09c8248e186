"""Per-shape timing of the tcgen05 GEMM on the shapes the server step runs (BERT linears at T = 4096, ResNet101 1x1
convolutions at batch 128), free of host launch latency: each shape is captured as a CUDA graph of REP back-to-back
launches on rotating operand sets (so consecutive launches do not hit the same L2 lines) and timed with CUDA events.
Writes gpurun_out/gemm_shapes.json.  Development aid; bench.py is the contract benchmark."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import ops  # noqa: E402

dev = torch.device('cuda:0')
REP, SETS = 24, 6
PEAK = 1435.6


def bench(name, make, call, flops, bytes_):
    sets = [make() for _ in range(SETS)]
    for s in sets:
        call(*s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(REP):
            call(*sets[i % SETS])
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / REP)
    us = sorted(ts)[len(ts) // 2] * 1e3
    r = {'us': round(us, 2), 'tflops': round(flops / us / 1e6, 1), 'GBps': round(bytes_ / us / 1e3, 1),
         'frac_tensor': round(flops / us / 1e6 / PEAK, 3), 'frac_hbm': round(bytes_ / us / 1e3 / 6447.5, 3)}
    print(name, r, flush=True)
    return r


def bf(*shape):
    return torch.randn(*shape, device=dev).to(torch.bfloat16)


res = {}
T = 4096
# ---- BERT forward / dgrad / wgrad (towers._BertFn)
for name, (m, n, k), kw in [
        ('bert_qkv_fwd', (T, 2304, 768), dict(bias=True)),
        ('bert_attnout_fwd_add', (T, 768, 768), dict(bias=True, add=True)),
        ('bert_ffn1_fwd_gelu_preact', (T, 3072, 768), dict(bias=True, act=ops.ACT_GELU, want_preact=True)),
        ('bert_ffn2_fwd_add', (T, 768, 3072), dict(bias=True, add=True)),
        ('plain_8192', (8192, 8192, 8192), {}),
        ('res_l3_1x1_256to1024_fwd', (25088, 1024, 256), {}),
        ('res_l3_1x1_1024to256_fwd', (25088, 256, 1024), {}),
        ('res_l2_1x1_128to512_fwd', (100352, 512, 128), {}),
        ('res_l1_1x1_64to256_fwd', (401408, 256, 64), {}),
        ('res_l4_1x1_512to2048_fwd', (6272, 2048, 512), {})]:
    def make(m=m, n=n, k=k, kw=kw):
        return (bf(m, k), bf(n, k), torch.randn(n, device=dev) if kw.get('bias') else None,
                bf(m, n) if kw.get('add') else None)

    def call(a, b, bias, add, kw=kw):
        return ops.gemm_bf16(a, b, bias=bias, add=add, act=kw.get('act', 0), want_preact=kw.get('want_preact', False))
    byt = 2 * (m * k + n * k + m * n * (1 + int(bool(kw.get('add'))) + int(bool(kw.get('want_preact')))))
    res[name] = dict(bench(name, make, call, 2.0 * m * n * k, byt), M=m, N=n, K=k)

for name, (m, n, k) in [('bert_ffn2_dgrad_dgelu', (T, 3072, 768)), ('bert_qkv_dgrad_add', (T, 768, 2304)),
                        ('res_l3_1x1_dgrad_256', (25088, 256, 1024)), ('res_l3_1x1_dgrad_1024', (25088, 1024, 256))]:
    gelu = 'dgelu' in name

    def make(m=m, n=n, k=k, gelu=gelu):
        return bf(m, k), bf(k, n), (bf(m, n) if gelu else None), (None if gelu else bf(m, n))

    def call(a, b, aux, add, gelu=gelu):
        if gelu:
            return ops.gemm_bf16(a, b, b_mn=True, act=ops.ACT_DGELU, aux=aux)
        return ops.gemm_bf16(a, b, b_mn=True, add=add)
    res[name] = dict(bench(name, make, call, 2.0 * m * n * k, 2 * (m * k + n * k + 2 * m * n)), M=m, N=n, K=k)

for name, (m, n, k) in [('bert_ffn1_wgrad', (3072, 768, T)), ('bert_ffn2_wgrad', (768, 3072, T)),
                        ('bert_qkv_wgrad', (2304, 768, T)), ('res_l3_wgrad_1024x256', (1024, 256, 25088)),
                        ('res_l3_wgrad_256x1024', (256, 1024, 25088)), ('res_l1_wgrad_256x64', (256, 64, 401408))]:
    def make(m=m, n=n, k=k):
        return bf(k, m), bf(k, n), torch.zeros(m, n, device=dev)

    def call(a, b, out):
        return ops.gemm_bf16(a, b, a_mn=True, b_mn=True, split_k=0, accumulate=True, out=out)
    res[name] = dict(bench(name, make, call, 2.0 * m * n * k, 2 * (m * k + n * k) + 4 * m * n), M=m, N=n, K=k)

Path('gpurun_out').mkdir(exist_ok=True)
Path('gpurun_out/gemm_shapes.json').write_text(json.dumps(res, indent=1))
