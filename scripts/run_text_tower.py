"""Microbenchmark of the GRU text tower (multimodal client, B = 128 captions of 5..30 words, D = 256): the whole
tower forward+backward through the public module, the two recurrent kernels alone, and - as a library yardstick only -
torch's nn.Embedding + packed cuDNN nn.GRU on the same shapes.  Development aid; prints one JSON object."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from creamfl_b200 import tower_ops as T  # noqa: E402
from creamfl_b200.text_towers import TextModel  # noqa: E402

dev = torch.device('cuda', 0)
B, L, D, V = 128, 30, 256, 11755
H = D // 2
g = torch.Generator().manual_seed(0)
lengths = torch.sort(torch.randint(5, L + 1, (B,), generator=g), descending=True).values
lengths[0] = L
x = (torch.randint(4, V, (B, L), generator=g) * (torch.arange(L)[None] < lengths[:, None])).to(dev)
coef = torch.randn(B, D, generator=g).to(dev)


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


model = TextModel(V, 300, D).to(dev).train()
st = model.store()


def tower_step():
    st.zero_grad()
    (model(x, lengths) * coef).sum().backward()


out = {'shape': {'B': B, 'L': L, 'D': D, 'tokens': int(lengths.sum())}}
out['tower_fwd_bwd_eager_ms'] = timeit(tower_step)
with torch.no_grad():
    out['tower_fwd_eval_ms'] = timeit(lambda: model(x, lengths))
# CUDA graph of the training step (how the client steps run in the engine)
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        tower_step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    tower_step()
out['tower_fwd_bwd_graph_ms'] = timeit(graph.replay)

# the recurrent kernels alone
tw = model.txt_enc
len32 = tw.lengths32(lengths, dev)
xproj = torch.randn(B * L, 6 * H, device=dev)
for rev, tag in ((1, 'rev1'), (0, 'full')):
    hseq, hlast, gates = T.gru_fwd(xproj, tw._whh, tw._bhh, len32, B, L, H, rev_steps=rev)
    dlast = torch.randn(B, 2 * H, device=dev)
    out[f'gru_fwd_{tag}_ms'] = timeit(lambda: T.gru_fwd(xproj, tw._whh, tw._bhh, len32, B, L, H, rev_steps=rev))
    out[f'gru_bwd_{tag}_ms'] = timeit(lambda: T.gru_bwd(gates, hseq, tw._whh, len32, None, dlast, B, L, H, rev_steps=rev))

# library yardstick: torch embedding + packed cuDNN GRU, fwd + bwd of the last-valid-step output
emb = torch.nn.Embedding(V, 300).to(dev)
gru = torch.nn.GRU(300, H, bidirectional=True, batch_first=True).to(dev)
idx = (lengths - 1).to(dev).view(-1, 1, 1).expand(-1, 1, D)


def cudnn_step():
    emb.zero_grad(set_to_none=True)
    gru.zero_grad(set_to_none=True)
    packed = torch.nn.utils.rnn.pack_padded_sequence(emb(x), lengths, batch_first=True)
    o, _ = gru(packed)
    padded, _ = torch.nn.utils.rnn.pad_packed_sequence(o, batch_first=True, total_length=L)
    (padded.gather(1, idx).squeeze(1) * coef).sum().backward()


out['torch_cudnn_embed_gru_fwd_bwd_ms'] = timeit(cudnn_step)
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}))
