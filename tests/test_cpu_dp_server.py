"""The data-parallel server of the N-GPU mini-round on the CPU (gloo, world size 2, kernels emulated by
tests/kernel_emulation.py): a train step and a distillation step with the public batches sharded over two ranks must
equal ONE process stepping on the mean of the two ranks' gradients - the arithmetic of the reference's single-process
server seeing both batches with a mean-reduced loss (retrieval_trainer.py:192-214, MMFL.py:346-391).  The 2-GPU NCCL
version of this test (tests/test_gpu_graph_parity.py) needs two GPUs; this one runs everywhere."""
import functools
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

N_PUB, DIM = 24, 32


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(seed, B=3, L=8):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, 224, 224, generator=g)
    ids = torch.randint(1000, 30522, (B, L), generator=g)
    ids[:, 0] = 101
    tok = {'input_ids': ids, 'attention_mask': torch.ones(B, L, dtype=torch.long)}
    d_idx = torch.randint(0, N_PUB, (B,), generator=g)
    return images, tok, d_idx


def _banks():
    g = torch.Generator().manual_seed(99)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    return unit(torch.randn(N_PUB, DIM, generator=g)), unit(torch.randn(N_PUB, DIM, generator=g))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import kernel_emulation as KE
    from creamfl_b200 import engine, towers
    patch = pytest.MonkeyPatch()
    KE.install_engine(patch, exact=True)
    patch.setattr(towers, 'BertEncoder', functools.partial(towers.BertEncoder, layers=2))
    agg_img, agg_txt = _banks()

    def make(dp):
        torch.manual_seed(3)
        return engine.ServerEngine(DIM, 'resnet18', data_parallel=dp, use_graphs=False, bert_dropout=0.0)

    def flat_grads(s):
        return [s.model.store().grad] + [p.grad for p in s.criterion.parameters()]

    def single():
        """one process: train on the mean gradient of both ranks' batches, then distill on the mean gradient"""
        s = make(False)
        for phase in ('train', 'distill'):
            acc = None
            for r in range(world):
                images, tok, d_idx = _inputs(50 + r)
                txt = {'ids': tok['input_ids'], 'mask': tok['attention_mask']}
                if phase == 'train':
                    s._train_fwd_bwd(images, txt)
                else:
                    s._distill_fwd_bwd(images, txt, None, d_idx, agg_img, agg_txt, 2, 1)
                got = [g.clone() for g in flat_grads(s)]
                acc = got if acc is None else [a + g for a, g in zip(acc, got)]
            for g, a in zip(flat_grads(s), acc):
                g.copy_(a / world)
            s.optimizer.step()
        return s

    ref = single()
    assert not ref.data_parallel
    s = make(True)
    assert s.data_parallel
    images, tok, d_idx = _inputs(50 + rank)
    s.train_step(images, tok)
    s.distill_step(images, tok, d_idx, agg_img, agg_txt, img_terms=2, txt_terms=1)
    flat, want = s.model.store().flat, ref.model.store().flat
    moved = float((want - make(False).model.store().flat).abs().max())
    crit = max(float((a - b).abs().max()) for a, b in zip(s.criterion.parameters(), ref.criterion.parameters()))
    other = flat.clone()
    dist.broadcast(other, 0)
    q.put((rank, float((flat - want).abs().max()), crit, float((flat - other).abs().max()), moved))
    dist.barrier()
    dist.destroy_process_group()
    patch.undo()


def test_data_parallel_server_equals_single_process_mean_gradient_step():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, err, crit_err, ranks_differ, moved in results:
        assert moved > 1e-4, moved                     # two AdamP steps at lr 2e-4 really moved the parameters
        assert err <= 1e-6, (rank, err)                # fp32 sum of two gradients in either order, same fp64 optimizer
        assert crit_err <= 1e-6, (rank, crit_err)
        assert ranks_differ == 0.0, (rank, ranks_differ)
