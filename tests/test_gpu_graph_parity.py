"""Parity of the configuration bench.py actually times: CUDA-graph replay of the step functions, the learning-rate
schedule reaching captured kernels, and the NCCL data-parallel server.

The kernels accumulate weight gradients with fp32 atomics (split-K `red.global.add`), so two EAGER runs of the same
step from the same state already differ in the last bits.  "Graph replay == eager" is therefore asserted against
that measured run-to-run noise (||graph - eager||_2 <= 3 * max ||eager' - eager||_2 over further eager runs, per
tensor group), on the flat parameter buffer, BatchNorm running statistics / counters, criterion
parameters, AdamP moments and the step's losses."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope='module')
def engine():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from creamfl_b200 import engine as e
    return e


def _server_inputs(B, L, seed):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, 224, 224, generator=g).cuda()
    lens = torch.sort(torch.randint(8, L + 1, (B,), generator=g), descending=True).values
    lens[0] = L
    mask = (torch.arange(L)[None] < lens[:, None]).long()
    ids = torch.randint(1000, 30522, (B, L), generator=g)
    ids[:, 0] = 101
    return images, {'input_ids': (ids * mask).cuda(), 'attention_mask': mask.cuda()}


def _groups(eng):
    st = eng.model.store()
    out = {'params': st.flat.clone(), 'shadow': st.shadow.float().clone(),
           'bn': torch.cat([b.detach().float().flatten() for b in eng.model.buffers()]),
           'moments': torch.cat([t.flatten() for t in eng.optimizer._keep])}
    if eng.criterion is not None:
        out['criterion'] = torch.cat([p.detach().flatten() for p in eng.criterion.parameters()])
    return out


def _l2diff(a, b):
    return {k: float((a[k].double() - b[k].double()).norm()) for k in a}


def _assert_within_noise(test, base, noise_runs, what):
    """||test - base||_2 <= 3 * max_i ||noise_i - base||_2 + floor, per tensor group.  L2 norms over 10^5..10^8
    elements concentrate, so a semantically identical run sits at ratio ~1; AdamP's first steps move every element by
    ~lr whatever the gradient magnitude, so the MAX difference is dominated by a few sign flips of near-zero gradients
    and is not a usable statistic."""
    d = _l2diff(test, base)
    noise = {k: max(_l2diff(n, base)[k] for n in noise_runs) for k in d}
    floor = {k: 1e-6 * float(base[k].double().norm()) for k in d}             # fp32 rounding of the group itself
    bad = {k: (d[k], noise[k]) for k in d if d[k] > 3 * noise[k] + floor[k]}
    assert not bad, (what, bad)


def _run_server(engine, graphs, steps, batches, sched=False, dropout=0.1):
    torch.manual_seed(7)                      # identical random init and dropout seed for every run
    srv = engine.ServerEngine(256, 'resnet101', lr=2e-4, use_graphs=graphs, bert_dropout=dropout)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(srv.optimizer, T_max=3) if sched else None
    losses, srv.moved = [], []
    for s in range(steps):
        images, tok = batches[s % len(batches)]
        before = srv.model.store().flat.clone()
        losses.append(srv.train_step(images, tok).item())
        srv.moved.append(float((srv.model.store().flat - before).abs().mean()))     # mean |update| of this step
        if scheduler is not None:
            scheduler.step()
    torch.cuda.synchronize()
    return srv, _groups(srv), losses


def test_server_graph_replay_equals_eager_over_three_steps(engine):
    batches = [_server_inputs(16, 32, 100 + s) for s in range(2)]
    _, eager, l_eager = _run_server(engine, False, 3, batches)
    noise = [_run_server(engine, False, 3, batches) for _ in range(3)]
    srv, graph, l_graph = _run_server(engine, True, 3, batches)
    assert int(srv.model.img_enc.cnn.bn1.num_batches_tracked) == 3          # the capture's warm-up was rolled back
    assert float(srv.optimizer._state[0]) == 3.0                            # AdamP step counter
    assert int(srv.model.txt_enc.dropout_state(srv.device).rng[1]) == 3     # dropout RNG step: one tick per step
    _assert_within_noise(graph, eager, [n[1] for n in noise], 'server graph vs eager')
    # step 1 runs on identical parameters: its loss (forward only, fp64 BatchNorm statistics) agrees to fp32 rounding
    assert abs(l_graph[0] - l_eager[0]) <= 1e-5 * abs(l_eager[0])
    # later losses sit on parameters that already carry the atomics noise of the earlier updates
    for a, b, *cs in zip(l_graph[1:], l_eager[1:], *[n[2][1:] for n in noise]):
        assert abs(a - b) <= 4 * max(abs(c - b) for c in cs) + 1e-2 * abs(b)


def test_lr_schedule_reaches_captured_optimizer(engine):
    """after lr_scheduler.step() a replayed graph must use the new learning rate.  AdamP's first updates
    have |update| ~ lr per element whatever the gradient, so the mean |update| of a step is a sharp read-out of the
    learning rate the kernels used: CosineAnnealingLR(T_max=3) runs the steps at 1, 0.75 and 0.25 of the base rate."""
    batches = [_server_inputs(8, 16, 200)]
    s_eager, eager, _ = _run_server(engine, False, 3, batches, sched=True, dropout=0.0)
    noise = [_run_server(engine, False, 3, batches, sched=True, dropout=0.0)[1] for _ in range(2)]
    s_graph, graph, _ = _run_server(engine, True, 3, batches, sched=True, dropout=0.0)
    _assert_within_noise(graph, eager, noise, 'cosine schedule under graphs')
    s_const, _, _ = _run_server(engine, True, 3, batches, sched=False, dropout=0.0)
    for k in range(3):
        assert abs(s_graph.moved[k] - s_eager.moved[k]) <= 0.03 * s_eager.moved[k], (k, s_graph.moved, s_eager.moved)
    # first Adam step: |update| = lr on every element with a gradient (the dead pooler and the embedding rows of absent
    # tokens - 15 % of the parameters - do not move)
    assert 0.6 * 2e-4 < s_graph.moved[0] < 1.01 * 2e-4
    assert s_graph.moved[2] / s_graph.moved[0] < 0.30                       # 0.25 x (|m/sqrt(v)| <= 1)
    assert s_const.moved[2] / s_const.moved[0] > 0.45                       # a graph stuck at the base rate would sit here


def _client_inputs(B, L, n_pub, seed):
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(B, 3, 224, 224, generator=g).cuda()
    lens = torch.sort(torch.randint(5, L + 1, (B,), generator=g), descending=True).values
    lens[0] = L
    caps = (torch.randint(4, 11755, (B, L), generator=g) * (torch.arange(L)[None] < lens[:, None])).cuda()
    d_idx = torch.randperm(n_pub, generator=g)[:B].cuda()
    return images, caps, lens, d_idx


def _run_client(engine, graphs, banks, steps=3):
    torch.manual_seed(9)
    cl = engine.MMClient(256, use_graphs=graphs)
    g_img, g_txt = banks
    losses = []
    cl.begin_round()
    for s in range(steps):
        images, caps, lens, d_idx = _client_inputs(16, 24, g_img.shape[0], 300 + s)
        losses.append(cl.private_step(images, caps, lens).item())
        losses.append(cl.contrast_step(images, caps, lens, d_idx, g_img, g_txt).item())
    fi, ft = cl.generate(*_client_inputs(16, 24, g_img.shape[0], 400)[:3])
    torch.cuda.synchronize()
    out = _groups(cl)
    out['generated'] = torch.cat([fi.flatten(), ft.flatten()]).clone()
    return cl, out, losses


def test_client_graph_replay_equals_eager(engine):
    g = torch.Generator().manual_seed(1)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    banks = (unit(torch.randn(4096, 256, generator=g)).cuda(), unit(torch.randn(4096, 256, generator=g)).cuda())
    # one step pair from identical state: only one update's worth of atomics noise separates two runs
    _, eager1, l_eager1 = _run_client(engine, False, banks, steps=1)
    noise1 = [_run_client(engine, False, banks, steps=1)[1] for _ in range(3)]
    _, graph1, l_graph1 = _run_client(engine, True, banks, steps=1)
    _assert_within_noise(graph1, eager1, noise1, 'client graph vs eager, 1 step')
    assert abs(l_graph1[0] - l_eager1[0]) <= 1e-5 * abs(l_eager1[0])        # private-step loss: identical parameters
    # three step pairs: the noise has been amplified by the earlier updates on both sides
    _, eager, _ = _run_client(engine, False, banks)
    noise = [_run_client(engine, False, banks)[1] for _ in range(4)]
    cl, graph, _ = _run_client(engine, True, banks)
    _assert_within_noise(graph, eager, noise, 'client graph vs eager')
    # different caption-length profiles reuse ONE graph per step kind (lengths are a graph input, not a key)
    assert len(cl._cache.graphs) == 3, list(cl._cache.graphs)
    assert float(cl.optimizer._state[0]) == 6.0


def test_clients_in_concurrent_lanes_equal_one_after_the_other(engine):
    """bench.py runs the clients hosted by one GPU in concurrent execution lanes (engine.lane: own stream, scratch
    buffers and graph pool per lane).  Three graphed clients stepping interleaved in three lanes end where the same
    clients end when they run one after the other on the default stream (within the atomics noise of repeated runs):
    no scratch buffer, graph pool or stream is shared across lanes.  Runs in a child process with a time limit
    (tests/lanes_worker.py): a GPU-side hang must fail this test, not stall the suite.  One bench process in ~10
    with lanes stalled once during development and never reproduced (known issue): a child that
    exceeds its time limit is killed and started ONCE more, with a warning; a wrong result or a second stall fails."""
    import warnings
    res = None
    for attempt in range(2):
        try:
            res = subprocess.run([sys.executable, str(ROOT / 'tests' / 'lanes_worker.py')], capture_output=True,
                                 text=True, timeout=300, cwd=str(ROOT))
            break
        except subprocess.TimeoutExpired:
            if attempt == 1:
                raise
            warnings.warn('lanes_worker.py exceeded 300 s (stalled lanes, DESIGN.md section 5); running it once more')
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert 'LANES OK' in res.stdout


def test_banks_refreshed_in_place_reach_the_captured_contrast_step(engine):
    g = torch.Generator().manual_seed(2)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    g_img, g_txt = unit(torch.randn(2048, 256, generator=g)).cuda(), unit(torch.randn(2048, 256, generator=g)).cuda()
    torch.manual_seed(11)
    cl = engine.MMClient(256, use_graphs=True)
    cl.begin_round()
    images, caps, lens, d_idx = _client_inputs(8, 16, 2048, 500)
    l0 = cl.contrast_step(images, caps, lens, d_idx, g_img, g_txt).item()
    torch.manual_seed(11)
    ce = engine.MMClient(256, use_graphs=False)
    ce.begin_round()
    assert abs(ce.contrast_step(images, caps, lens, d_idx, g_img, g_txt).item() - l0) < 1e-3 * abs(l0)
    # a new round's server features (new tensors): the graph must see them
    g_img2, g_txt2 = unit(g_img + 0.5 * torch.randn_like(g_img)), unit(g_txt + 0.5 * torch.randn_like(g_txt))
    l1 = cl.contrast_step(images, caps, lens, d_idx, g_img2, g_txt2).item()
    l1e = ce.contrast_step(images, caps, lens, d_idx, g_img2, g_txt2).item()
    assert abs(l1 - l1e) < 2e-3 * abs(l1e)
    assert len(cl._cache.graphs) == 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_data_parallel_server_step_equals_single_process_mean_gradient(engine, tmp_path):
    """2 ranks over NCCL, one batch each: the replicated server's step == one process averaging the two gradients."""
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run with gpurun --gpus 2)')
    out = tmp_path / 'dp.pt'
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                        '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
                        str(ROOT / 'tests' / 'dp_worker.py'), str(out)], capture_output=True, text=True, env=env,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = torch.load(out)
    assert res['ranks_equal'] <= 1e-7, res                      # both replicas hold the same parameters after the step
    assert res['dp_vs_single'] <= 4 * res['single_vs_single'] + 1e-7, res
    assert res['graph_dp_vs_single'] <= 4 * res['single_vs_single'] + 1e-7, res
