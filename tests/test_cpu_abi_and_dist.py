"""CPU tests (no GPU): the C-ABI library loads and exports every symbol the header declares; the product refuses to
run without CUDA; the multi-rank exchange plumbing (gloo, world size 2) reproduces the single-process con_w
aggregation of the oracle."""
import ctypes
import os
import re
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_library_builds_and_exports_every_declared_symbol():
    from creamfl_b200 import build, _lib
    lib_path = build.build()
    handle = ctypes.CDLL(str(lib_path))
    header = (ROOT / 'include' / 'creamfl_b200.h').read_text()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(creamfl_\w+)\s*\(', header))
    assert len(declared) >= 45
    assert declared == set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.creamfl_abi_version() == 1


def test_error_channel_without_gpu():
    """Argument validation happens before any CUDA call: the status / message channel works on a CPU box."""
    from creamfl_b200 import _lib
    lib = _lib.load()
    rc = lib.creamfl_conw_reduce(None, None, 0, 0, 0, None, None, None)
    assert rc == -1 and b'null pointer' in lib.creamfl_last_error()
    assert lib.creamfl_rowlse_workspace_bytes(0, 0) == 0
    assert lib.creamfl_conv2d_workspace_bytes(2, 56, 56, 64, 64, 3, 3, 1, 1) == 0          # implicit GEMM: no patch matrix
    assert lib.creamfl_conv2d_workspace_bytes(2, 56, 56, 64, 128, 3, 3, 2, 1) == 2 * 28 * 28 * 576 * 2


def test_no_cpu_fallback():
    from creamfl_b200 import ops, towers
    with pytest.raises(RuntimeError):
        ops.pcme_loss(torch.zeros(4, 8), torch.zeros(4, 8), torch.tensor(1.0), torch.tensor(1.0))
    with pytest.raises(RuntimeError):
        ops.conw_score(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))
    model = towers.ImageModel({'embed_dim': 64, 'cnn_type': 'resnet18'})
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 3, 224, 224))


def test_state_dict_keys_match_reference_modules():
    """Checkpoint compatibility (MMFL.py:281): PCME keys = torchvision ResNet + HF BERT + the reference's heads."""
    import torchvision
    from transformers import BertConfig, BertModel
    from creamfl_b200 import towers
    sd = towers.PCME(None, {'embed_dim': 256, 'cnn_type': 'resnet18'}).state_dict()
    tv = {k for k in torchvision.models.resnet18(weights=None).state_dict() if not k.startswith('fc.')}
    assert {k[len('img_enc.cnn.'):] for k in sd if k.startswith('img_enc.cnn.')} == tv
    hf = set(BertModel(BertConfig(num_hidden_layers=2)).state_dict())
    mine = {k[len('txt_enc.'):] for k in sd if k.startswith('txt_enc.') and not re.search(r'layer\.([2-9]|1[01])\.', k)}
    assert mine == hf
    for k in ('linear.weight', 'linear.bias', 'img_enc.fc.weight', 'img_enc.pie_net.attention.w_1.weight',
              'img_enc.pie_net.attention.w_2.weight', 'img_enc.pie_net.fc.bias', 'img_enc.pie_net.layer_norm.weight'):
        assert k in sd


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, d, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from creamfl_b200 import engine
    from oracle import creamfl_oracle as O
    from conftest import conw_inputs
    g_img, g_txt, i_vecs, _ = conw_inputs(11, n, d, world)
    own = torch.from_numpy(i_vecs[rank])
    score = O.conw_scores(own.double(), torch.from_numpy(g_txt).double()).float()
    scores, vecs = engine.gather_client_rows(score, own)          # the exchange under test
    w = torch.softmax(scores.double(), 0)
    agg = (vecs.double() * w[:, :, None]).sum(0)
    grad = torch.full((5,), float(rank + 1))
    engine.average_gradients(grad)
    q.put((rank, agg.numpy(), tuple(scores.shape), tuple(vecs.shape), grad.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_gloo_world2_matches_single_process_oracle():
    from oracle import creamfl_oracle as O
    from conftest import conw_inputs
    n, d, world = 384, 64, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g_img, g_txt, i_vecs, _ = conw_inputs(11, n, d, world)
    want, _ = O.conw_aggregate([torch.from_numpy(v).double() for v in i_vecs], torch.from_numpy(g_txt).double())
    for rank, agg, s_shape, v_shape, grad in results:
        assert s_shape == (world, n) and v_shape == (world, n, d)
        np.testing.assert_allclose(agg, want.numpy(), rtol=1e-5, atol=1e-7)     # every rank holds the full ensemble (fp32 scores on the wire)
        np.testing.assert_allclose(grad, np.full(5, 1.5))                       # mean of the rank gradients


def _worker_absent_slots(rank, world, port, n, d, q):
    """configs[2] on two ranks: each hosts 3 client slots, some without the modality (image-only / text-only clients
    return None for the other one, ClientTrainer.py:622-629)."""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, str(ROOT / 'tests'))
    import kernel_emulation as KE
    from creamfl_b200 import engine, ops
    from conftest import conw_inputs
    for name in ('conw_score', 'conw_reduce', 'to_bf16'):         # test-only CPU stand-ins of the CUDA wrappers
        setattr(ops, name, getattr(KE, name))
    g_img, g_txt, vecs, _ = conw_inputs(23, n, d, 4)
    layout = [[True, False, True], [False, False, True]]         # slot 1 is empty on every rank: it must not travel
    clients = {(0, 0): 0, (0, 2): 1, (1, 2): 2}
    own = [torch.from_numpy(vecs[clients[(rank, s)]]) if layout[rank][s] else None for s in range(3)]
    g16 = torch.from_numpy(g_txt).to(torch.bfloat16)
    agg_given = engine.exchange_and_aggregate_clients(own, g16, layout)
    agg_found = engine.exchange_and_aggregate_clients(own, g16)          # layout exchanged by the function itself
    none = engine.exchange_and_aggregate_clients([None, None, None], g16)
    try:
        engine.exchange_and_aggregate_clients(own, g16, [[True, True, True], [True, True, True]])
        mismatch = False
    except ValueError:
        mismatch = True
    q.put((rank, agg_given.numpy(), agg_found.numpy(), none is None, mismatch))
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_with_absent_modality_slots_gloo_world2():
    """BASELINE configs[2] (image-only + text-only clients): the exchange aggregates over the clients that carry the
    modality, wherever they are hosted (MMFL.py:226-247,298-331) - equal to the single-process aggregation of exactly
    those clients."""
    sys.path.insert(0, str(ROOT / 'tests'))
    import kernel_emulation as KE
    from conftest import conw_inputs
    n, d, world = 256, 32, 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_absent_slots, args=(r, world, port, n, d, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g_img, g_txt, vecs, _ = conw_inputs(23, n, d, 4)
    want = KE.conw_aggregate([torch.from_numpy(vecs[c]) for c in (0, 1, 2)], torch.from_numpy(g_txt)).numpy()
    for rank, agg_given, agg_found, none_is_none, mismatch in results:
        np.testing.assert_allclose(agg_given, want, rtol=1e-5, atol=1e-7)
        np.testing.assert_array_equal(agg_given, agg_found)
        assert none_is_none and mismatch


def test_split_k_planner_fills_whole_waves():
    """Host logic of the accumulating (weight-gradient) GEMMs: `creamfl_plan_split_k` must return a distinct
    partition of the k-blocks (>= 4 per unit) whose tiles * split units fill whole waves of the 148 SMs on the shapes
    the server step issues most (ResNet101 layer 3 / 4 1x1 convolutions at batch 128, BERT-base at 4096 tokens)."""
    import math
    from creamfl_b200 import _lib
    lib = _lib.load()
    shapes = [(1024, 256, 25088), (256, 1024, 25088), (2048, 512, 6272), (512, 2048, 6272), (768, 768, 4096),
              (3072, 768, 4096), (768, 3072, 4096), (2304, 768, 4096), (512, 128, 100352), (128, 512, 100352),
              (256, 64, 401408), (64, 256, 401408)]
    for m, n, k in shapes:
        s = lib.creamfl_plan_split_k(m, n, k)
        bn = 64 if n <= 64 else (256 if ((n >= 256 and n % 256 == 0) or n >= 1024) and k >= 512 else 128)
        nkb = math.ceil(k / 64)
        per = math.ceil(nkb / s)
        assert s >= 1 and per >= 4 and math.ceil(nkb / per) == s, (m, n, k, s)
        units = math.ceil(m / 128) * math.ceil(n / bn) * s
        waves = math.ceil(units / 148)
        assert units / (waves * 148) >= 0.85, (m, n, k, s, units)
    assert lib.creamfl_plan_split_k(256, 768, 128) == 1          # two k-blocks: nothing to split
    assert lib.creamfl_plan_split_k(0, 5, 5) == 1
