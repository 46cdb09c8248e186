#!/usr/bin/env python
"""Benchmark of the CreamFL hot path on B200 (contract: python bench.py --gpus N --steps K --warmup W).

Workload = BASELINE.json configs[1]: ResNet101+BERT server, 8 multimodal clients (ResNet18+GRU), COCO-shape synthetic
public batches of 128 pairs, inter+intra contrast against a 50 000-row public bank, con_w aggregation.  (--config 2:
configs[2], 4 image + 4 text unimodal clients.)  The 8 clients are sharded over the N GPUs (8/N per GPU; 8 GPUs = one
client per GPU as the config says), the total work is the same at every N: "scaling": "strong".

One "step" is one mini-round over a FIXED public set of S = 8 batches (what the reference's MMFL.train does per round,
with the 391-batch public loader shortened to S; the public bank keeps its full 50 000 rows):
  A  server train       the server sees the public set once: S/N data-parallel steps (rank r takes batches r, r+N, ..;
                        ResNet101+BERT fwd/bwd with BERT dropout, PCME loss, gradient all-reduce, clip, AdamP)
                                                                                       retrieval_trainer.py:192-214
  B  server extraction  S/N eval forwards per rank, rows all-gathered into the global banks    MMFL.py:194-221
  C  clients            every client: old_model <- model; 1 private step; S contrast steps over ALL public batches
                        (client fwd/bwd + old-model fwd + inter/intra + AdamP)         MMClientTrainer.py:91-222
  D  client generate    every client: S eval forwards -> rows of its public representations   MMClientTrainer.py:326-359
  E  con_w aggregation  score [50000 x 50000] per client and modality, all-gather of scores + representations over
                        NCCL (the one exchange of the path), softmax-over-clients reduce          MMFL.py:298-335
  F  server distill     S/N data-parallel steps (fwd/bwd, kd MSE to aggregated rows, clip, AdamP)     MMFL.py:346-391
value = pairs pushed through an encoder forward(+backward) step per second, whole job:
3*S*B server pair-passes + C*(1+2S)*B client pair-passes per mini-round (the old-model forward is not counted), inputs
resident in HBM; e2e = the same with the step's inputs copied from pinned host memory inside the timed region and the
step's losses read back.  `public_pairs_per_s` (S*B public pairs per mini-round) is kept for comparison with round 1.

--impl reference times the reference's CPU implementation of the same mini-round (torch restatement in oracle/, fp32,
all host threads, AdamP restated) on a bounded sample per step; see cpu_baseline.sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D = 256
N_PUB = 50000
BERT_L = 32
CAP_L = 30
TXT_L = 60                       # AG_NEWS-shape private captions of the unimodal text clients (configs[2])
VOCAB = 11755
METRIC = 'image-text pairs/sec per FL round'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='creamfl_b200', choices=['creamfl_b200', 'reference'])
    ap.add_argument('--config', type=int, default=1, choices=[1, 2], help='BASELINE.json configs[1] or configs[2]')
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--public-batches', type=int, default=8, help='S: public batches per mini-round')
    ap.add_argument('--clients', type=int, default=8)
    ap.add_argument('--client-lanes', type=int, default=2,
                    help='clients hosted by one GPU run in this many concurrent execution lanes (1: one after the other)')
    ap.add_argument('--lane-guard-seconds', type=int, default=240,
                    help='with --client-lanes > 1: if the measured part has not finished after this many seconds the '
                         'process re-executes itself with --client-lanes 1 (0: no guard)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-gpu-reference', action='store_true')
    ap.add_argument('--phases', default='ABCDEF', help='debug: subset of phases to run')
    ap.add_argument('--no-graphs', action='store_true', help='launch every kernel eagerly instead of CUDA graphs')
    ap.add_argument('--no-dropout', action='store_true', help='debug: BERT dropout off (the reference trains with 0.1)')
    return ap.parse_args()


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def workload_config(args, world):
    S, B, C = args.public_batches, args.batch, args.clients
    kind = ('8 multimodal clients (ResNet18+GRU)' if args.config == 1 else
            f'{C // 2} image (ResNet18, CIFAR-shape) + {C - C // 2} text (GRU, AG_NEWS-shape) unimodal clients')
    return {'workload': f'configs[{args.config}]: ResNet101+BERT server, {kind}, COCO-shape synthetic public batches of '
                        f'{B} pairs, inter+intra contrast vs N_pub={N_PUB}, con_w aggregation',
            'global_batch': B, 'public_batches_per_step': S, 'clients': C, 'n_pub': N_PUB, 'embed_dim': D,
            'bert_seq_len': BERT_L, 'bert_dropout': 0.0 if args.no_dropout else 0.1, 'phases': args.phases,
            'cuda_graphs': not args.no_graphs,
            'pairs_definition': 'SURVEY 8d: pairs pushed through an encoder fwd(+bwd) step = 3*S*B (server A, B, F) + '
                                'per client (1+2S)*B (private, contrast, generate); old-model forward not counted',
            'l2': 'inputs_exceed_l2 (0.6 GB images + 0.6 GB parameters per step >> 126 MB L2)',
            'parallelism': f'{C} clients sharded over {world} GPU(s) ({C // world} per GPU); server phases data-parallel '
                           f'over the fixed public set with flat-gradient all-reduce; NCCL all-gather of public '
                           f'representations' if world > 1 else 'single gpu',
            'client_lanes': f'the {C // world} client(s) of a GPU run in {max(1, min(args.client_lanes, C // world))} '
                            f'concurrent execution lane(s) (own stream, scratch and graph pool each)'}


def pairs_per_step(args):
    S, B, C = args.public_batches, args.batch, args.clients
    return 3 * S * B + C * (1 + 2 * S) * B


class stdout_to_stderr:
    """File-descriptor level redirect of stdout to stderr (catches C-level printf of libraries, e.g. the
    `NCCL version ...` banner NCCL writes to stdout when NCCL_DEBUG=VERSION)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)                       # flush C stdio buffers before restoring the fd
        except OSError:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


# ===================================================================================================== synthetic data
def make_public(S, B, seed, pin=True):
    """The fixed public set of a mini-round (identical on every rank) - SURVEY.md 8d recipe: N(0,1) images, BERT ids
    U[1000, 30522) with CLS/SEP, lengths U{8..32} sorted descending, vocab-id captions U[4, 11755) with lengths
    U{5..30} sorted descending, bank rows = a random subset."""
    g = torch.Generator().manual_seed(seed)
    p = (lambda t: t.pin_memory()) if (pin and torch.cuda.is_available()) else (lambda t: t)
    out = {'images': p(torch.randn(S, B, 3, 224, 224, generator=g))}
    lens = torch.sort(torch.randint(8, BERT_L + 1, (S, B), generator=g), dim=1, descending=True).values
    lens[:, 0] = BERT_L
    ids = torch.randint(1000, 30522, (S, B, BERT_L), generator=g)
    mask = (torch.arange(BERT_L)[None, None, :] < lens[:, :, None]).long()
    ids[:, :, 0] = 101
    ids.scatter_(2, (lens - 1).unsqueeze(-1), 102)
    out['ids'], out['mask'] = p(ids * mask), p(mask)
    clen = torch.sort(torch.randint(5, CAP_L + 1, (S, B), generator=g), dim=1, descending=True).values
    clen[:, 0] = CAP_L
    cmask = (torch.arange(CAP_L)[None, None, :] < clen[:, :, None]).long()
    out['caps'] = p(torch.randint(4, VOCAB, (S, B, CAP_L), generator=g) * cmask)
    out['cap_lens'] = p(clen.to(torch.int32))
    out['d_idx'] = p(torch.stack([torch.randperm(N_PUB, generator=g)[:B] for _ in range(S)]))
    return out


def make_private(kind, B, seed, pin=True):
    """One private batch of a client: multimodal (COCO/Flickr-shape pair batch), image (CIFAR-100-shape, upsampled to
    224 here like the public images so one graph serves both passes) or text (AG_NEWS-shape)."""
    g = torch.Generator().manual_seed(seed)
    p = (lambda t: t.pin_memory()) if (pin and torch.cuda.is_available()) else (lambda t: t)
    if kind == 'mm':
        clen = torch.sort(torch.randint(5, CAP_L + 1, (B,), generator=g), descending=True).values
        clen[0] = CAP_L
        cmask = (torch.arange(CAP_L)[None, :] < clen[:, None]).long()
        return {'images': p(torch.randn(B, 3, 224, 224, generator=g)),
                'caps': p(torch.randint(4, VOCAB, (B, CAP_L), generator=g) * cmask), 'cap_lens': p(clen.to(torch.int32))}
    if kind == 'image':
        return {'images': p(torch.randn(B, 3, 224, 224, generator=g)), 'labels': p(torch.randint(0, 100, (B,), generator=g))}
    clen = torch.sort(torch.randint(10, TXT_L + 1, (B,), generator=g), descending=True).values
    clen[0] = TXT_L
    cmask = (torch.arange(TXT_L)[None, :] < clen[:, None]).long()
    return {'caps': p(torch.randint(4, VOCAB, (B, TXT_L), generator=g) * cmask), 'cap_lens': p(clen.to(torch.int32)),
            'labels': p(torch.randint(0, 4, (B,), generator=g))}


def nbytes(d):
    return sum(v.numel() * v.element_size() for v in d.values())


def make_banks(device, seed):
    g = torch.Generator().manual_seed(seed)
    g_img = unit(torch.randn(N_PUB, D, generator=g))
    g_txt = unit(0.7 * g_img + 0.5 * unit(torch.randn(N_PUB, D, generator=g)))
    return g_img.to(device), g_txt.to(device)


def client_kinds(args):
    if args.config == 1:
        return ['mm'] * args.clients
    half = args.clients // 2
    return ['image'] * half + ['text'] * (args.clients - half)          # ranks 0-3 image, 4-7 text at 8 GPUs


# ===================================================================================================== clocks
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '100'], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        # median of the busy half (the sampler also sees the idle gaps around the timed region)
        busy = sm[len(sm) // 2:] if sm else []
        return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def ncu_traffic_bytes(csv_path, column=0):
    """dram read + write bytes of launch `column` in a profiles/*_ncu_*.csv summary (None if absent)."""
    try:
        tot = 0.0
        for ln in Path(csv_path).read_text().splitlines():
            k = ln.split(',')
            if k[0] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[k[1]]
                tot += float(k[2 + column]) * scale
        return int(tot) if tot else None
    except (OSError, KeyError, ValueError, IndexError):
        return None


def ncu_counters(csv_path, column=0):
    """A few headline counters of launch `column` in a profiles/*_ncu_*.csv summary ({} if absent): evidence that
    rides with the roofline entry, measured under ncu in an earlier call (never a timing of this run)."""
    want = {'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed': 'tensor_pipe_pct_elapsed',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active': 'tensor_pipe_pct_active',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct',
            'lts__t_sector_hit_rate.pct': 'l2_hit_pct', 'gpu__time_duration.sum': 'ncu_duration_us'}
    out = {}
    try:
        for ln in Path(csv_path).read_text().splitlines():
            k = ln.split(',')
            if k[0] in want:
                out[want[k[0]]] = round(float(k[2 + column]), 2)
    except (OSError, ValueError, IndexError):
        return {}
    return out


def load_peaks():
    try:
        return json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())
    except OSError:
        return {}


# ===================================================================================================== our arm
class Round:
    """One mini-round on this rank: owns the server replica, the local clients, the banks and the step function."""

    def __init__(self, args, rank, world, dev):
        import torch.distributed as dist
        from creamfl_b200 import engine, ops
        self.args, self.rank, self.world, self.dev = args, rank, world, dev
        self.engine, self.ops, self.dist = engine, ops, dist
        S, B, C = args.public_batches, args.batch, args.clients
        if S % world or C % world:
            raise SystemExit(f'bench.py: --public-batches {S} and --clients {C} must be multiples of --gpus {world}')
        graphs = not args.no_graphs
        torch.manual_seed(1234)                                   # identical server replicas on every rank
        self.server = engine.ServerEngine(D, 'resnet101', device=dev, data_parallel=world > 1, use_graphs=graphs,
                                          bert_dropout=0.0 if args.no_dropout else 0.1)
        kinds = client_kinds(args)
        self.local = [c for c in range(C) if c % world == rank]     # client ids hosted by this rank
        self.kinds = kinds
        self.clients = []
        for c in self.local:
            torch.manual_seed(5000 + c)
            if kinds[c] == 'mm':
                self.clients.append(engine.MMClient(D, device=dev, use_graphs=graphs))
            else:
                self.clients.append(engine.UnimodalClient(kinds[c], 100 if kinds[c] == 'image' else 4, embed_dim=D,
                                                          device=dev, use_graphs=graphs))
        # which (client, modality) slots exist, for every rank (static: known from the configuration)
        per_rank = C // world
        self.layout_img = [[kinds[c] in ('mm', 'image') for c in range(C) if c % world == r] for r in range(world)]
        self.layout_txt = [[kinds[c] in ('mm', 'text') for c in range(C) if c % world == r] for r in range(world)]
        assert all(len(row) == per_rank for row in self.layout_img)
        self.g_img, self.g_txt = make_banks(dev, 99)                # server features (same on every rank)
        self.g_img16 = torch.empty_like(self.g_img, dtype=torch.bfloat16)
        self.g_txt16 = torch.empty_like(self.g_txt, dtype=torch.bfloat16)
        gen = torch.Generator().manual_seed(777 + rank)
        self.c_img = [unit(self.g_img.cpu() + 0.5 * unit(torch.randn(N_PUB, D, generator=gen))).to(dev)
                      if kinds[c] in ('mm', 'image') else None for c in self.local]
        self.c_txt = [unit(self.g_txt.cpu() + 0.5 * unit(torch.randn(N_PUB, D, generator=gen))).to(dev)
                      if kinds[c] in ('mm', 'text') else None for c in self.local]
        self.public_host = make_public(S, B, 4321)
        self.private_host = [make_private(kinds[c], B, 9000 + c) for c in self.local]
        self.h2d_bytes = nbytes(self.public_host) + sum(nbytes(p) for p in self.private_host)
        self.coll_ms = {}

    def to_device(self, non_blocking=True):
        dev = self.dev
        pub = {k: v.to(dev, non_blocking=non_blocking) for k, v in self.public_host.items()}
        priv = [{k: v.to(dev, non_blocking=non_blocking) for k, v in p.items()} for p in self.private_host]
        return pub, priv

    def _timed_collective(self, name, fn):
        """CUDA-event time of one exchange (recorded only when asked: the events stay off the hot loop otherwise)."""
        if self.coll_ms is None:
            return fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        self.coll_ms.setdefault(name, []).append((a, b))
        return out

    def step(self, pub, priv, marks=None):
        args, server, dist, engine, ops = self.args, self.server, self.dist, self.engine, self.ops
        S, world, rank, phases = args.public_batches, self.world, self.rank, args.phases

        def mark(name):
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev))
        mark('start')
        losses = []
        tok = lambda s: {'input_ids': pub['ids'][s], 'attention_mask': pub['mask'][s]}
        mine = list(range(rank, S, world))                          # this rank's share of the public set
        if 'A' in phases:
            for s in mine:
                losses.append(server.train_step(pub['images'][s], tok(s)))
        mark('A_server_train')
        if 'B' in phases:
            fi, ft = [], []
            for s in mine:
                a, b = server.extract(pub['images'][s], tok(s))
                fi.append(a.clone())
                ft.append(b.clone())
            fi, ft = torch.stack(fi), torch.stack(ft)               # [S/N, B, D]
            if world > 1:
                def gather():
                    gi = torch.empty((world,) + tuple(fi.shape), dtype=fi.dtype, device=fi.device)
                    gt = torch.empty_like(gi)
                    dist.all_gather_into_tensor(gi.view(-1), fi.view(-1))
                    dist.all_gather_into_tensor(gt.view(-1), ft.view(-1))
                    return gi, gt
                gi, gt = self._timed_collective('extract_allgather', gather)
                for r in range(world):
                    for k, s in enumerate(range(r, S, world)):
                        self.g_img.index_copy_(0, pub['d_idx'][s], gi[r, k])
                        self.g_txt.index_copy_(0, pub['d_idx'][s], gt[r, k])
            else:
                for k, s in enumerate(mine):
                    self.g_img.index_copy_(0, pub['d_idx'][s], fi[k])
                    self.g_txt.index_copy_(0, pub['d_idx'][s], ft[k])
        ops.cast_into(self.g_img.view(-1), self.g_img16.view(-1))
        ops.cast_into(self.g_txt.view(-1), self.g_txt16.view(-1))
        mark('B_server_extract')
        # clients hosted by this rank share nothing inside a round: they run in `--client-lanes` execution lanes (own
        # stream, scratch and graph pool each - engine.lane), client i in lane 1 + i % lanes, enqueued step by step in
        # round-robin so every lane always has work queued; steps of one client stay ordered on its lane's stream
        lanes = max(1, min(args.client_lanes, len(self.clients)))
        lane_ids = [1 + k for k in range(lanes)]
        lane_of = lambda ci: engine.lane(self.dev, 1 + ci % lanes)
        per_client = [[] for _ in self.clients]
        if 'C' in phases:
            engine.lane.fork(self.dev, lane_ids)
            for ci, cl in enumerate(self.clients):
                kind, pv = self.kinds[self.local[ci]], priv[ci]
                with lane_of(ci):
                    cl.begin_round()
                    if kind == 'mm':
                        per_client[ci].append(cl.private_step(pv['images'], pv['caps'], pv['cap_lens']))
                    elif kind == 'image':
                        per_client[ci].append(cl.supervised_step(pv['images'], pv['labels']))
                    else:
                        per_client[ci].append(cl.supervised_step(pv['caps'], pv['labels'], pv['cap_lens']))
            for s in range(S):
                for ci, cl in enumerate(self.clients):
                    kind = self.kinds[self.local[ci]]
                    with lane_of(ci):
                        if kind == 'mm':
                            per_client[ci].append(cl.contrast_step(pub['images'][s], pub['caps'][s], pub['cap_lens'][s],
                                                                   pub['d_idx'][s], self.g_img, self.g_txt))
                        elif kind == 'image':
                            per_client[ci].append(cl.contrast_step(pub['images'][s], None, pub['d_idx'][s], self.g_img,
                                                                   self.g_txt))
                        else:
                            per_client[ci].append(cl.contrast_step(pub['caps'][s], pub['cap_lens'][s], pub['d_idx'][s],
                                                                   self.g_txt, self.g_img))
            engine.lane.join(self.dev, lane_ids)
            for lst in per_client:
                losses.extend(lst)
        mark('C_client_private_contrast')
        if 'D' in phases:
            engine.lane.fork(self.dev, lane_ids)
            for s in range(S):
                for ci, cl in enumerate(self.clients):
                    kind = self.kinds[self.local[ci]]
                    with lane_of(ci):
                        if kind == 'mm':
                            a, b = cl.generate(pub['images'][s], pub['caps'][s], pub['cap_lens'][s])
                            self.c_img[ci].index_copy_(0, pub['d_idx'][s], a)
                            self.c_txt[ci].index_copy_(0, pub['d_idx'][s], b)
                        elif kind == 'image':
                            self.c_img[ci].index_copy_(0, pub['d_idx'][s], cl.generate(pub['images'][s]))
                        else:
                            self.c_txt[ci].index_copy_(0, pub['d_idx'][s], cl.generate(pub['caps'][s], pub['cap_lens'][s]))
            engine.lane.join(self.dev, lane_ids)
        mark('D_client_generate')
        agg_img = agg_txt = None
        if 'E' in phases:
            agg_img = self._timed_collective('conw_img', lambda: engine.exchange_and_aggregate_clients(
                self.c_img, self.g_txt16, layout=self.layout_img))
            agg_txt = self._timed_collective('conw_txt', lambda: engine.exchange_and_aggregate_clients(
                self.c_txt, self.g_img16, layout=self.layout_txt))
        mark('E_conw_aggregate')
        if 'F' in phases:
            if agg_img is None:
                agg_img, agg_txt = self.g_img, self.g_txt
            # the reference adds each MSE once per client TYPE carrying the modality (MMFL.py:361-378)
            types = set(self.kinds)
            it = int('image' in types) + int('mm' in types)
            tt = int('text' in types) + int('mm' in types)
            for s in mine:
                losses.append(server.distill_step(pub['images'][s], tok(s), pub['d_idx'][s], agg_img, agg_txt,
                                                  img_terms=it, txt_terms=tt))
        mark('F_server_distill')
        return torch.stack([l.reshape(()) for l in losses]) if losses else torch.zeros(1, device=self.dev)


def family_table(rnd, pub, peaks):
    """One eager (un-graphed, single-stream) server train step with CUDA events around every C-ABI call, the host
    running ahead of the device (calltimer.py): where the time of the dominant phase goes, by KERNEL family, with each
    family's algorithmic FLOP/s or GB/s against the measured peaks.  gemm_tc = BERT linears + 1x1 convolutions (+ the
    strided convolutions through im2col); conv_tc = 3x3 implicit GEMM; bn = BatchNorm statistics / apply / backward."""
    from creamfl_b200.calltimer import CallTimer
    server = rnd.server
    txt = {'ids': pub['ids'][0], 'mask': pub['mask'][0]}
    was_dp, server.data_parallel = server.data_parallel, False
    server.model.overlap_towers = False                     # one stream: an event pair brackets one call's kernels only
    for _ in range(2):                                      # re-warm the eager (non-graph) allocator pool
        server._train_step(pub['images'][0], txt)
    torch.cuda.synchronize()
    with CallTimer() as t:
        t.stall(120.0)
        server._train_step(pub['images'][0], txt)
        raw = t.families()
        overhead_us = t.overhead_ms() * 1e3
    server.data_parallel = was_dp
    server.model.overlap_towers = True
    group = lambda k: 'gemm_tc' if k.startswith('gemm_tc') else ('conv_tc' if k.startswith('conv_tc') else
                                                                  ('bn' if k.startswith('bn_') else k))
    fam = {}
    for k, f in raw.items():
        g = fam.setdefault(group(k), {'ms': 0.0, 'calls': 0, 'flops': 0.0, 'bytes': 0.0, 'parts': {}})
        for key in ('ms', 'calls', 'flops', 'bytes'):
            g[key] += f[key]
        if group(k) != k:
            g['parts'][k] = round(f['ms'], 3)
    total = sum(f['ms'] for f in fam.values())
    peak_tf, peak_gb = peaks.get('bf16_tflops_sustained', 1400.0), peaks.get('hbm_gbs', 6400.0)
    table = {}
    for name, f in sorted(fam.items(), key=lambda kv: -kv[1]['ms']):
        row = {'ms': round(f['ms'], 3), 'share': round(f['ms'] / total, 4), 'calls': f['calls']}
        if f['flops'] > 0:
            tf = f['flops'] / (f['ms'] * 1e-3) / 1e12
            row.update(tflops=round(tf, 1), frac_tensor=round(tf / peak_tf, 4))
        if f['bytes'] > 0:
            gb = f['bytes'] / (f['ms'] * 1e-3) / 1e9
            row.update(gbps=round(gb, 1), frac_hbm=round(gb / peak_gb, 4))
        if f['parts']:
            row['parts_ms'] = f['parts']
        table[name] = row
    return table, total, fam, overhead_us


def run_ours(args):
    import torch.distributed as dist
    from creamfl_b200 import ops
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        with stdout_to_stderr():     # stdout carries exactly one JSON line: NCCL's version banner goes to stderr
            dist.init_process_group('nccl', device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))          # forces communicator creation inside the redirect
            torch.cuda.synchronize()
    # Concurrent client lanes hung ONCE in ~10 runs during development (one of three consecutive bench processes on a
    # box; never reproduced, cause unknown - DESIGN.md section 5).  A hung GPU context cannot be recovered in-process, so
    # the measured part runs under a timer that re-executes this very process (same PID, same torchrun rendezvous
    # environment) with the lanes switched off.  Nothing has been printed at that point.
    guard = None
    if args.client_lanes > 1 and args.lane_guard_seconds > 0:
        import threading

        def _reexec():
            argv, skip = [], False
            for a in sys.argv[1:]:
                if skip:
                    skip = False
                elif a == '--client-lanes':
                    skip = True
                elif not a.startswith('--client-lanes='):
                    argv.append(a)
            sys.stderr.write(f'bench.py: no result after {args.lane_guard_seconds} s with --client-lanes '
                             f'{args.client_lanes}; re-executing with --client-lanes 1\n')
            sys.stderr.flush()
            os.execv(sys.executable, [sys.executable, os.path.abspath(sys.argv[0])] + argv + ['--client-lanes', '1'])
        guard = threading.Timer(args.lane_guard_seconds, _reexec)
        guard.daemon = True
        guard.start()
    rnd = Round(args, rank, world, dev)
    pub, priv = rnd.to_device(non_blocking=False)
    S, B = args.public_batches, args.batch

    from creamfl_b200.prefetch import Prefetcher
    pf = Prefetcher(dev, depth=2)

    def timed(copy_in, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launches()
        t0.record()
        out = None
        if copy_in:
            # every step's inputs travel from pinned host memory inside the timed region; the copy of step k + 1 runs on
            # the prefetcher's stream while step k computes (creamfl_b200/prefetch.py), the first one is exposed
            host = (rnd.public_host, rnd.private_host)
            pf.submit(host)
        for i in range(n):
            if copy_in:
                p, q = pf.next()
                if i + 1 < n:
                    pf.submit(host)
                out = rnd.step(p, q).cpu()             # device -> host read of the step's losses
                pf.release()
            else:
                out = rnd.step(pub, priv)
        t1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, ops.launches() - l0, out

    rnd.coll_ms = None
    for _ in range(args.warmup):
        rnd.step(pub, priv)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ms_res, launches, out = timed(False, args.steps)
    clocks = sampler.stop()
    # where the step's time goes: one more resident step with CUDA events at the phase boundaries and around the
    # collectives (kept out of the timed region)
    marks = []
    rnd.coll_ms = {}
    rnd.step(pub, priv, marks)
    torch.cuda.synchronize()
    phase_ms = {marks[i][0]: round(marks[i - 1][1].elapsed_time(marks[i][1]), 2) for i in range(1, len(marks))}
    coll = {k: round(sum(a.elapsed_time(b) for a, b in v) / len(v), 3) for k, v in rnd.coll_ms.items()}
    rnd.coll_ms = None
    p2, q2 = rnd.to_device()
    rnd.step(p2, q2)
    del p2, q2
    ms_e2e, _, out_e2e = timed(True, args.steps)
    finite = bool(torch.isfinite(out_e2e).all())
    if guard is not None:
        guard.cancel()

    pairs = pairs_per_step(args) * args.steps
    value = pairs / (ms_res / 1e3)
    e2e = pairs / (ms_e2e / 1e3)
    peaks = load_peaks()
    peak_tf = peaks.get('bf16_tflops_sustained', 1400.0)
    peak_gbs = peaks.get('hbm_gbs', 6400.0)
    line = {
        'metric': METRIC, 'value': round(value, 1), 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(ms_res / args.steps, 2), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': workload_config(args, world),
        'e2e': {'value': round(e2e, 1), 'unit': 'pairs/s', 'h2d_bytes_per_step': int(rnd.h2d_bytes),
                'd2h_bytes_per_step': int(out_e2e.numel() * 4), 'ms_per_step': round(ms_e2e / args.steps, 2),
                'h2d': 'pinned host -> device inside the timed region, every step; the copy of step k + 1 runs on the '
                       'prefetch stream under step k (creamfl_b200.prefetch.Prefetcher, depth 2), the first is exposed'},
        'gpu_launches': int(launches), 'clocks': clocks, 'losses_finite': finite, 'phase_ms': phase_ms,
        'public_pairs_per_s': round(S * B * args.steps / (ms_res / 1e3), 1),
    }
    if world > 1:
        # achieved bus bandwidth of the representation exchange (nccl-tests convention for all-gather:
        # (n-1)/n * total bytes / time); the con_w entries also contain the local scoring GEMM, so the all-gather
        # figure is taken from the extraction gather (pure collective) and from a stand-alone gather of the con_w payload
        slots = max(1, sum(rnd.layout_img[0]))
        payload = torch.empty((slots, N_PUB, D), dtype=torch.float32, device=dev)
        gathered = torch.empty((world,) + tuple(payload.shape), dtype=torch.float32, device=dev)
        ts = []
        for _ in range(5):
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dist.all_gather_into_tensor(gathered.view(-1), payload.view(-1))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ag_ms = sorted(ts)[len(ts) // 2]
        flat = rnd.server.model.store().grad
        ts = []
        for _ in range(5):
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ar_ms = sorted(ts)[len(ts) // 2]
        tot = gathered.numel() * 4
        line['collectives'] = {
            'conw_allgather': {'bytes_total': tot, 'ms': round(ag_ms, 3),
                               'bus_GBps': round((world - 1) / world * tot / (ag_ms * 1e-3) / 1e9, 1)},
            'grad_allreduce': {'bytes': flat.numel() * 4, 'ms': round(ar_ms, 3),
                               'bus_GBps': round(2 * (world - 1) / world * flat.numel() * 4 / (ar_ms * 1e-3) / 1e9, 1)},
            'in_step_ms': coll, 'nvlink_reference_GBps': {'allreduce_bus_measured': 725, 'p2p_per_direction': 770}}
        del payload, gathered
    if args.config == 1 and 'A' in args.phases:
        table, total, fam, overhead_us = family_table(rnd, pub, peaks)
        line['kernel_families'] = {'phase': 'one eager single-stream server train step (A), CUDA events around every '
                                            'C-ABI call, event-pair overhead subtracted',
                                   'event_overhead_us': round(overhead_us, 2), 'sum_ms': round(total, 2),
                                   'families': table}
        # roofline of the dominant-by-time tensor-core kernel family, all its launches of the step
        tc = {k: v for k, v in fam.items() if v['flops'] > 0}
        dom = max(tc, key=lambda k: tc[k]['ms'])
        ach = tc[dom]['flops'] / (tc[dom]['ms'] * 1e-3) / 1e12
        line['roofline'] = {'bound': 'tensor', 'kernel': f'{dom}_kernel (dominant kernel family of the server train '
                            f'step by time: {tc[dom]["calls"]} launches, {tc[dom]["ms"]:.2f} ms of {total:.2f} ms; '
                            f'algorithmic FLOPs of all its launches / their summed CUDA-event time)',
                            'achieved': round(ach, 1), 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': round(ach / peak_tf, 4),
                            'traffic': None,
                            'avg_launch_ms': round(tc[dom]['ms'] / tc[dom]['calls'], 4),
                            'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained' if peaks else 'fallback'}
        if dom == 'gemm_tc':          # the committed ncu --set full capture is of this family's top kernel
            cap = ROOT / 'profiles' / 'r02_ncu_dominant.csv'
            line['roofline'].update({
                'traffic': ncu_traffic_bytes(cap), 'ncu': ncu_counters(cap),
                'traffic_of': 'one launch of the family\'s top kernel by time, gemm_tc_kernel<256,1,1,0,0,1,2> (split-K '
                              'weight gradient 1024x256x25088, 65.3 MB algorithmic: both operands once + the fp32 '
                              'output), ncu --set full capture profiles/r02_ncu_dominant.csv'})
        if 'bn' in fam:
            gb = fam['bn']['bytes'] / (fam['bn']['ms'] * 1e-3) / 1e9
            line['roofline_bn'] = {'bound': 'hbm', 'kernel': 'bn_* kernels (BatchNorm statistics, apply, backward reduce / apply: '
                                   f'{fam["bn"]["calls"]} calls, {fam["bn"]["ms"]:.2f} ms of {total:.2f} ms)',
                                   'achieved': round(gb, 1), 'peak': peak_gbs, 'unit': 'GB/s', 'frac': round(gb / peak_gbs, 4),
                                   'algorithmic_bytes': int(fam['bn']['bytes']),
                                   'note': 'algorithmic bytes = forward read x + write y, backward read dy, x + write dx '
                                           '(2 B each); the kernels move 3-8 passes (statistics, residual, ReLU mask)'}
        if phase_ms.get('A_server_train'):
            step_ms = phase_ms['A_server_train'] / (S // world)
            step_tf = 63.8e9 * B / (step_ms * 1e-3) / 1e12
            line['roofline_step'] = {'bound': 'tensor', 'kernel': 'one server train step (all kernels, CUDA graph '
                                     'replay, BERT dropout on' + (', + gradient all-reduce' if world > 1 else '') + ')',
                                     'achieved': round(step_tf, 1), 'peak': peak_tf, 'unit': 'TFLOP/s',
                                     'frac': round(step_tf / peak_tf, 4), 'ms': round(step_ms, 2), 'traffic': None}
    if rank == 0 and world == 1:
        line.update(single_kernel_rooflines(rnd, peaks))
    if rank == 0 and world == 1 and not args.no_gpu_reference and args.config == 1:
        # the reference arm gets the device to itself: our engines, graph pools and prefetch rings (tens of GB) are
        # released first - with them resident its first backward passes ran 5x slower (allocator retries)
        import gc
        from creamfl_b200 import engine as _engine
        del rnd, pf, pub, priv, timed, marks, out, out_e2e
        _engine.lane._streams.clear()
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        try:
            line['gpu_reference'] = gpu_reference(args, dev)
        except Exception as e:                                    # the reference arm must not take the bench down
            line['gpu_reference'] = {'unavailable': f'{type(e).__name__}: {e}'[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_baseline(args, steps=1, warmup=0)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def single_kernel_rooflines(rnd, peaks):
    """The two stand-alone kernels of the con_w aggregation, timed live and eagerly with the L2 flushed by a 0.6 GB
    write in between: the similarity-matrix kernel (tensor-bound) and the softmax-over-clients reduce (HBM-bound)."""
    from creamfl_b200 import ops
    dev = rnd.dev
    peak_tf, peak_gbs = peaks.get('bf16_tflops_sustained', 1400.0), peaks.get('hbm_gbs', 6400.0)
    flush = rnd.server.model.store().grad
    src = next(c for c in (rnd.c_img + rnd.c_txt) if c is not None)
    c16 = ops.to_bf16(src)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        a.record()
        ops.conw_score(c16, rnd.g_txt16)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    sim_ms = sorted(ts)[len(ts) // 2]
    sim_tf = 2.0 * N_PUB * N_PUB * D / (sim_ms * 1e-3) / 1e12
    out = {'roofline_sim': {'bound': 'tensor', 'kernel': 'sim_tc_kernel<0> + lse_combine (con_w scoring, N_pub=50000, D=256)',
                            'achieved': round(sim_tf, 1), 'peak': peak_tf, 'unit': 'TFLOP/s',
                            'frac': round(sim_tf / peak_tf, 4), 'avg_launch_ms': round(sim_ms, 4),
                            'traffic': ncu_traffic_bytes(ROOT / 'profiles' / 'r01_ncu_prof_sim2.csv')}}
    vecs8 = [torch.empty_like(src).copy_(src) for _ in range(8)]
    scores8 = torch.randn(8, N_PUB, device=dev)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        a.record()
        ops.conw_reduce(vecs8, scores8)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    del vecs8
    hbm_ms = sorted(ts)[len(ts) // 2]
    hbm_bytes = 8 * N_PUB * D * 4 + 8 * N_PUB * 4 + N_PUB * D * 4
    gbs = hbm_bytes / (hbm_ms * 1e-3) / 1e9
    out['roofline_hbm'] = {'bound': 'hbm', 'kernel': 'conw_reduce_kernel (softmax over C = 8 clients + weighted sum, N_pub=50000, D=256)',
                           'achieved': round(gbs, 1), 'peak': peak_gbs, 'unit': 'GB/s', 'frac': round(gbs / peak_gbs, 4),
                           'avg_launch_ms': round(hbm_ms, 4), 'algorithmic_bytes': hbm_bytes,
                           'traffic': ncu_traffic_bytes(ROOT / 'profiles' / 'r01_ncu_hbm_kernels.csv', column=2)}
    return out


# ===================================================================================================== reference on the GPU
def gpu_reference(args, dev, steps=2):
    """the reference's path on the SAME B200 - the torch restatement of its modules
    (oracle/torch_towers.py: torchvision ResNet + HF BertModel + the reference's glue) in eager torch under bf16
    autocast (the stand-in for apex O2, which is not installable offline) with channels_last, cuDNN / cuBLAS kernels,
    AdamP restated (the adamp package is absent).  Same mini-round accounting as the main arm, measured on one
    public batch per phase and one client, scaled to S batches and C clients (every phase is a loop over identical
    batches).  This is the denominator of the north-star target '>= 10x the reference's single-GPU pairs/sec'."""
    from oracle import torch_towers as RT, creamfl_oracle as O
    S, B, C = args.public_batches, args.batch, args.clients
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    server = RT.RefPCME('resnet101', D, bert_dropout=0.1).to(dev).to(memory_format=torch.channels_last)
    client_img = RT.RefEncoderImage('resnet18', D).to(dev).to(memory_format=torch.channels_last)
    client_txt = RT.RefGRUEncoderText(VOCAB, 300, D).to(dev)
    shift = torch.nn.Parameter(torch.tensor(15.0, device=dev))
    scale = torch.nn.Parameter(torch.tensor(15.0, device=dev))
    s_params = list(server.parameters()) + [shift, scale]
    c_params = list(client_img.parameters()) + list(client_txt.parameters())
    s_opt, c_opt = O.AdamPRestated(s_params, lr=2e-4), O.AdamPRestated(c_params, lr=2e-4)
    pub = {k: v[0].to(dev) for k, v in make_public(1, B, 7, pin=False).items()}
    prv = {k: v.to(dev) for k, v in make_private('mm', B, 8, pin=False).items()}
    images = pub['images'].contiguous(memory_format=torch.channels_last)
    g_img, g_txt = make_banks(dev, 5)
    import copy
    amp = lambda: torch.autocast('cuda', dtype=torch.bfloat16)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def server_train(loss_fn):
        server.train()
        with amp():
            o = server(images, pub['ids'], pub['mask'])
        loss = loss_fn(o['image_features'].float(), o['caption_features'].float())
        s_opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(server.parameters(), 2.0)
        s_opt.step()

    def phase_a():
        server_train(lambda i, t: O.pcme_loss(i, t, shift, scale)[0])

    def phase_b():
        server.eval()
        with torch.no_grad(), amp():
            o = server(images, pub['ids'], pub['mask'])
        g_img[pub['d_idx']] = o['image_features'].float()
        g_txt[pub['d_idx']] = o['caption_features'].float()

    def client_private():
        client_img.train(); client_txt.train()
        with amp():
            zi = client_img(prv['images'].contiguous(memory_format=torch.channels_last))['embedding']
            zt = client_txt(prv['caps'], prv['cap_lens'].cpu())['embedding']
        loss, _ = O.pcme_loss(zi.float(), zt.float(), shift, scale)
        c_opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(c_params, 2.0)
        c_opt.step()

    state = {}

    def client_contrast():
        client_img.train(); client_txt.train()
        with amp():
            zi = client_img(images)['embedding']
            zt = client_txt(pub['caps'], pub['cap_lens'].cpu())['embedding']
            with torch.no_grad():
                oi, ot = state['old_i'](images)['embedding'], state['old_t'](pub['caps'], pub['cap_lens'].cpu())['embedding']
        parts = O.mm_client_contrast_loss(zi.float(), zt.float(), oi.float(), ot.float(), g_img, g_txt,
                                          pub['d_idx'], 0.5, False)
        c_opt.zero_grad(set_to_none=True)
        parts['loss'].backward()
        torch.nn.utils.clip_grad_norm_(c_params, 2.0)
        c_opt.step()

    def client_generate():
        client_img.eval(); client_txt.eval()
        with torch.no_grad(), amp():
            client_img(images)['embedding'], client_txt(pub['caps'], pub['cap_lens'].cpu())['embedding']

    def phase_f():
        server_train(lambda i, t: 0.3 * O.distill_mse(i, g_img, pub['d_idx']) + 0.3 * O.distill_mse(t, g_txt, pub['d_idx']))

    def conw():          # the reference aggregates on CPU tensors (MMFL.py:302-314); on the GPU in chunks of 4096 rows
        out = 0.0
        for gsrc, gother in ((g_img, g_txt), (g_txt, g_img)):
            rows = []
            for r0 in range(0, N_PUB, 4096):
                lg = gsrc[r0:r0 + 4096] @ gother.t()
                rows.append(lg[torch.arange(lg.shape[0]), torch.arange(r0, r0 + lg.shape[0])] - torch.logsumexp(lg, 1))
            out = out + torch.cat(rows).sum()
        return out

    def time_of(fn, n):
        for _ in range(2):                      # cuDNN autotuning and allocator growth stay outside the timed calls
            fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    state['old_i'], state['old_t'] = copy.deepcopy(client_img).eval(), copy.deepcopy(client_txt).eval()
    t = {'A': time_of(phase_a, steps), 'B': time_of(phase_b, steps), 'F': time_of(phase_f, steps),
         'private': time_of(client_private, steps), 'contrast': time_of(client_contrast, steps),
         'generate': time_of(client_generate, steps), 'conw_per_client': time_of(conw, 1)}
    round_ms = S * (t['A'] + t['B'] + t['F']) + C * (t['private'] + S * (t['contrast'] + t['generate']) + t['conw_per_client'])
    return {'value': round(pairs_per_step(args) / (round_ms * 1e-3), 1), 'unit': 'pairs/s', 'ms_per_step': round(round_ms, 1),
            'per_call_ms': {k: round(v, 2) for k, v in t.items()},
            'how': 'torch eager restatement of the reference modules (oracle/torch_towers.py) on this GPU, bf16 autocast '
                   '+ channels_last (stand-in for apex O2), cuDNN/cuBLAS, BERT dropout on, AdamP restated in torch '
                   f'(for-loop over tensors like the adamp package); each phase timed on one batch of {B} ({steps} '
                   f'calls after 2 warm-ups, our engines released first) and scaled to S={S} public batches and C={C} clients; con_w scoring chunked '
                   'on the GPU (the reference runs it on the CPU)'}


# ===================================================================================================== reference arm (CPU)
def cpu_round_sampler(b, n_conw, seed=0):
    """The reference's per-round path on the host (fp32 torch, restated in oracle/): returns closures timing each
    phase on one batch of b pairs; con_w at n_conw rows."""
    from oracle import torch_towers as RT, creamfl_oracle as O
    torch.manual_seed(seed)
    server = RT.RefPCME('resnet101', D, bert_dropout=0.1)
    client_img = RT.RefEncoderImage('resnet18', D)
    client_txt = RT.RefGRUEncoderText(VOCAB, 300, D)
    shift = torch.nn.Parameter(torch.tensor(15.0))
    scale = torch.nn.Parameter(torch.tensor(15.0))
    s_params = list(server.parameters()) + [shift, scale]
    c_params = list(client_img.parameters()) + list(client_txt.parameters())
    s_opt, c_opt = O.AdamPRestated(s_params, lr=2e-4), O.AdamPRestated(c_params, lr=2e-4)
    pub = {k: v[0] for k, v in make_public(1, b, 7, pin=False).items()}
    prv = make_private('mm', b, 8, pin=False)
    g = torch.Generator().manual_seed(5)
    g_img = unit(torch.randn(N_PUB, D, generator=g))
    g_txt = unit(0.7 * g_img + 0.5 * unit(torch.randn(N_PUB, D, generator=g)))
    cv_i = unit(g_img[:n_conw] + 0.5 * unit(torch.randn(n_conw, D, generator=g)))
    cv_t = unit(g_txt[:n_conw] + 0.5 * unit(torch.randn(n_conw, D, generator=g)))
    import copy
    old = {}
    images, ids, mask = pub['images'], pub['ids'], pub['mask']
    caps, lens, d_idx = pub['caps'], pub['cap_lens'].long(), pub['d_idx']

    def server_train(loss_fn):
        server.train()
        o = server(images, ids, mask)
        loss = loss_fn(o['image_features'], o['caption_features'])
        s_opt.zero_grad(); loss.backward()
        torch.nn.utils.clip_grad_norm_(server.parameters(), 2.0); s_opt.step()

    def A():
        server_train(lambda i, t: O.pcme_loss(i, t, shift, scale)[0])

    def Bx():
        server.eval()
        with torch.no_grad():
            o = server(images, ids, mask)
            g_img[d_idx] = o['image_features']; g_txt[d_idx] = o['caption_features']

    def P():
        old['i'], old['t'] = copy.deepcopy(client_img).eval(), copy.deepcopy(client_txt).eval()
        client_img.train(); client_txt.train()
        zi, zt = client_img(prv['images'])['embedding'], client_txt(prv['caps'], prv['cap_lens'].long())['embedding']
        loss, _ = O.pcme_loss(zi, zt, shift, scale)
        c_opt.zero_grad(); loss.backward()
        torch.nn.utils.clip_grad_norm_(c_params, 2.0); c_opt.step()

    def Cx():
        client_img.train(); client_txt.train()
        zi, zt = client_img(images)['embedding'], client_txt(caps, lens)['embedding']
        with torch.no_grad():
            oi, ot = old['i'](images)['embedding'], old['t'](caps, lens)['embedding']
        parts = O.mm_client_contrast_loss(zi, zt, oi, ot, g_img, g_txt, d_idx, 0.5, False)
        c_opt.zero_grad(); parts['loss'].backward()
        torch.nn.utils.clip_grad_norm_(c_params, 2.0); c_opt.step()

    def Dx():
        client_img.eval(); client_txt.eval()
        with torch.no_grad():
            client_img(images)['embedding'], client_txt(caps, lens)['embedding']

    def E():
        O.conw_aggregate([cv_i], g_txt[:n_conw])
        O.conw_aggregate([cv_t], g_img[:n_conw])

    def F():
        tgt = d_idx % n_conw
        server_train(lambda i, t: 0.3 * O.distill_mse(i, cv_i, tgt) + 0.3 * O.distill_mse(t, cv_t, tgt))
    return {'A': A, 'B': Bx, 'P': P, 'C': Cx, 'D': Dx, 'E': E, 'F': F}


def cpu_baseline(args, steps=1, warmup=0, b=16, n_conw=8192):
    """pairs/s of the host path on the main arm's accounting.  Each sample runs every phase once on one batch of b
    pairs (server A, B, F; one client's private / contrast / generate; con_w of one client at n_conw rows); per-batch
    phases scale linearly in pairs, con_w quadratically in rows; the mini-round is then assembled exactly like the
    GPU arm's: S * B * (A + B + F) + C * B * (P + S * (C + D)) + C * E(N_pub)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    S, B, C = args.public_batches, args.batch, args.clients
    fns = cpu_round_sampler(b, n_conw)
    tot = {k: 0.0 for k in fns}
    t_all = time.perf_counter()
    for it in range(warmup + steps):
        for k, fn in fns.items():
            t0 = time.perf_counter()
            fn()
            if it >= warmup:
                tot[k] += time.perf_counter() - t0
    wall = time.perf_counter() - t_all
    per = {k: v / steps for k, v in tot.items()}
    scale_pairs = B / b
    round_s = S * scale_pairs * (per['A'] + per['B'] + per['F']) + \
        C * scale_pairs * (per['P'] + S * (per['C'] + per['D'])) + C * per['E'] * (N_PUB / n_conw) ** 2
    return {'value': round(pairs_per_step(args) / round_s, 3), 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
            'sample': f'{steps} sample(s) of the torch-fp32 restatement (oracle/): every phase once on one batch of {b} '
                      f'pairs (BERT dropout on, AdamP restated): ' +
                      ', '.join(f'{k} {v:.2f}s' for k, v in per.items()) +
                      f'; con_w (E) at {n_conw} rows scaled x{(N_PUB / n_conw) ** 2:.1f} to N_pub={N_PUB}; assembled to the '
                      f'mini-round of S={S} x {B} public pairs and C={C} clients ({round_s:.0f} s per mini-round on {cores} cores)',
            'wall_s': round(wall, 1), 'seconds_per_step_extrapolated': round(round_s, 1)}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    t0 = time.perf_counter()
    base = cpu_baseline(args, steps=max(1, args.steps), warmup=min(1, args.warmup))
    wall = time.perf_counter() - t0
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'pairs/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': round(1e3 * wall / max(1, args.steps + min(1, args.warmup)), 1), 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, max(1, args.gpus)),
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


if __name__ == '__main__':
    a = parse()
    if os.environ.get('CREAMFL_BENCH_WATCHDOG'):
        # development aid: dump every Python thread's stack and exit if the run has not finished after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ['CREAMFL_BENCH_WATCHDOG']), exit=True)
    if a.impl == 'reference':
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device - the creamfl_b200 hot path has no CPU fallback '
                             '(use --impl reference for the host baseline)')
        run_ours(a)
