"""Worker of tests/test_gpu_graph_parity.py::test_data_parallel_server_step...: launched by torchrun on 2 GPUs."""
import os
import sys

import torch
import torch.distributed as dist


def main(out_path):
    rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    from creamfl_b200 import engine

    def inputs(seed, B=8, L=16):
        g = torch.Generator().manual_seed(seed)
        images = torch.randn(B, 3, 224, 224, generator=g).to(dev)
        ids = torch.randint(1000, 30522, (B, L), generator=g)
        ids[:, 0] = 101
        return images, {'input_ids': ids.to(dev), 'attention_mask': torch.ones(B, L, dtype=torch.long, device=dev)}

    def make(dp, graphs):
        torch.manual_seed(3)
        return engine.ServerEngine(64, 'resnet18', device=dev, data_parallel=dp, use_graphs=graphs, bert_dropout=0.0)

    def single():
        """one process, mean of the two ranks' gradients, one optimizer step"""
        s = make(False, False)
        acc = torch.zeros_like(s.model.store().grad)
        cacc = [torch.zeros_like(p) for p in s.criterion.parameters()]
        for r in range(2):
            s._train_fwd_bwd(*_split(inputs(50 + r)))
            acc += s.model.store().grad
            for a, p in zip(cacc, s.criterion.parameters()):
                a += p.grad
        s.model.store().grad.copy_(acc / 2)
        for a, p in zip(cacc, s.criterion.parameters()):
            p.grad.copy_(a / 2)
        s.optimizer.step()
        return s.model.store().flat.clone()

    def _split(pair):
        images, tok = pair
        return images, {'ids': tok['input_ids'], 'mask': tok['attention_mask']}

    ref, ref2 = single(), single()
    res = {'single_vs_single': float((ref - ref2).abs().max())}
    for name, graphs in (('dp_vs_single', False), ('graph_dp_vs_single', True)):
        s = make(True, graphs)
        s.train_step(*inputs(50 + rank))
        flat = s.model.store().flat
        res[name] = float((flat - ref).abs().max())
        other = flat.clone()
        dist.broadcast(other, 0)
        res['ranks_equal'] = max(res.get('ranks_equal', 0.0), float((flat - other).abs().max()))
    vals = torch.tensor([res[k] for k in sorted(res)], device=dev)
    dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save({k: float(v) for k, v in zip(sorted(res), vals)}, out_path)
    dist.destroy_process_group()


if __name__ == '__main__':
    main(sys.argv[1])
