"""Prefetcher (creamfl_b200/prefetch.py): copies issued on the side stream arrive intact, ring slots are not
overwritten before the consuming step released them, pageable and pinned sources both work."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_prefetcher_ring_order_and_values():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    from creamfl_b200.prefetch import Prefetcher
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(0)
    host = [{'images': torch.randn(4, 3, 64, 64, generator=g).pin_memory(),
             'ids': torch.randint(0, 1000, (4, 16), generator=g),                  # pageable on purpose
             'nested': [torch.full((8,), float(k)), (torch.arange(5) + k,)]} for k in range(5)]
    pf = Prefetcher(dev, depth=2)
    pf.submit(host[0])
    with pytest.raises(RuntimeError):
        pf.submit(host[1]); pf.submit(host[2])               # third in flight: refused
    seen = []
    pf2 = Prefetcher(dev, depth=2)
    pf2.submit(host[0])
    for k in range(5):
        b = pf2.next()
        if k + 1 < 5:
            pf2.submit(host[k + 1])
        torch.cuda._sleep(2_000_000)                          # the "step": keeps the stream busy while the next copy runs
        seen.append({'images': b['images'].clone(), 'ids': b['ids'].clone(), 'f': b['nested'][0].clone(),
                     'a': b['nested'][1][0].clone()})
        pf2.release()
    torch.cuda.synchronize()
    for k in range(5):
        assert torch.equal(seen[k]['images'].cpu(), host[k]['images'])
        assert torch.equal(seen[k]['ids'].cpu(), host[k]['ids'])
        assert torch.equal(seen[k]['f'].cpu(), host[k]['nested'][0])
        assert torch.equal(seen[k]['a'].cpu(), host[k]['nested'][1][0])
    assert pf2.bytes_copied == sum(sum(t.numel() * t.element_size() for t in (h['images'], h['ids'], h['nested'][0],
                                                                               h['nested'][1][0])) for h in host)


def test_prefetcher_refuses_cpu_device():
    from creamfl_b200.prefetch import Prefetcher
    with pytest.raises(RuntimeError):
        Prefetcher(torch.device('cpu'))
