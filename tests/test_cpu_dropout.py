"""BERT dropout (reference: HF BertConfig defaults p = 0.1 under model.train(), pcme.py:31 + retrieval_trainer.py:187)
checked WITHOUT a GPU:

  * the numpy Philox4x32-10 of tests/kernel_emulation.py against the Random123 known-answer vectors;
  * csrc/philox.cuh compiled for the HOST by nvcc against that numpy restatement (same header the kernels include);
  * the host sequencing of the BERT tower with dropout on (towers._BertFn on the emulated C ABI) against the HF
    BertModel oracle fed the identical keep masks through oracle.torch_towers.frozen_dropout: forward values and every
    parameter gradient, exact mode;
  * keep-rate statistics and the eval-mode no-op.
The CUDA kernels themselves are compared with the same oracle on the GPU (tests/test_gpu_dropout.py)."""
import copy
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

import kernel_emulation as KE
from test_cpu_tower_host import _inputs, _pcme_pair, _rel

ROOT = Path(__file__).resolve().parent.parent


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 with 10 rounds."""
    kat = [((0, 0, 0, 0), (0, 0), '6627e8d5 e169c58d bc57ac4c 9b00dbd8'),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), '408f276d 41c83b0e a20bc7c6 6d5451fd'),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            'd16cfe09 94fdcceb 5001e420 24126ea1')]
    for ctr, key, want in kat:
        r = KE.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert ' '.join('%08x' % int(x[0]) for x in r) == want


def test_philox_header_matches_numpy(tmp_path):
    """The header the CUDA kernels include, compiled as host code: keep masks equal the numpy restatement."""
    src = tmp_path / 'k.cu'
    src.write_text('''
#include <cstdio>
#include <cstdlib>
#include "philox.cuh"
int main(int argc, char** argv) {
  unsigned long long seed = strtoull(argv[1], 0, 10);
  unsigned step = (unsigned)strtoul(argv[2], 0, 10), site = (unsigned)strtoul(argv[3], 0, 10);
  long long n = atoll(argv[4]);
  unsigned thresh = cfl::drop_thresh16((float)atof(argv[5]));
  for (long long b = 0; b < (n + 7) / 8; ++b) {
    unsigned k = cfl::drop_keep8(seed, step, site, (unsigned long long)b, thresh);
    for (int i = 0; i < 8 && b * 8 + i < n; ++i) putchar('0' + ((k >> i) & 1));
  }
  return 0;
}''')
    exe = tmp_path / 'k'
    r = subprocess.run(['nvcc', '-O1', '-I', str(ROOT / 'creamfl_b200' / 'csrc'), str(src), '-o', str(exe)],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip('nvcc host build unavailable: ' + r.stderr[-300:])
    for seed, step, site, n, p in [(1234, 1, 0, 1000, 0.1), (2 ** 40 + 17, 70000, 36, 4099, 0.1), (5, 3, 7, 257, 0.5)]:
        out = subprocess.run([str(exe), str(seed), str(step), str(site), str(n), str(p)], capture_output=True,
                             text=True, check=True).stdout
        got = np.frombuffer(out.encode(), dtype=np.uint8) - ord('0')
        assert np.array_equal(got, KE.keep_mask_np(seed, step, site, n, p)), (seed, step, site)


def test_keep_rate_and_site_independence():
    n = 1 << 20
    m0 = KE.keep_mask_np(99, 1, 0, n, 0.1)
    m1 = KE.keep_mask_np(99, 1, 1, n, 0.1)
    m2 = KE.keep_mask_np(99, 2, 0, n, 0.1)
    for m in (m0, m1, m2):
        assert abs(m.mean() - (1 - 6554 / 65536)) < 4 * (0.09 / n) ** 0.5 + 1e-3       # ~4 sigma of a Bernoulli(0.9)
    # different sites / steps are different masks that agree at the chance rate 0.9^2 + 0.1^2 = 0.82
    assert abs((m0 == m1).mean() - 0.82) < 5e-3 and abs((m0 == m2).mean() - 0.82) < 5e-3


def _provider(mine, batch, seq, layers):
    """Keep masks of the product's dropout sites in the shapes HF asks for (see frozen_dropout)."""
    bert = mine.txt_enc
    seed, p = bert._seed & 0x7fffffffffffffff, bert.dropout_p

    def provider(call, shape):
        site = call
        last_dense = site in (2 + 3 * (layers - 1), 3 + 3 * (layers - 1))
        if last_dense:
            # the product runs the last layer's row-wise ops on the [CLS] rows only (element index = b * 768 + col);
            # every other token row of the last layer is dead (pcme.py:44 reads last_hidden_state[:, 0])
            full = torch.ones(shape)
            small = torch.from_numpy(KE.keep_mask_np(seed, 1, site, batch * shape[-1], p)).view(batch, shape[-1])
            full[:, 0, :] = small.float()
            return full
        n = int(np.prod(shape))
        return torch.from_numpy(KE.keep_mask_np(seed, 1, site, n, p)).view(*shape).float()
    return provider


def test_bert_tower_dropout_matches_oracle_with_same_masks(monkeypatch):
    from oracle import torch_towers as RT
    KE.install(monkeypatch, exact=True)
    layers, batch, seq = 2, 3, 8
    ref, mine = _pcme_pair(layers=layers, seed=11, dropout=0.1)
    images, ids, mask, cot_i, cot_t = _inputs(batch=batch, seq=seq, seed=12)
    ref64 = copy.deepcopy(ref).double()
    with RT.frozen_dropout(_provider(mine, batch, seq, layers)) as fd:
        o64 = ref64(images.double(), ids, mask, torch.zeros_like(ids))
        assert fd.calls == 1 + 3 * layers
    ((o64['image_features'] * cot_i.double()).sum() + (o64['caption_features'] * cot_t.double()).sum()).backward()
    with RT.frozen_dropout(_provider(mine, batch, seq, layers)):
        o32 = ref(images, ids, mask, torch.zeros_like(ids))
    ((o32['image_features'] * cot_i).sum() + (o32['caption_features'] * cot_t).sum()).backward()
    mine.store().zero_grad()
    o = mine(images, None, {'input_ids': ids, 'attention_mask': mask}, None)
    assert int(mine.txt_enc.dropout_state(ids.device).rng[1]) == 1          # one tick per training forward
    assert _rel(o['caption_features'], o64['caption_features']) < 1e-4
    ((o['image_features'] * cot_i).sum() + (o['caption_features'] * cot_t).sum()).backward()
    # text-tower parameters (the image tower has no dropout; its ill-conditioned 3-image BatchNorm gradients are the
    # subject of test_cpu_tower_host.py)
    p64, p32 = dict(ref64.named_parameters()), dict(ref.named_parameters())
    checked = 0
    for name, prm in mine.named_parameters():
        if not (name.startswith('txt_enc.') or name.startswith('linear.')) or name.endswith('attention.self.key.bias'):
            continue
        g64 = p64[name].grad
        if g64 is None or float(g64.abs().max()) == 0.0:
            assert float(prm.grad.abs().max()) == 0.0, name
            continue
        e_mine, e_torch = _rel(prm.grad, g64), _rel(p32[name].grad, g64)
        assert e_mine <= 2.0 * e_torch + 2e-4, (name, e_mine, e_torch)
        checked += 1
    assert checked >= 30
    # and it IS a different function from the dropout-free tower
    ref0, _ = _pcme_pair(layers=layers, seed=11, dropout=0.0)
    with torch.no_grad():
        o0 = ref0(images, ids, mask, torch.zeros_like(ids))
    assert _rel(o['caption_features'], o0['caption_features']) > 1e-2


def test_dropout_fresh_mask_every_step_and_off_in_eval(monkeypatch):
    KE.install(monkeypatch, exact=True)
    _, mine = _pcme_pair(layers=2, seed=13, dropout=0.1)
    images, ids, mask, _, _ = _inputs(seed=14)
    tok = {'input_ids': ids, 'attention_mask': mask}
    with torch.no_grad():
        a = mine(images, None, tok, None)['caption_features']
        b = mine(images, None, tok, None)['caption_features']
        assert _rel(a, b) > 1e-3                                            # step 1 vs step 2: new masks
        mine.eval()
        c = mine(images, None, tok, None)['caption_features']
        d = mine(images, None, tok, None)['caption_features']
        assert torch.equal(c, d)
        assert int(mine.txt_enc.dropout_state(ids.device).rng[1]) == 2      # eval forwards do not tick
        _, mine0 = _pcme_pair(layers=2, seed=13, dropout=0.0)
        mine0.eval()
        assert torch.equal(c, mine0(images, None, tok, None)['caption_features'])
