"""CPU tests: the product's index producers (creamfl_b200/partition.py) are bit-identical to the reference's
(golden vectors in tests/golden/partition.npz were produced by the reference's own functions)."""
import numpy as np
import pytest
import torch

from creamfl_b200 import partition as P
from oracle import creamfl_oracle as O


@pytest.mark.parametrize('tag', ['synth', 'cifar', 'agnews'])
def test_hetero_partition_bit_exact(golden, tag):
    g = golden('partition')
    n, k, nets, seed = (int(v) for v in g[f'{tag}_args'])
    name = {'synth': 'synth', 'cifar': 'cifar100', 'agnews': 'AG_NEWS'}[tag]
    part = P.data_partitioner(name, n, nets, 'hetero', float(g[f'{tag}_alpha']), np.arange(n) % k, seed=seed)
    assert [len(part[j]) for j in range(nets)] == list(g[f'{tag}_sizes'])
    assert O.partition_digest(part) == str(g[f'{tag}_sha256'])


def test_flickr_shards_bit_exact(golden):
    g = golden('partition')
    part = P.shard_partition(145000, 15, 150, seed=2021)
    assert O.partition_digest(part) == str(g['f30k_sha256'])


def test_flickr_shards_reproduce_the_fixture_the_reference_ships(golden):
    """/root/reference/data_partition/client_noniid_flicker30k.pkl is the one index fixture of the reference that does
    not depend on a dataset (flickr30k.py:79-102 needs only len(data) = 145 000 and numpy's legacy RNG at seed 2021):
    its sha256 - taken over the pickle's own arrays by tests/golden/make_golden.py - must come out of the product's
    and the oracle's shard partition.  (The CIFAR-100 / AG_NEWS fixtures were drawn from the real datasets' label
    vectors, which are not on the box; their hashes are recorded in the golden file for a maintainer who has them.)"""
    g = golden('partition')
    shipped = str(g['fixture_client_noniid_flicker30k_sha256'])
    assert shipped.startswith('fcd3455d84de4a2b')
    assert list(g['fixture_client_noniid_flicker30k_sizes']) == [9660] * 14 + [9760]
    assert O.partition_digest(P.shard_partition(145000, 15, 150, seed=2021)) == shipped
    assert O.partition_digest(O.shard_partition(145000, 15, 150, seed=2021)) == shipped
    for name, n in (('client_cifar100_noniid', 50000), ('client_AG_NEWS_noniid', 120000)):
        assert int(g[f'fixture_{name}_sizes'].sum()) == n and len(g[f'fixture_{name}_sizes']) == 10


def test_public_subset_is_the_coco_subset_idx_file_the_reference_ships(golden):
    """/root/reference/coco_subset_idx_file (50 000 sorted caption indices, load_datasets.py:148-157): sha256 of its
    int64 bytes recorded by tests/golden/make_golden.py; product and oracle reproduce the file bit for bit."""
    import hashlib
    g = golden('partition')
    want = str(g['fixture_coco_subset_sha256'])
    assert want.startswith('8ffcd8243cd86772') and int(g['fixture_coco_subset_len']) == 50000
    sha = lambda idx: hashlib.sha256(np.asarray(idx, dtype=np.int64).tobytes()).hexdigest()
    mine = P.public_subset_indices()
    assert len(mine) == 50000 and mine == sorted(mine) and mine[:8] == [9, 20, 33, 46, 76, 90, 97, 99]
    assert sha(mine) == want
    assert sha(O.public_subset_indices()) == want
    small = P.public_subset_indices(256)
    assert len(set(small)) == 256 and small == sorted(small) and max(small) < P.COCO_TRAIN_CAPTIONS
    with pytest.raises(ValueError):
        P.public_subset_indices(0)


def test_distill_lookup_matches_dict_semantics():
    rng = np.random.default_rng(0)
    distill_index = rng.permutation(5000)[:777].tolist()
    d = {b: a for a, b in enumerate(distill_index)}                      # MMFL.py:343
    lut = P.distill_lookup(distill_index)
    batch = [distill_index[i] for i in (5, 700, 0, 42)]
    assert lut[torch.tensor(batch)].tolist() == [d[b] for b in batch]
    assert lut[torch.tensor(batch[:1])].tolist() == [d[batch[0]]]        # batch of one (itemgetter returns a bare int)
