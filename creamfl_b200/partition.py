"""Client-partition index producers of the hot path's data side (host integer work, numpy legacy global RNG exactly
like the reference, so the index tensors are bit-identical):

  data_partitioner   - src/datasets/load_FL_datasets.py:79-122 ("hetero" Dirichlet, "homo"), without the pickle cache
  shard_partition    - src/datasets/flickr30k.py:79-102 (150 shards, 10 per client, leftovers to the last client)
  distill_lookup     - {dataset index -> row} of MMFL.py:343 / MMClientTrainer.py:152 as an int64 lookup table
  public_subset_indices - src/utils/load_datasets.py:148-157: the sorted public subset of the 566 435 COCO training
                       captions (Python's `random` at seed 2021 reproduces the `coco_subset_idx_file` the reference ships)
"""
from __future__ import annotations

import random
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

_MIN_SIZE = {'cifar100': 10, 'AG_NEWS': 3000}


def data_partitioner(dataset: str, num_samples: int, num_nets: int, partition: str = 'hetero', alpha: float = 0.1,
                     y_train: Optional[np.ndarray] = None, seed: Optional[int] = None,
                     min_size: Optional[int] = None) -> Dict[int, List[int]]:
    if seed is not None:
        np.random.seed(seed)
    if partition == 'homo':
        idxs = np.random.permutation(num_samples)
        return {i: part.tolist() for i, part in enumerate(np.array_split(idxs, num_nets))}
    if partition != 'hetero':
        raise ValueError(f'unknown partition {partition}')
    if y_train is None:
        raise ValueError('hetero partition needs the label vector')
    classes = int(np.max(y_train)) + 1
    floor = _MIN_SIZE.get(dataset, 500) if min_size is None else min_size   # reference thresholds (:98-101)
    smallest = 0
    buckets: List[List[int]] = []
    while smallest < floor:
        buckets = [[] for _ in range(num_nets)]
        for k in range(classes):
            members = np.where(y_train == k)[0]
            np.random.shuffle(members)
            share = np.random.dirichlet(np.repeat(alpha, num_nets))
            # clients that already hold their fair share get nothing more of this class
            share = np.array([s * (len(b) < num_samples / num_nets) for s, b in zip(share, buckets)])
            share = share / share.sum()
            cuts = (np.cumsum(share) * len(members)).astype(int)[:-1]
            buckets = [b + piece.tolist() for b, piece in zip(buckets, np.split(members, cuts))]
            smallest = min(len(b) for b in buckets)
    out = {}
    for j in range(num_nets):
        np.random.shuffle(buckets[j])
        out[j] = buckets[j]
    return out


def shard_partition(n_items: int, num_users: int = 15, num_shards: int = 150,
                    seed: Optional[int] = None) -> Dict[int, np.ndarray]:
    if seed is not None:
        np.random.seed(seed)
    per_shard = int(n_items / num_shards)
    free = list(range(num_shards))
    users = {i: np.array([], dtype=int) for i in range(num_users)}
    flat = np.arange(num_shards * per_shard)
    left = list(range(n_items))
    last = 0
    for last in range(num_users):
        picked = set(np.random.choice(free, int(num_shards / num_users), replace=False))
        free = list(set(free) - picked)
        for sh in picked:                                # CPython set order, as in the reference
            block = flat[sh * per_shard:(sh + 1) * per_shard]
            users[last] = np.concatenate((users[last], block), axis=0)
            left = list(set(left) - set(block))
    users[last] = np.concatenate([users[last], left])
    return users


COCO_TRAIN_CAPTIONS = 566435          # len(CocoCaptionsCap) of the training split (load_datasets.py:150)


def public_subset_indices(subset_num: int = 50000, n_total: int = COCO_TRAIN_CAPTIONS, seed: int = 2021) -> List[int]:
    """load_datasets.py:148-157: shuffle range(n_total) with Python's Mersenne Twister, keep the first `subset_num`,
    sort.  With the defaults this IS the `coco_subset_idx_file` of the reference repository (sha256 of the int64
    bytes 8ffcd824...; tests/test_cpu_partition.py) - the reference draws it once with whatever state `random` is in
    and pickles it; the shipped file corresponds to seed 2021.  A private generator is used, the global `random`
    state is left alone."""
    if not 0 < subset_num <= n_total:
        raise ValueError(f'public_subset_indices: subset_num {subset_num} outside (0, {n_total}]')
    full = list(range(n_total))
    random.Random(seed).shuffle(full)
    return sorted(full[:subset_num])


def distill_lookup(distill_index: Sequence[int], device=None) -> torch.Tensor:
    """int64 table `lut[dataset_index] = row in the public bank` (-1 elsewhere): `lut[index]` replaces
    `operator.itemgetter(*index)(distill_dict)` (MMClientTrainer.py:156) and works for batches of one."""
    idx = torch.as_tensor(list(distill_index), dtype=torch.long)
    lut = torch.full((int(idx.max()) + 1,), -1, dtype=torch.long)
    lut[idx] = torch.arange(len(idx))
    return lut.to(device) if device is not None else lut
