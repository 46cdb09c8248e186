"""Synthetic COCO / Flickr30k / CIFAR-100 / AG_NEWS-shape loaders.  No dataset is available on the
box; batches have the reference's collate layout (src/datasets/_dataloader.py:35-64):
    (images, targets, captions, cap_lengths, ann_ids, image_ids, index)
with `captions` being the BERT token dict the CUDA PCME consumes instead of raw strings (pcme.py:40 tokenises strings
on the host; the tokenizer vocabulary is not on the box either)."""
from __future__ import annotations

import torch

VOCAB = 11755
BERT_L = 32


class SyntheticPairs:
    """Deterministic image/caption pairs addressed by dataset index; `indices` plays the role of the public-subset
    index file (src/utils/load_datasets.py:148-162)."""

    def __init__(self, indices, batch_size, image_size=224, cap_len=30, seed=0, shuffle=False, drop_last=False):
        self.indices = list(indices)
        self.batch_size, self.image_size, self.cap_len = batch_size, image_size, cap_len
        self.seed, self.shuffle, self.drop_last = seed, shuffle, drop_last
        self.epoch = 0
        self.n_images = len(self.indices)

    def __len__(self):
        n = len(self.indices)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = list(range(len(self.indices)))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + 7919 * self.epoch)
            order = torch.randperm(len(order), generator=g).tolist()
            self.epoch += 1
        for b in range(len(self)):
            rows = order[b * self.batch_size:(b + 1) * self.batch_size]
            yield self.batch([self.indices[r] for r in rows])

    def batch(self, index):
        n = len(index)
        g = torch.Generator().manual_seed(self.seed * 1000003 + int(index[0]))
        images = torch.randn(n, 3, self.image_size, self.image_size, generator=g)
        lens = torch.sort(torch.randint(5, self.cap_len + 1, (n,), generator=g), descending=True).values
        lens[0] = self.cap_len                    # per-batch length-descending sort of the collate fn (:49)
        cmask = (torch.arange(self.cap_len)[None] < lens[:, None]).long()
        targets = torch.randint(4, VOCAB, (n, self.cap_len), generator=g) * cmask
        blens = torch.clamp(lens + 2, max=BERT_L)
        bmask = (torch.arange(BERT_L)[None] < blens[:, None]).long()
        ids = torch.randint(1000, 30522, (n, BERT_L), generator=g)
        ids[:, 0] = 101
        ids.scatter_(1, (blens - 1).unsqueeze(1), 102)
        captions = {'input_ids': ids * bmask, 'attention_mask': bmask}
        return images, targets, captions, lens, list(index), list(index), list(index)


class SyntheticLabelled:
    """CIFAR-100-shape (images, labels) or AG_NEWS-shape (token ids, labels, lengths) private client batches."""

    def __init__(self, kind, indices, batch_size, num_class, image_size=64, seq_len=60, seed=0):
        self.kind, self.indices, self.batch_size, self.num_class = kind, list(indices), batch_size, num_class
        self.image_size, self.seq_len, self.seed = image_size, seq_len, seed

    def __len__(self):
        return (len(self.indices) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        for b in range(len(self)):
            idx = self.indices[b * self.batch_size:(b + 1) * self.batch_size]
            g = torch.Generator().manual_seed(self.seed * 7 + int(idx[0]))
            labels = torch.tensor([i % self.num_class for i in idx])
            if self.kind == 'image':
                yield torch.randn(len(idx), 3, self.image_size, self.image_size, generator=g), labels
            else:
                lens = torch.sort(torch.randint(10, self.seq_len + 1, (len(idx),), generator=g), descending=True).values
                lens[0] = self.seq_len
                mask = (torch.arange(self.seq_len)[None] < lens[:, None]).long()
                yield torch.randint(4, VOCAB, (len(idx), self.seq_len), generator=g) * mask, labels, lens
