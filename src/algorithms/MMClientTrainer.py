"""Mirror of the reference's src/algorithms/MMClientTrainer.py (run :91-114, train_epoch :116-324,
generate_logits :326-359) on creamfl_b200.engine.MMClient."""
from __future__ import annotations

import torch

from creamfl_b200 import ops
from creamfl_b200.engine import MMClient
from creamfl_b200.partition import distill_lookup


class MMClientTrainer:
    def __init__(self, args, train_loader, client=-1, device='cuda', logger=None):
        self.args, self.train_loader, self.client, self.logger = args, train_loader, client, logger
        self.client_idx = client
        self.local_epochs, self.local_epoch, self.cur_epoch = args.local_epochs, 0, 0
        self._core = MMClient(embed_dim=args.feature_dim, interintra_weight=args.interintra_weight,
                              use_graphs=not getattr(args, 'no_cuda_graphs', True))
        self.model, self.criterion, self.optimizer = self._core.model, self._core.criterion, self._core.optimizer
        self.device = self._core.device

    def run(self, global_img_feature, global_txt_feature, distill_index, global_train_loader, prefix=''):
        self._core.begin_round()
        for _ in range(self.local_epochs):
            self.local_epoch += 1
            if self.logger is not None:
                self.logger.log(f'Epoch {self.local_epoch}')
            self.train_epoch(global_img_feature, global_txt_feature, distill_index, global_train_loader)

    def train_epoch(self, global_img_feature, global_txt_feature, distill_index, global_train_loader, prefix=''):
        dev = self.device
        for images, captions, captions_word, caption_lens, _, _, index in self.train_loader:
            self._core.private_step(images.to(dev, non_blocking=True), captions.to(dev, non_blocking=True), caption_lens)
        intra, inter = self.args.contrast_local_intra, self.args.contrast_local_inter
        if not (intra or inter):
            return
        g_img, g_txt = global_img_feature.to(dev).float(), global_txt_feature.to(dev).float()
        g_img16 = g_txt16 = None
        if not self._core.use_graphs:          # (the graphed client keeps persistent fp32 + bf16 copies itself)
            g_img16, g_txt16 = ops.to_bf16(g_img), ops.to_bf16(g_txt)
        lut = distill_lookup(distill_index, dev)
        for images, captions, captions_word, caption_lens, _, _, index in global_train_loader:
            d_idx = lut[torch.as_tensor(index, device=dev)]
            self._core.contrast_step(images.to(dev, non_blocking=True), captions.to(dev, non_blocking=True),
                                     caption_lens, d_idx, g_img, g_txt, g_img16, g_txt16, intra=intra, inter=inter,
                                     loss_scale=self.args.loss_scale)

    def generate_logits(self, dataloader):
        dev = self.device
        img, txt, distill_index = [], [], []
        for images, captions, captions_word, caption_lens, _, _, index in dataloader:
            fi, ft = self._core.generate(images.to(dev, non_blocking=True), captions.to(dev, non_blocking=True),
                                         caption_lens)
            img.append(fi.clone())
            txt.append(ft.clone())
            distill_index.extend(index)
        return {'img': torch.cat(img), 'txt': torch.cat(txt)}, distill_index
