"""Mirror of the reference's src/algorithms/ClientTrainer.py for the unimodal clients (run :193-215, lr_scheduler
:291-302, tra :307-510, generate_logits / extract_pub_feature :622-664) on creamfl_b200.engine.UnimodalClient:
supervised pass with the margin / class-centre losses, then the inter + intra contrast pass against the server's
public features, SGD(lr 1e-4, momentum 0.9, wd 5e-5) with the reference's two-step decay (x0.1 from round 15, x0.01
from round 24 of num_epochs = 30)."""
from __future__ import annotations

import torch

from creamfl_b200.engine import UnimodalClient
from creamfl_b200.partition import distill_lookup

try:                                                    # imported as the package src.algorithms
    from .. import losses
except (ImportError, ValueError):                       # `python src/main.py`: src/ itself is on the path
    import losses


def seed_torch(seed=2021):
    """ClientTrainer.py:35-41."""
    import os
    import random
    import numpy as np
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class ClientTrainer:
    def __init__(self, args, dataset, dst, RGBmean, RGBstdv, data_dict, logger, global_test_set=None, inter_distance=4,
                 loss='softmax', gpuid='cuda:0', num_epochs=30, init_lr=0.0001, decay=0.1, batch_size=512, imgsize=256,
                 num_workers=4, print_freq=10, save_step=10, scale=128, pool_type='max_avg', client_id=-1, wandb=None):
        # ClientTrainer.py:140 - every unimodal client re-seeds ALL global generators to 2021 in its constructor: clients
        # of one type start from identical weights, and the per-round `random.sample` of MMFL.py:191 no longer depends
        # on --seed once a unimodal client exists; kept, because it decides who trains when
        seed_torch()
        self.args, self.dset_name, self.logger, self.client_idx = args, dataset, logger, client_id
        self.train_loader = data_dict
        self.device = torch.device(gpuid)
        self.inter_distance, self.scale, self.init_lr = inter_distance, scale, init_lr
        self.num_epochs, self.decay_rate, self.decay_time = num_epochs, decay, [False, False]
        self.local_epochs, self.local_epoch, self.cur_epoch = args.local_epochs, 0, 0
        self.classSize = {'Cifar100': 100, 'Cifar10': 10, 'AG_NEWS': 4, 'YelpReviewPolarity': 2}[dataset]
        self.is_image = dataset in ('Cifar100', 'Cifar10')
        self.loss = loss
        self.setModel()

    def setModel(self):
        """ClientTrainer.py:274-289."""
        self._core = UnimodalClient('image' if self.is_image else 'text', self.classSize,
                                    embed_dim=self.args.feature_dim, lr=self.init_lr, scale=self.scale,
                                    interintra_weight=self.args.interintra_weight, inter_distance=self.inter_distance,
                                    device=self.device, use_graphs=not getattr(self.args, 'no_cuda_graphs', True))
        self.model, self.optimizer = self._core.model, self._core.optimizer
        self.criterion = losses.create(self.loss)           # :280,284; the engine's steps call the same kernel fused

    def lr_scheduler(self, epoch):
        """ClientTrainer.py:291-302 (FusedOptimizer reads param_groups['lr'] before every step / graph replay)."""
        if epoch >= 0.5 * self.num_epochs and not self.decay_time[0]:
            self.decay_time[0] = True
            for group in self._core.optimizer.param_groups:
                group['lr'] = self.init_lr * self.decay_rate
        if epoch >= 0.8 * self.num_epochs and not self.decay_time[1]:
            self.decay_time[1] = True
            for group in self._core.optimizer.param_groups:
                group['lr'] = self.init_lr * self.decay_rate * self.decay_rate

    def run(self, global_img_feature, global_txt_feature, distill_index, global_train_loader):
        self._core.begin_round()                                           # old_model = deepcopy(model).eval()
        self.lr_scheduler(self.cur_epoch)
        for _ in range(self.local_epochs):
            self.local_epoch += 1
            self.tra(global_img_feature, global_txt_feature, distill_index, global_train_loader)

    def tra(self, global_img_feature, global_txt_feature, distill_index, global_train_loader):
        dev, core = self.device, self._core
        for data in self.train_loader:                                     # supervised pass (:322-363)
            if self.is_image:
                inputs, labels = data
                core.supervised_step(inputs.to(dev, non_blocking=True), labels.to(dev, non_blocking=True))
            else:
                inputs, labels, caplens = data
                core.supervised_step(inputs.to(dev, non_blocking=True), labels.to(dev, non_blocking=True), caplens)
        intra, inter = self.args.contrast_local_intra, self.args.contrast_local_inter
        if not (intra or inter):
            return
        g_img, g_txt = global_img_feature.to(dev).float(), global_txt_feature.to(dev).float()
        g_same, g_other = (g_img, g_txt) if self.is_image else (g_txt, g_img)
        lut = distill_lookup(distill_index, dev)
        for images, captions, captions_word, caption_lens, _, _, index in global_train_loader:
            d_idx = lut[torch.as_tensor(index, device=dev)]
            x = (images if self.is_image else captions).to(dev, non_blocking=True)
            core.contrast_step(x, caption_lens, d_idx, g_same, g_other, intra=intra, inter=inter,
                               loss_scale=self.args.loss_scale)

    def generate_logits(self, dataloader):
        vec, idx = self.extract_pub_feature(dataloader)
        return ({'img': vec, 'txt': None} if self.is_image else {'img': None, 'txt': vec}), idx

    def extract_pub_feature(self, dataloader):
        """ClientTrainer.py:631-664 - note: no .eval() in the reference, BatchNorm keeps using batch statistics."""
        dev = self.device
        feats, distill_index = [], []
        for images, captions, captions_word, caption_lens, _, _, index in dataloader:
            x = (images if self.is_image else captions).to(dev, non_blocking=True)
            feats.append(self._core.generate(x, caption_lens).clone())
            distill_index.extend(index)
        return torch.cat(feats), distill_index
