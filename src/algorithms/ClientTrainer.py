"""Mirror of the reference's src/algorithms/ClientTrainer.py for the unimodal clients (run :193-215, tra :307-510,
generate_logits / extract_pub_feature :622-664): supervised pass with the margin / class-centre losses, then the
inter + intra contrast pass against the server's public features, SGD(lr 1e-4, momentum 0.9, wd 5e-5)."""
from __future__ import annotations

import copy

import torch

from creamfl_b200 import ops
from creamfl_b200.clients import TextClient, resnet18_client, text_supervised_loss, unimodal_supervised_loss
from creamfl_b200.optim import FusedOptimizer
from creamfl_b200.partition import distill_lookup


class ClientTrainer:
    def __init__(self, args, dataset, dst, RGBmean, RGBstdv, data_dict, logger, global_test_set=None, inter_distance=4,
                 loss='softmax', gpuid='cuda:0', num_epochs=30, init_lr=0.0001, decay=0.1, batch_size=512, imgsize=256,
                 num_workers=4, print_freq=10, save_step=10, scale=128, pool_type='max_avg', client_id=-1, wandb=None):
        self.args, self.dset_name, self.logger, self.client_idx = args, dataset, logger, client_id
        self.train_loader = data_dict
        self.device = torch.device(gpuid)
        self.inter_distance, self.scale, self.init_lr = inter_distance, scale, init_lr
        self.local_epochs, self.local_epoch = args.local_epochs, 0
        self.classSize = {'Cifar100': 100, 'Cifar10': 10, 'AG_NEWS': 4, 'YelpReviewPolarity': 2}[dataset]
        self.is_image = dataset in ('Cifar100', 'Cifar10')
        self.setModel()

    def setModel(self):
        """ClientTrainer.py:274-289."""
        if self.is_image:
            self.model = resnet18_client(pretrained=True, num_class=self.classSize, is_train=True, scale=self.scale,
                                         embed_dim=self.args.feature_dim).to(self.device)
            self.model.store()
            self.optimizer = FusedOptimizer(self.model.parameters(), lr=self.init_lr, momentum=0.9, weight_decay=5e-5,
                                            mode='sgd').attach_stores(self.model)
        else:
            self.model = TextClient(embed_dim=self.args.feature_dim, num_class=self.classSize,
                                    scale=self.scale).to(self.device)
            self.model.store()
            self.optimizer = FusedOptimizer(self.model.parameters(), lr=self.init_lr, momentum=0.9, weight_decay=5e-5,
                                            mode='sgd').attach_stores(self.model)

    def run(self, global_img_feature, global_txt_feature, distill_index, global_train_loader):
        self.old_model = copy.deepcopy(self.model)
        self.old_model.eval()
        for _ in range(self.local_epochs):
            self.local_epoch += 1
            self.tra(global_img_feature, global_txt_feature, distill_index, global_train_loader)
        del self.old_model

    def _embed(self, model, images, captions, caption_lens):
        if self.is_image:
            return model(images)
        return model(captions, caption_lens)

    def tra(self, global_img_feature, global_txt_feature, distill_index, global_train_loader):
        dev = self.device
        self.model.train()
        self.model.phase, self.model.is_train = 'None', True
        for data in self.train_loader:                                     # supervised pass (:322-363)
            self.optimizer.zero_grad()
            if self.is_image:
                inputs, labels = data
                loss, _ = unimodal_supervised_loss(self.model, inputs.to(dev), labels.to(dev), self.inter_distance)
            else:
                inputs, labels, caplens = data
                loss, _ = text_supervised_loss(self.model, inputs.to(dev), caplens, labels.to(dev), self.inter_distance)
            loss.backward()
            self.optimizer.step()
        intra, inter = self.args.contrast_local_intra, self.args.contrast_local_inter
        if not (intra or inter):
            return
        g_img, g_txt = global_img_feature.to(dev).float(), global_txt_feature.to(dev).float()
        g_same, g_other = (g_img, g_txt) if self.is_image else (g_txt, g_img)
        g_other16 = ops.to_bf16(g_other)
        lut = distill_lookup(distill_index, dev)
        for m in (self.model, self.old_model):
            m.phase, m.is_train = 'extract_conv_feature', False            # :372-375
        for images, captions, captions_word, caption_lens, _, _, index in global_train_loader:
            self.optimizer.zero_grad()
            d_idx = lut[torch.as_tensor(index, device=dev)]
            images, captions = images.to(dev, non_blocking=True), captions.to(dev, non_blocking=True)
            feat = self._embed(self.model, images, captions, caption_lens)
            loss_inter = loss_moon = None
            if inter:
                loss_inter = ops.infonce_loss(feat, g_other16, d_idx, 2.0)                    # :388,398-401
            if intra:
                with torch.no_grad():
                    old = self._embed(self.old_model, images, captions, caption_lens)
                loss_moon = ops.moon_intra_loss(feat, old, g_same, d_idx, 2.0, feat.shape[0])  # :404-414
            if intra and inter:
                if not self.args.loss_scale:
                    loss = (loss_moon + loss_inter) * self.args.interintra_weight
                else:
                    loss = (loss_moon + loss_inter / (loss_inter / loss_moon).detach()) * self.args.interintra_weight
            else:
                loss = loss_moon if intra else loss_inter
            loss.backward()
            self.optimizer.step()
        for m in (self.model, self.old_model):
            m.phase, m.is_train = 'None', True

    def generate_logits(self, dataloader):
        vec, idx = self.extract_pub_feature(dataloader)
        return ({'img': vec, 'txt': None} if self.is_image else {'img': None, 'txt': vec}), idx

    def extract_pub_feature(self, dataloader):
        """ClientTrainer.py:631-664 - note: no .eval() in the reference, BatchNorm keeps using batch statistics."""
        dev = self.device
        self.model.phase, self.model.is_train = 'extract_conv_feature', False
        feats, distill_index = [], []
        with torch.no_grad():
            for images, captions, captions_word, caption_lens, _, _, index in dataloader:
                f = self._embed(self.model, images.to(dev, non_blocking=True), captions.to(dev, non_blocking=True),
                                caption_lens)
                feats.append(f.clone())
                distill_index.extend(index)
        self.model.phase, self.model.is_train = 'None', True
        return torch.cat(feats), distill_index
