"""Mirror of the reference's src/algorithms/MMFL.py orchestrator (MMFL.__init__ :40-68, create_model :116-178,
load_dataset :90-114, train :180-289, distill :291-391) on the creamfl_b200 kernels.

Differences that are deliberate and stated: datasets are synthetic COCO / Flickr / CIFAR / AG_NEWS-shape generators
(no dataset is on the box; `--pub_data_num` is honoured instead of the reference's hard-coded 50000,
MMFL.py:302,319); wandb logging is optional; public features never leave the device (the reference round-trips them
through the host, MMFL.py:209-210)."""
from __future__ import annotations

import random
import time

import torch

from creamfl_b200 import engine as _engine, ops
from creamfl_b200.partition import data_partitioner, distill_lookup, public_subset_indices, shard_partition

from .ClientTrainer import ClientTrainer
from .MMClientTrainer import MMClientTrainer
from .eval_coco import COCOEvaluator
from .retrieval_trainer import TrainerEngine

import importlib.util
from pathlib import Path

# `datasets` is also the name of an installed third-party package: load the synthetic loaders by path
_spec = importlib.util.spec_from_file_location('creamfl_src_synthetic',
                                               Path(__file__).resolve().parent.parent / 'datasets' / 'synthetic.py')
_synthetic = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_synthetic)
SyntheticLabelled, SyntheticPairs = _synthetic.SyntheticLabelled, _synthetic.SyntheticPairs


class _Logger:
    def log(self, msg):
        print(f'[{time.strftime("%H:%M:%S")}] {msg}', flush=True)


class MMFL:
    def __init__(self, args, wandb=None):
        self.args, self.wandb = args, wandb
        self.device = None
        self.img_local_trainers = self.txt_local_trainers = self.mm_local_trainers = None
        self.engine = None
        self.best_score, self.cur_epoch = 0, 0
        self.best_scores, self.best_metadata = {}, {}
        self.img_vec = self.txt_vec = None
        self.logger = _Logger()
        self.set_config()

    def set_config(self, img='cifar100', txt='AG_NEWS'):
        """MMFL.py:70-88: coco.yaml with embed_dim = --feature_dim, ResNet101 + BERT unless --not_bert."""
        a = self.args
        self.config = {
            'model': {'embed_dim': a.feature_dim, 'cnn_type': 'resnet50' if a.not_bert else 'resnet101',
                      'not_bert': bool(a.not_bert), 'n_samples_inference': 7},           # MMFL.py:82-85
            'cuda_graphs': not getattr(a, 'no_cuda_graphs', True),
            'optimizer': {'name': 'adamp', 'learning_rate': a.server_lr, 'weight_decay': 0.0},
            'lr_scheduler': {'name': 'cosine_annealing', 'T_max': 30},
            'criterion': {'name': 'pcme', 'init_negative_scale': 15, 'init_shift': 15},
            'train': {'grad_clip': 2, 'use_fp16': True},
            'kd_weight': a.kd_weight,
        }

    # ------------------------------------------------------------------------------------------------ data
    def load_dataset(self, args):
        """MMFL.py:90-114.  Public subset = `pub_data_num` pairs; test = COCO-1K shape folds."""
        n = args.pub_data_num
        bs = getattr(args, 'pub_batch_size', 128)
        subset = public_subset_indices(n)         # load_datasets.py:148-157; n = 50000: the shipped coco_subset_idx_file
        self.dataloaders_global = {
            f'train_subset_{n}': SyntheticPairs(subset, bs, seed=1, shuffle=True, image_size=args.image_size),
            f'train_subset_eval_{n}': SyntheticPairs(subset, 2 * bs, seed=1, image_size=args.image_size),
            'test': SyntheticPairs(list(range(1000000, 1000000 + args.test_images)), 2 * bs, seed=2,
                                   image_size=args.image_size),
        }
        self.engine = TrainerEngine()
        self.engine.set_logger(self.logger)
        self.engine.create(self.config, None, COCOEvaluator(), args.mlp_local)
        self.engine.model_to_device()
        self.engine.to_half()

    def create_model(self, args):
        """MMFL.py:116-178: Dirichlet(0.1) partitions for the unimodal clients, Flickr shards for the multimodal ones."""
        self.logger.log('start creating model and partition datasets')
        self.device = _engine.default_device(args.device)
        self.img_local_trainers, self.txt_local_trainers, self.mm_local_trainers = [], [], []
        n_priv = args.private_samples
        if args.num_img_clients > 0:
            import numpy as np
            part = data_partitioner('cifar100', n_priv, args.num_img_clients, 'hetero', 0.1, np.arange(n_priv) % 100,
                                    seed=2021, min_size=min(10, n_priv // 400))     # get_FL_trainloader(.., num_clients)
            for i in range(args.num_img_clients):
                loader = SyntheticLabelled('image', part[i], 512, 100, image_size=args.client_image_size, seed=i)
                self.img_local_trainers.append(ClientTrainer(args, 'Cifar100', 'Cifar100', None, None, loader,
                                                             self.logger, gpuid=str(self.device), client_id=i))
        if args.num_txt_clients > 0:
            import numpy as np
            part = data_partitioner('AG_NEWS', n_priv, args.num_txt_clients, 'hetero', 0.1, np.arange(n_priv) % 4,
                                    seed=2021, min_size=min(3000, n_priv // 40))
            for i in range(args.num_txt_clients):
                loader = SyntheticLabelled('text', part[i], 512, 4, seed=100 + i)
                self.txt_local_trainers.append(ClientTrainer(args, 'AG_NEWS', 'AG_NEWS', None, None, loader,
                                                             self.logger, gpuid=str(self.device), client_id=i))
        if args.num_mm_clients > 0:
            shards = shard_partition(max(150, n_priv), 15, 150, seed=2021)
            for i in range(args.num_mm_clients):
                loader = SyntheticPairs(shards[i].tolist(), 128, seed=200 + i, image_size=args.image_size)
                self.mm_local_trainers.append(MMClientTrainer(args, loader, client=i, logger=self.logger))
        self.total_local_trainers = self.img_local_trainers + self.txt_local_trainers + self.mm_local_trainers
        for i, t in enumerate(self.total_local_trainers):
            t.client_idx = i + 1

    # ------------------------------------------------------------------------------------------------ round
    def train(self, round_n):
        """MMFL.py:180-289."""
        args = self.args
        n = args.pub_data_num
        self.cur_epoch = round_n
        self.cur_trainers = self.total_local_trainers
        self.engine.train(self.dataloaders_global[f'train_subset_{n}'])                                   # step 1
        self.cur_trainers = random.sample(self.total_local_trainers, args.client_num_per_round)           # step 2
        core = self.engine._core                                                                          # step 3
        img_feature, txt_feature, distill_index = [], [], []
        for images, captions, captions_word, caption_lens, _, _, index in self.dataloaders_global[f'train_subset_eval_{n}']:
            fi, ft = core.extract(images.to(core.device, non_blocking=True),
                                  self.engine.text_input(captions, captions_word, caption_lens))
            img_feature.append(fi.clone())
            txt_feature.append(ft.clone())
            distill_index.extend(index)
        self.global_img_feature, self.global_txt_feature = torch.cat(img_feature), torch.cat(txt_feature)
        self.distill_index = distill_index
        img_vec, img_num, txt_vec, txt_num = [], [], [], []                                               # step 4
        for trainer in self.cur_trainers:
            self.logger.log(f'Training Client {trainer.client_idx}!')
            trainer.cur_epoch = round_n
            trainer.run(self.global_img_feature, self.global_txt_feature, distill_index,
                        self.dataloaders_global[f'train_subset_{n}'])
            vec, i = trainer.generate_logits(self.dataloaders_global[f'train_subset_eval_{n}'])
            assert i == self.distill_index                                                                # MMFL.py:239
            if vec['img'] is not None:
                img_vec.append(vec['img'])
                img_num.append(len(trainer.train_loader.indices))
            if vec['txt'] is not None:
                txt_vec.append(vec['txt'])
                txt_num.append(len(trainer.train_loader.indices))
        if not args.disable_distill:                                                                      # step 5
            self.distill(round_n, img_vec, txt_vec, img_num, txt_num, self.distill_index)
        test_scores = self.engine.evaluate({'test': self.dataloaders_global['test']},                     # step 6
                                           n_crossfolds=args.test_folds,
                                           n_images_per_crossfold=args.test_images // max(1, args.test_folds),
                                           n_captions_per_crossfold=args.test_images // max(1, args.test_folds))
        self.engine.report_scores(step=round_n + 1, scores=test_scores, metadata=self.best_metadata)
        rsum = test_scores['test']['n_fold']['rsum'] if args.test_folds else test_scores['test']['rsum']
        if self.best_score < rsum:
            self.best_score, self.best_scores = rsum, test_scores        # (the reference never updates best_score, :276)
            self.best_metadata = {'best_score': rsum, 'best_epoch': round_n + 1}
            core.save_checkpoint(args.name + '-best_model.pt')                                            # :281
        if round_n == args.comm_rounds - 1:
            core.save_checkpoint(args.name + '-last_model.pt')                                            # :284
        self.engine.lr_scheduler.step()                                                                   # step 7
        return test_scores

    def distill(self, round_n, img_vec, txt_vec, img_num, txt_num, distill_index):
        """MMFL.py:291-391: con_w aggregation (:298-335), then one public epoch of kd-MSE distillation (:346-391)."""
        args, core = self.args, self.engine._core
        n = args.pub_data_num
        if args.agg_method != 'con_w':
            raise NotImplementedError('the accelerated path implements agg_method = con_w')
        agg_img = ops.conw_aggregate(img_vec, self.global_txt_feature) if img_vec else None               # :298-314
        agg_txt = ops.conw_aggregate(txt_vec, self.global_img_feature) if txt_vec else None               # :317-331
        self.img_vec, self.txt_vec = agg_img, agg_txt
        lut = distill_lookup(distill_index, core.device)
        # the reference adds the image (text) MSE once per configured client type carrying it (:361-378)
        img_terms = int(args.num_img_clients > 0) + int(args.num_mm_clients > 0)
        txt_terms = int(args.num_txt_clients > 0) + int(args.num_mm_clients > 0)
        self.logger.log('start distilling')
        for images, captions, captions_word, caption_lens, _, _, index in self.dataloaders_global[f'train_subset_{n}']:
            d_idx = lut[torch.as_tensor(index, device=core.device)]
            core.distill_step(images.to(core.device, non_blocking=True),
                              self.engine.text_input(captions, captions_word, caption_lens), d_idx, agg_img, agg_txt,
                              img_terms=img_terms if agg_img is not None else 0,
                              txt_terms=txt_terms if agg_txt is not None else 0)
