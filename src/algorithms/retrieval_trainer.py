"""Mirror of the reference's src/algorithms/retrieval_trainer.py (EngineBase.create :53-84, TrainerEngine.train
:185-214, evaluate :113-135) on creamfl_b200.engine.ServerEngine."""
from __future__ import annotations

import torch

from creamfl_b200.engine import ServerEngine


class TrainerEngine:
    def __init__(self, device='cuda', partition_train_distill=-1.):
        self.device = device
        self.model = self.criterion = self.optimizer = self.lr_scheduler = self.evaluator = self.logger = None
        self.config = None
        self._core = None

    def set_logger(self, logger):
        self.logger = logger

    def create(self, config, word2idx, evaluator, mlp_local):
        """get_model / get_criterion / get_optimizer('adamp') / get_lr_scheduler('cosine_annealing', T_max=30)
        (retrieval_trainer.py:53-84; coco.yaml:30-38)."""
        self.config = config
        model_cfg = config['model']
        opt_cfg = config.get('optimizer', {})
        self._core = ServerEngine(embed_dim=model_cfg['embed_dim'], cnn_type=model_cfg.get('cnn_type', 'resnet101'),
                                  lr=opt_cfg.get('learning_rate', 2e-4),
                                  grad_clip=config.get('train', {}).get('grad_clip', 2.0),
                                  kd_weight=config.get('kd_weight', 0.3), not_bert=model_cfg.get('not_bert', False),
                                  vocab_size=len(word2idx) if word2idx else 11755,
                                  use_graphs=config.get('cuda_graphs', False))
        self.model, self.criterion, self.optimizer = self._core.model, self._core.criterion, self._core.optimizer
        self.lr_scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(
            self.optimizer, T_max=config.get('lr_scheduler', {}).get('T_max', 30))
        evaluator.set_model(self.model)
        evaluator.set_criterion(self.criterion)
        self.evaluator = evaluator
        self.evaluator.set_logger(self.logger)

    def model_to_device(self):
        self.model.to(self._core.device)

    def to_half(self):
        """apex O2 in the reference (retrieval_trainer.py:107-111); the CUDA towers already run bf16 storage with
        fp32 accumulation and need no loss scaling."""

    def text_input(self, captions, captions_word, caption_lens):
        """What the server's text tower consumes from a loader batch: BERT tokens / caption strings
        (pcme.py:40-43), or the vocabulary-id sentences + lengths of the `--not_bert` GRU server (pcme.py:37-38)."""
        if self._core.not_bert:
            return captions.to(self._core.device, non_blocking=True), caption_lens
        return captions_word

    def train(self, tr_loader, pub_data_ratio=1.):
        """One public-data epoch (retrieval_trainer.py:185-214)."""
        dev = self._core.device
        last = None
        for idx, (images, captions, captions_word, caption_lens, _a, _b, index) in enumerate(tr_loader):
            if idx == int(len(tr_loader) * pub_data_ratio):
                break
            last = self._core.train_step(images.to(dev, non_blocking=True),
                                         self.text_input(captions, captions_word, caption_lens))
        return last

    @torch.no_grad()
    def evaluate(self, val_loaders, n_crossfolds=None, **kwargs):
        """retrieval_trainer.py:113-135."""
        if self.evaluator is None:
            raise RuntimeError('evaluator is not set')
        self.model.eval()
        scores = {}
        for key, loader in val_loaders.items():
            scores[key] = self.evaluator.evaluate(loader, n_crossfolds=n_crossfolds, key=key, **kwargs)
        return scores

    def report_scores(self, step, scores, metadata, prefix=''):
        if self.logger is not None:
            self.logger.log(f'[Eval] Report @step {step}: {scores}')
