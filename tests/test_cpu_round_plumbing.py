"""BASELINE.json configs[0]: "1 FL round, 2 clients, ResNet101+BERT server, 16 synthetic COCO-shape pairs, CPU
(plumbing, no GPU)".

The product's own orchestrator (src/algorithms/MMFL.py: create_model -> load_dataset -> train(round) with public
training, extraction, client rounds, con_w aggregation, distillation, Recall@K evaluation) runs one communication
round on the CPU.  Every C-ABI entry point is replaced by the test-only emulation of tests/kernel_emulation.py (kernel
semantics from include/creamfl_b200.h, losses / aggregation / optimizer through the CPU oracle); everything above the
ABI - engines, trainers, parameter stores, index plumbing - is the shipped code.  Nothing here is a product CPU path:
without the emulation the same call raises (last test)."""
import sys
import types
from pathlib import Path

import pytest
import torch

import kernel_emulation as KE  # tests/ is on sys.path

ROOT = Path(__file__).resolve().parent.parent


def _args(**over):
    a = dict(name='plumbing', local_epochs=1, comm_rounds=1, seed=0, device=0, num_img_clients=0, num_txt_clients=0,
             num_mm_clients=2, client_num_per_round=2, server_lr=1e-5, disable_distill=False, agg_method='con_w',
             contrast_local_intra=True, contrast_local_inter=True, mlp_local=False, kd_weight=0.3,
             interintra_weight=0.5, loss_scale=False, save_client=False, pub_data_num=16, feature_dim=256,
             not_bert=False, private_samples=150, image_size=224, client_image_size=64, test_images=16, test_folds=2,
             pub_batch_size=16)
    a.update(over)
    return types.SimpleNamespace(**a)


def _mmfl():
    for p in (str(ROOT), str(ROOT / 'src')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from src.algorithms.MMFL import MMFL
    return MMFL


def test_one_round_two_clients_on_cpu(monkeypatch, tmp_path):
    KE.install_engine(monkeypatch)
    import random
    random.seed(0)
    torch.manual_seed(0)
    MMFL = _mmfl()
    args = _args(name=str(tmp_path / 'plumbing'))        # MMFL.py:281,284 saves '<name>-{best,last}_model.pt' (0.6 GB)
    algo = MMFL(args, None)
    algo.create_model(args)
    algo.load_dataset(args)
    # shrink the two multimodal clients' private shards to one batch each: this is a plumbing run
    for t in algo.mm_local_trainers:
        t.train_loader.indices = t.train_loader.indices[:16]
        t.train_loader.batch_size = 16
    server = algo.engine.model
    assert type(server).__name__ == 'PCME' and sum(p.numel() for p in server.parameters()) > 150e6   # ResNet101 + BERT
    before = server.store().flat.clone()
    scores = algo.train(0)
    # public representations of the server and of both clients were produced for the same 16 items, in loader order
    assert algo.global_img_feature.shape == (16, 256) and algo.global_txt_feature.shape == (16, 256)
    assert len(algo.distill_index) == 16 and sorted(algo.distill_index) == algo.distill_index
    # con_w aggregates: convex combinations of unit-norm client rows
    for agg in (algo.img_vec, algo.txt_vec):
        assert agg.shape == (16, 256) and torch.isfinite(agg).all() and float(agg.norm(dim=1).max()) <= 1.0 + 1e-4
    # the server moved (public training + distillation), shadows follow the masters, the scheduler stepped
    assert not torch.equal(before, server.store().flat)
    w = server.linear.weight
    assert torch.equal(w._w16, w.data.to(torch.bfloat16))
    assert algo.engine.optimizer.param_groups[0]['lr'] < args.server_lr
    # Recall@K report of the COCO-1K style protocol (2 folds of 8 here)
    nf = scores['test']['n_fold']
    for d in ('i2t', 't2i'):
        assert 0.0 <= nf[d]['recall_1'] <= nf[d]['recall_5'] <= nf[d]['recall_10'] <= 100.0
    assert scores['test']['rsum'] == scores['test']['i2t']['rsum'] + scores['test']['t2i']['rsum']
    # checkpoints in the reference's format: {'net': state_dict} with torchvision / HF key names
    ck = torch.load(tmp_path / 'plumbing-last_model.pt', map_location='cpu')
    assert set(ck) == {'net'} and 'img_enc.cnn.layer3.22.conv2.weight' in ck['net'] and \
        'txt_enc.encoder.layer.11.output.dense.weight' in ck['net'] and ck['net']['linear.weight'].shape == (256, 768)
    assert (tmp_path / 'plumbing-best_model.pt').exists()
    # clients: models trained, old-model snapshot kept by address for the next round
    for t in algo.mm_local_trainers:
        assert t._core.old_model is not None and t.local_epoch == 1


def test_one_round_image_text_and_multimodal_client_on_cpu(monkeypatch, tmp_path):
    """The mix tests/test_gpu_round.py runs on the GPU through `python src/main.py` (one CIFAR-shape image client, one
    AG_NEWS-shape text client, one multimodal client), here through the same orchestrator on the CPU: unimodal
    trainers (ClientTrainer.run / generate_logits with an absent modality), con_w over the clients that carry each
    modality, distillation with the per-client-type term counts (2 + 2 with all three types, MMFL.py:361-378)."""
    KE.install_engine(monkeypatch)
    import random
    random.seed(0)
    torch.manual_seed(0)
    MMFL = _mmfl()
    args = _args(name=str(tmp_path / 'mix'), num_img_clients=1, num_txt_clients=1, num_mm_clients=1,
                 client_num_per_round=3, private_samples=3000, client_image_size=32)
    algo = MMFL(args, None)
    algo.create_model(args)
    algo.load_dataset(args)
    for t in algo.total_local_trainers:                  # one private batch of 16 per client: a plumbing run
        t.train_loader.indices = t.train_loader.indices[:16]
        t.train_loader.batch_size = 16
    kinds = [type(t).__name__ for t in algo.total_local_trainers]
    assert kinds == ['ClientTrainer', 'ClientTrainer', 'MMClientTrainer']
    assert [t.client_idx for t in algo.total_local_trainers] == [1, 2, 3]
    # the public subset comes from the reference's index producer (sorted caption indices below 566 435)
    pub = algo.dataloaders_global['train_subset_16'].indices
    assert pub == sorted(pub) and len(set(pub)) == 16 and max(pub) < 566435
    seen = {}
    real_distill = algo.engine._core.distill_step

    def spy(images, tokens, d_idx, agg_img, agg_txt, img_terms=1, txt_terms=1):
        seen['terms'] = (img_terms, txt_terms)
        return real_distill(images, tokens, d_idx, agg_img, agg_txt, img_terms=img_terms, txt_terms=txt_terms)
    monkeypatch.setattr(algo.engine._core, 'distill_step', spy)
    scores = algo.train(0)
    assert seen['terms'] == (2, 2)
    assert sorted(type(t).__name__ for t in algo.cur_trainers) == sorted(kinds)          # all three were selected
    for agg in (algo.img_vec, algo.txt_vec):             # image client + mm client / text client + mm client
        assert agg.shape == (16, 256) and torch.isfinite(agg).all() and float(agg.norm(dim=1).max()) <= 1.0 + 1e-4
    img_t, txt_t, mm_t = algo.total_local_trainers
    assert img_t.is_image and not txt_t.is_image
    assert type(img_t.criterion).__name__ == 'CrossEntropyLoss'
    assert img_t.local_epoch == txt_t.local_epoch == mm_t.local_epoch == 1
    assert 0.0 <= scores['test']['i2t']['recall_1'] <= 100.0


def test_orchestrator_refuses_cpu_without_emulation():
    MMFL = _mmfl()
    args = _args()
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present: the orchestrator runs on it')
    algo = MMFL(args, None)
    with pytest.raises(RuntimeError, match='CUDA'):
        algo.create_model(args)
