"""Generate tests/golden/*.npz by executing the REFERENCE's own code (FLAIR-THU/CreamFL at /root/reference).

Run once in the build container (the reference tree does not travel to the GPU box):

    python tests/golden/make_golden.py [case ...]

The reference has no tests, so these files are the pin for oracle/creamfl_oracle.py and, through it, for the
CUDA path.  Where the reference code is a method with inline arithmetic (MMClientTrainer.train_epoch,
ClientTrainer.tra, MMFL.distill) the method itself is driven with stub models, loaders and optimizers, and loss /
gradients are recorded by instrumenting Tensor.backward - the arithmetic executed is the reference's, untouched.

Import shims: torchtext, apex, adamp, munch, nltk, pycocotools, fire are not installed in this image; stub modules
are inserted so that the reference packages import.  `.cuda()` is patched to the identity (no GPU here).
"""
from __future__ import annotations

import hashlib
import os
import pickle
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

REF = Path('/root/reference')
OUT = Path(__file__).resolve().parent


# --------------------------------------------------------------------------------------------- import shims
class Munch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class _Anything:
    """Placeholder for any attribute of a stubbed third-party module (class, function or constant)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything()


class _StubModule(types.ModuleType):
    __path__ = []  # behave like a package so that `import a.b.c` works

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Anything


class _StubFinder:
    """Serves stub modules for the third-party packages the image lacks (and all their submodules)."""
    ROOTS = ('torchtext', 'apex', 'nltk', 'pycocotools', 'fire', 'wandb')

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        if name.split('.')[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        pass


def install_shims():
    sys.meta_path.insert(0, _StubFinder())
    m = types.ModuleType('munch')
    m.Munch = Munch
    m.munchify = lambda d: Munch(d)
    sys.modules['munch'] = m
    a = types.ModuleType('adamp')
    a.AdamP = torch.optim.Adam
    sys.modules['adamp'] = a
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    sys.path[:0] = [str(REF / 'src'), str(REF)]
    os.chdir(REF)


class BackwardRecorder:
    """Records the value of every tensor on which .backward() is called."""

    def __init__(self):
        self.values = []
        self._orig = torch.Tensor.backward

    def __enter__(self):
        rec = self

        def backward(t, *a, **k):
            rec.values.append(float(t.detach()))
            return rec._orig(t, *a, **k)

        torch.Tensor.backward = backward
        return self

    def __exit__(self, *exc):
        torch.Tensor.backward = self._orig


class NullOptimizer:
    param_groups = [{'lr': 0.0}]

    def zero_grad(self):
        pass

    def step(self):
        pass


class NullLogger:
    def log(self, *a, **k):
        pass


class FeatureStub(nn.Module):
    """Stands in for a PCME model: looks features up by the integer ids carried in `images`; keeps the output
    tensors so that their gradients can be read back after backward."""

    def __init__(self, img, txt, train_param=True):
        super().__init__()
        self.img = nn.Parameter(img.clone(), requires_grad=train_param)
        self.txt = nn.Parameter(txt.clone(), requires_grad=train_param)
        self.last = None

    def forward(self, images, sentences, captions_word, lengths):
        ids = images.long().view(-1)
        oi = self.img[ids] * 1.0
        ot = self.txt[ids] * 1.0
        if oi.requires_grad:
            oi.retain_grad()
            ot.retain_grad()
        self.last = (oi, ot)
        return {'image_features': oi, 'caption_features': ot}


class UniStub(nn.Module):
    def __init__(self, feats, train_param=True):
        super().__init__()
        self.f = nn.Parameter(feats.clone(), requires_grad=train_param)
        self.phase = 'None'
        self.is_train = True
        self.last = None

    def forward(self, x, lengths=None):
        ids = x.long().view(-1)
        o = self.f[ids] * 1.0
        if o.requires_grad:
            o.retain_grad()
        self.last = o
        return o


def unit(x):
    return x / x.norm(dim=-1, keepdim=True)


# --------------------------------------------------------------------------------------------- cases
def case_pcme():
    from criterions.probemb import MCSoftContrastiveLoss
    out = {}
    for tag, (n, d, s, b, normalise) in {'a': (16, 32, 15.0, 15.0, True), 'b': (5, 8, 3.0, 2.0, False),
                                          'c': (33, 64, 15.0, 15.0, True)}.items():
        g = torch.Generator().manual_seed(100 + n)
        img = torch.randn(n, d, generator=g, dtype=torch.float64)
        txt = torch.randn(n, d, generator=g, dtype=torch.float64) * 0.5 + img * 0.7
        if normalise:
            img, txt = unit(img), unit(txt)
        crit = MCSoftContrastiveLoss(Munch(init_shift=b, init_negative_scale=s, num_samples=7)).double()
        img.requires_grad_(True)
        txt.requires_grad_(True)
        loss, info = crit(img, txt, None, None)
        loss.backward()
        out.update({f'{tag}_img': img.detach().numpy(), f'{tag}_txt': txt.detach().numpy(),
                    f'{tag}_shift': np.float64(b), f'{tag}_scale': np.float64(s),
                    f'{tag}_loss': np.float64(loss.item()),
                    f'{tag}_i2t_pos': np.float64(info['i2t_pos_loss']), f'{tag}_i2t_neg': np.float64(info['i2t_neg_loss']),
                    f'{tag}_t2i_loss': np.float64(info['t2i_loss']),
                    f'{tag}_d_img': img.grad.numpy(), f'{tag}_d_txt': txt.grad.numpy(),
                    f'{tag}_d_shift': crit.shift.grad.numpy(), f'{tag}_d_scale': crit.negative_scale.grad.numpy()})
    np.savez_compressed(OUT / 'pcme.npz', **out)


def _public_setup(seed, n_pub=96, d=32, b=8):
    g = torch.Generator().manual_seed(seed)
    g_img = unit(torch.randn(n_pub, d, generator=g, dtype=torch.float64))
    g_txt = unit(0.7 * g_img + 0.5 * torch.randn(n_pub, d, generator=g, dtype=torch.float64))
    cur_img = unit(g_img + 0.4 * torch.randn(n_pub, d, generator=g, dtype=torch.float64))
    cur_txt = unit(g_txt + 0.4 * torch.randn(n_pub, d, generator=g, dtype=torch.float64))
    old_img = unit(cur_img + 0.2 * torch.randn(n_pub, d, generator=g, dtype=torch.float64))
    old_txt = unit(cur_txt + 0.2 * torch.randn(n_pub, d, generator=g, dtype=torch.float64))
    # dataset indices are arbitrary ints; distill_index lists them in bank-row order
    distill_index = (torch.randperm(10 * n_pub, generator=g)[:n_pub] + 9).tolist()
    rows = torch.randperm(n_pub, generator=g)[:b].tolist()
    batch_index = [distill_index[r] for r in rows]
    return dict(g_img=g_img, g_txt=g_txt, cur_img=cur_img, cur_txt=cur_txt, old_img=old_img, old_txt=old_txt,
                distill_index=distill_index, rows=rows, batch_index=batch_index)


def case_mm_contrast():
    import src.algorithms.MMClientTrainer as M
    out = {}
    variants = {'both': (True, True, False), 'intra': (True, False, False), 'inter': (False, True, False),
                'both_scaled': (True, True, True)}
    for tag, (intra, inter, scaled) in variants.items():
        s = _public_setup(7)
        tr = M.MMClientTrainer.__new__(M.MMClientTrainer)
        tr.args = Munch(contrast_local_intra=intra, contrast_local_inter=inter, interintra_weight=0.5,
                        loss_scale=scaled, feature_dim=32)
        tr.config = Munch(train=Munch(use_fp16=False, grad_clip=0))
        tr.device = 'cpu'
        tr.cur_epoch = 0
        tr.optimizer = NullOptimizer()
        tr.model = FeatureStub(s['cur_img'], s['cur_txt'])
        tr.old_model = FeatureStub(s['old_img'], s['old_txt'], train_param=False)
        tr.criterion = lambda **kw: (kw['image_features'].sum() * 0.0, {'loss': 0.0})
        rows = torch.tensor(s['rows'], dtype=torch.float64)
        dummy = torch.zeros(len(s['rows']), 4, dtype=torch.long)
        lens = torch.full((len(s['rows']),), 4)
        private = [(rows, dummy, ('x',) * len(s['rows']), lens, None, None, s['batch_index'])]
        tr.train_loader = private
        public = [(rows, dummy, ('x',) * len(s['rows']), lens, None, None, s['batch_index'])]
        with BackwardRecorder() as rec:
            tr.train_epoch(s['g_img'].clone(), s['g_txt'].clone(), s['distill_index'], public)
        oi, ot = tr.model.last
        out.update({f'{tag}_loss': np.float64(rec.values[-1]),
                    f'{tag}_d_img': oi.grad.numpy(), f'{tag}_d_txt': ot.grad.numpy()})
    s = _public_setup(7)
    out.update({k: s[k].numpy() for k in ('g_img', 'g_txt', 'cur_img', 'cur_txt', 'old_img', 'old_txt')})
    out['distill_index'] = np.asarray(s['distill_index'], dtype=np.int64)
    out['rows'] = np.asarray(s['rows'], dtype=np.int64)
    out['batch_index'] = np.asarray(s['batch_index'], dtype=np.int64)
    np.savez_compressed(OUT / 'mm_contrast.npz', **out)


def case_uni_contrast():
    import src.algorithms.ClientTrainer as CT
    out = {}
    variants = {'both': (True, True, False), 'intra': (True, False, False), 'inter': (False, True, False),
                'both_scaled': (True, True, True)}
    for dset in ('Cifar100', 'AG_NEWS'):
        for tag, (intra, inter, scaled) in variants.items():
            s = _public_setup(11)
            tr = CT.ClientTrainer.__new__(CT.ClientTrainer)
            tr.args = Munch(contrast_local_intra=intra, contrast_local_inter=inter, interintra_weight=0.5,
                            loss_scale=scaled, feature_dim=32)
            tr.gpuid = 'cpu'
            tr.dset_name = dset
            tr.logger = NullLogger()
            tr.local_epoch = 0
            tr.losses, tr.top1, tr.top5 = CT.AverageMeter(), CT.AverageMeter(), CT.AverageMeter()
            tr.optimizer = NullOptimizer()
            tr.criterion = nn.CrossEntropyLoss()
            own_cur = s['cur_img'] if dset == 'Cifar100' else s['cur_txt']
            own_old = s['old_img'] if dset == 'Cifar100' else s['old_txt']
            tr.model = UniStub(own_cur)
            tr.old_model = UniStub(own_old, train_param=False)
            tr.train_loader = []
            rows = torch.tensor(s['rows'], dtype=torch.float64)
            lens = torch.full((len(s['rows']),), 4)
            # the unimodal text client reads `captions` (2nd slot) as model input, the image client `images`
            public = [(rows, rows, ('x',) * len(s['rows']), lens, None, None, s['batch_index'])]
            with BackwardRecorder() as rec:
                tr.tra(s['g_img'].clone(), s['g_txt'].clone(), s['distill_index'], public)
            out[f'{dset}_{tag}_loss'] = np.float64(rec.values[-1])
            out[f'{dset}_{tag}_d_feat'] = tr.model.last.grad.numpy()
    np.savez_compressed(OUT / 'uni_contrast.npz', **out)


def conw_inputs(seed, n, d, n_clients):
    """Deterministic inputs shared by the golden generator and the tests (numpy Generator, PCG64)."""
    rng = np.random.default_rng(seed)

    def unit_np(x):
        return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)

    g_img = unit_np(rng.standard_normal((n, d)))
    g_txt = unit_np(0.7 * g_img + 0.5 * unit_np(rng.standard_normal((n, d))))
    i_vecs = [unit_np(g_img + (0.3 + 0.35 * c) * unit_np(rng.standard_normal((n, d)))) for c in range(n_clients)]
    t_vecs = [unit_np(g_txt + (0.3 + 0.35 * c) * unit_np(rng.standard_normal((n, d)))) for c in range(n_clients)]
    return g_img, g_txt, i_vecs, t_vecs


def case_conw():
    """Drives MMFL.distill (aggregation closure, MMFL.py:298-335) at the only size it supports: N = 50000."""
    import src.algorithms.MMFL as F
    n, d, c = 50000, 64, 3
    g_img, g_txt, i_vecs, t_vecs = conw_inputs(2024, n, d, c)
    algo = F.MMFL.__new__(F.MMFL)
    algo.args = Munch(agg_method='con_w', pub_data_num=n, num_img_clients=0, num_txt_clients=0, num_mm_clients=0,
                      kd_weight=0.3)
    algo.config = Munch(model=Munch(use_img_client=True, use_txt_client=True, use_mm_client=True),
                        train=Munch(use_fp16=False, grad_clip=0))
    algo.engine = Munch(model=nn.Identity())
    algo.logger = NullLogger()
    algo.global_img_feature = torch.from_numpy(g_img)
    algo.global_txt_feature = torch.from_numpy(g_txt)
    algo.dataloaders_global = {f'train_subset_{n}': []}
    algo.distill(0, [torch.from_numpy(v) for v in i_vecs], [torch.from_numpy(v) for v in t_vecs], [1] * c,
                 [1] * c, list(range(n)))
    sel = np.arange(0, n, 397)
    np.savez_compressed(OUT / 'conw.npz', seed=np.int64(2024), n=np.int64(n), d=np.int64(d), c=np.int64(c),
                        rows=sel, img_rows=algo.img_vec.numpy()[sel], txt_rows=algo.txt_vec.numpy()[sel],
                        img_sum=np.float64(algo.img_vec.double().sum().item()),
                        txt_sum=np.float64(algo.txt_vec.double().sum().item()),
                        img_abs_sum=np.float64(algo.img_vec.double().abs().sum().item()),
                        txt_abs_sum=np.float64(algo.txt_vec.double().abs().sum().item()))


def case_recall():
    from src.algorithms.eval_coco import COCOEvaluator
    out = {}
    for tag, (n_img, per, d, n_emb) in {'a': (120, 5, 32, 1), 'b': (64, 5, 16, 7)}.items():
        g = torch.Generator().manual_seed(5 + n_img)
        img = unit(torch.randn(n_img, d, generator=g))
        cap = unit(img.repeat_interleave(per, 0) + 0.9 * torch.randn(n_img * per, d, generator=g))
        img_lab = np.arange(n_img)
        cap_lab = np.repeat(np.arange(n_img), per)
        ev = COCOEvaluator(eval_method='matmul', verbose=False, eval_device='cpu', n_crossfolds=5)
        ev.n_embeddings = n_emb
        qi = img.double()[:, None, :].repeat(1, n_emb, 1)
        qc = cap.double()[:, None, :].repeat(1, n_emb, 1)
        i2t = ev.evaluate_recall(qi, qc, img_lab, cap_lab)
        t2i = ev.evaluate_recall(qc, qi, cap_lab, img_lab)
        out.update({f'{tag}_img': img.numpy(), f'{tag}_cap': cap.numpy(), f'{tag}_img_lab': img_lab,
                    f'{tag}_cap_lab': cap_lab})
        for name, sc in (('i2t', i2t), ('t2i', t2i)):
            for k, v in sc.items():
                out[f'{tag}_{name}_{k}'] = np.float64(v)
    np.savez_compressed(OUT / 'recall.npz', **out)


def case_partition():
    from src.datasets.load_FL_datasets import data_partitioner
    from src.datasets.flickr30k import F30kCaptionsCap
    out = {}

    def digest(part):
        h = hashlib.sha256()
        for j in sorted(part):
            h.update(np.asarray(part[j], dtype=np.int64).tobytes())
        return h.hexdigest()

    for tag, (name, n, k_classes, nets, alpha, seed) in {
            'synth': ('synth', 50000, 100, 4, 0.1, 2021),
            'cifar': ('cifar100', 50000, 100, 10, 0.1, 2021),
            'agnews': ('AG_NEWS', 120000, 4, 10, 0.1, 7)}.items():
        with tempfile.TemporaryDirectory() as tmp:
            np.random.seed(seed)
            y = np.arange(n) % k_classes
            part = data_partitioner(name, n, nets, partition='hetero', check_dir=tmp + '/', alpha=alpha, y_train=y)
        out[f'{tag}_sizes'] = np.asarray([len(part[j]) for j in range(nets)], dtype=np.int64)
        out[f'{tag}_head'] = np.asarray([part[j][:8] for j in range(nets)], dtype=np.int64)
        out[f'{tag}_sha256'] = np.asarray(digest(part))
        out[f'{tag}_args'] = np.asarray([n, k_classes, nets, seed], dtype=np.int64)
        out[f'{tag}_alpha'] = np.float64(alpha)
    # Flickr30k shard partition
    stub = types.SimpleNamespace(data=[None] * 145000)
    with tempfile.TemporaryDirectory() as tmp:
        np.random.seed(2021)
        part = F30kCaptionsCap.non_iid(stub, root=tmp + '/', num_users=15)
    out['f30k_sizes'] = np.asarray([len(part[j]) for j in range(15)], dtype=np.int64)
    out['f30k_head'] = np.asarray([np.asarray(part[j][:8]) for j in range(15)], dtype=np.int64)
    out['f30k_sha256'] = np.asarray(digest(part))
    # hashes of the fixtures the reference ships (they pin the index semantics of the real datasets)
    for fn in ('client_cifar100_noniid.pkl', 'client_AG_NEWS_noniid.pkl', 'client_noniid_flicker30k.pkl'):
        part = pickle.load(open(REF / 'data_partition' / fn, 'rb'))
        out['fixture_' + fn.split('.')[0] + '_sha256'] = np.asarray(digest(part))
        out['fixture_' + fn.split('.')[0] + '_sizes'] = np.asarray([len(part[j]) for j in sorted(part)], dtype=np.int64)
    sub = pickle.load(open(REF / 'coco_subset_idx_file', 'rb'))
    out['fixture_coco_subset_sha256'] = np.asarray(hashlib.sha256(np.asarray(sub, dtype=np.int64).tobytes()).hexdigest())
    out['fixture_coco_subset_len'] = np.int64(len(sub))
    np.savez_compressed(OUT / 'partition.npz', **out)


def case_towers():
    """The reference's own PCME (ResNet18 + BERT glue, pcme.py / image_encoder.py / pie_model.py) and resnet18_client
    (resnet_client.py) on deterministic weights: pins oracle/torch_towers.py (the restatement the CUDA towers are
    tested against).  torchvision / transformers are the versions in this image (the reference pins older ones)."""
    import torchvision
    from transformers import BertConfig, BertModel
    import importlib.util          # by path: /root/repo must not join sys.path (its `src` package would shadow the reference's)
    spec = importlib.util.spec_from_file_location('oracle_torch_towers', OUT.parent.parent / 'oracle' / 'torch_towers.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fill_deterministic = mod.fill_deterministic
    orig18 = torchvision.models.resnet18
    torchvision.models.resnet18 = lambda pretrained=True, **k: orig18(weights=None)
    import src.networks.models.pcme as ref_pcme
    cfg = BertConfig(num_hidden_layers=2)

    class FakeTokenizer:
        """Captions are strings of space-separated ids; pads with 0 like BertTokenizer(padding=True)."""

        def __call__(self, captions, padding=True, return_tensors='pt'):
            rows = [[int(t) for t in c.split()] for c in captions]
            width = max(len(r) for r in rows)
            ids = torch.tensor([r + [0] * (width - len(r)) for r in rows])
            mask = torch.tensor([[1] * len(r) + [0] * (width - len(r)) for r in rows])
            return {'input_ids': ids, 'token_type_ids': torch.zeros_like(ids), 'attention_mask': mask}

    ref_pcme.BertModel.from_pretrained = staticmethod(lambda name: BertModel(cfg))
    ref_pcme.BertTokenizer.from_pretrained = staticmethod(lambda name: FakeTokenizer())
    model = ref_pcme.PCME(None, Munch(embed_dim=64, cnn_type='resnet18', not_bert=False, n_samples_inference=0), False)
    fill_deterministic(model, seed=31)
    g = torch.Generator().manual_seed(32)
    images = torch.randn(2, 3, 224, 224, generator=g)
    ids = torch.randint(1000, 30000, (2, 8), generator=g)
    ids[:, 0] = 101
    captions = (' '.join(str(int(t)) for t in ids[0]), ' '.join(str(int(t)) for t in ids[1][:5]))
    model.eval()
    with torch.no_grad():
        o = model(images, None, captions, None)
    model.img_enc.train()
    with torch.no_grad():
        tr = model.img_enc(images)['embedding']
    out = {'images': images.numpy(), 'ids': ids.numpy(), 'len1': np.int64(5),
           'eval_image_features': o['image_features'].numpy(), 'eval_caption_features': o['caption_features'].numpy(),
           'train_image_embedding': tr.numpy(), 'none_keys': np.array(sorted(k for k, v in o.items() if v is None))}
    # unimodal image client
    import src.networks.resnet_client as rc
    rc.model_zoo.load_url = lambda url: {}
    client = rc.resnet18_client(pretrained=False, num_class=10, is_train=True, scale=128, embed_dim=64)
    fill_deterministic(client, seed=33)
    with torch.no_grad():
        client.linear.weight.mul_(0.05)
    small = torch.randn(4, 3, 64, 64, generator=g)
    client.train()
    x1, x2, w1, w2 = client(small)
    client.phase, client.is_train = 'extract_conv_feature', False
    with torch.no_grad():
        emb = client(small)
    out.update({'client_images': small.numpy(), 'client_x1': x1.detach().numpy(), 'client_w_min': np.float64(w1.min().item()),
                'client_embedding': emb.numpy()})
    np.savez_compressed(OUT / 'towers.npz', **out)


def case_text_towers():
    """The reference's own GRU text towers - caption_encoder.EncoderText (multimodal client) and
    language_model.EncoderText (unimodal text client) - on deterministic weights and ragged, length-sorted captions:
    pins RefGRUEncoderText / RefTextClient of oracle/torch_towers.py (forward values and parameter gradients)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('oracle_torch_towers', OUT.parent.parent / 'oracle' / 'torch_towers.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    fill_deterministic = mod.fill_deterministic
    import src.networks.models.caption_encoder as ce
    vocab = 500
    word2idx = {f'w{i}': i for i in range(vocab)}
    opt = Munch(wemb_type=None, word_dim=300, embed_dim=64, cache_dir='')
    enc = ce.EncoderText(word2idx, opt, False)
    fill_deterministic(enc, seed=41)
    g = torch.Generator().manual_seed(42)
    lengths = torch.tensor([9, 9, 7, 4, 2, 1])
    x = torch.randint(4, vocab, (6, 9), generator=g)
    for i, n in enumerate(lengths.tolist()):
        x[i, n:] = 0
    coef = torch.randn(6, 64, generator=g)
    enc.train()
    emb = enc(x, lengths)['embedding']
    (emb * coef).sum().backward()
    names = ['embed.weight', 'rnn.weight_ih_l0', 'rnn.weight_hh_l0', 'rnn.bias_ih_l0', 'rnn.bias_hh_l0',
             'rnn.weight_ih_l0_reverse', 'rnn.weight_hh_l0_reverse', 'rnn.bias_ih_l0_reverse', 'rnn.bias_hh_l0_reverse',
             'pie_net.attention.w_1.weight', 'pie_net.attention.w_2.weight', 'pie_net.fc.weight', 'pie_net.fc.bias',
             'pie_net.layer_norm.weight', 'pie_net.layer_norm.bias']
    params = dict(enc.named_parameters())
    out = {'x': x.numpy(), 'lengths': lengths.numpy(), 'coef': coef.numpy(), 'mm_embedding': emb.detach().numpy()}
    for nme in names:
        gr = params[nme].grad
        out['mm_grad.' + nme] = (gr if gr is not None else torch.zeros_like(params[nme])).numpy()
    # unimodal text client (opens src/datasets/vocabs/coco_vocab.pkl relative to the reference root: 11755 words)
    import src.networks.language_model as lm
    client = lm.EncoderText(wemb_type=None, word_dim=300, embed_dim=64, num_class=4, scale=128)
    fill_deterministic(client, seed=43)
    xv = x.clone()
    client.train()
    x1, x2, w1, w2 = client(xv, lengths)
    labels = torch.tensor([0, 3, 1, 2, 2, 0])
    onehot = torch.nn.functional.one_hot(labels, 4).float()
    loss = torch.nn.functional.cross_entropy(x1 - 4.0 * onehot, labels) + \
        0.5 * torch.nn.functional.cross_entropy(w1 @ w1.t(), torch.arange(4))        # ClientTrainer.py:344-355
    loss.backward()
    cparams = dict(client.named_parameters())
    out.update({'uni_x1': x1.detach().numpy(), 'uni_x2': x2.detach().numpy(), 'uni_loss': np.float64(loss.item()),
                'uni_labels': labels.numpy(), 'uni_vocab': np.int64(client.embed.weight.shape[0]),
                'uni_w_min': np.float64(min(w1.min().item(), w2.min().item()))})
    for nme in ['rnn.weight_hh_l0', 'rnn.weight_ih_l0_reverse', 'pie_net.fc.weight', 'class_fc.weight', 'class_fc.bias']:
        out['uni_grad.' + nme] = cparams[nme].grad.numpy()
    out['uni_grad_embed_rows'] = cparams['embed.weight'].grad[:vocab].numpy()
    client.is_train = False
    with torch.no_grad():
        out['uni_embedding'] = client(xv, lengths).numpy()
    np.savez_compressed(OUT / 'text_towers.npz', **out)


def case_match_prob():
    """criterion.match_prob (probemb.py:210-219) the way MatchingProbModule calls it (eval_coco.py:66-69): one query
    [1, K, D] broadcast against the gallery [Ng, K, D], K = n_embeddings; plus the same-N and 2-D call forms."""
    from criterions.probemb import MCSoftContrastiveLoss
    out = {}
    for tag, (nq, ng, k, d, s, b) in {'a': (1, 12, 3, 16, 15.0, 15.0), 'b': (7, 7, 2, 8, 3.0, 2.0),
                                       'c': (9, 1, 1, 32, 5.0, 4.0)}.items():
        g = torch.Generator().manual_seed(700 + ng)
        q = unit(torch.randn(nq, k, d, generator=g, dtype=torch.float64))
        gal = unit(torch.randn(ng, k, d, generator=g, dtype=torch.float64))
        crit = MCSoftContrastiveLoss(Munch(init_shift=b, init_negative_scale=s, num_samples=k)).double()
        with torch.no_grad():
            prob = crit.match_prob(q, gal, None, None)
        out.update({f'{tag}_q': q.numpy(), f'{tag}_g': gal.numpy(), f'{tag}_shift': np.float64(b),
                    f'{tag}_scale': np.float64(s), f'{tag}_prob': prob.numpy()})
    g = torch.Generator().manual_seed(77)
    q2, g2 = unit(torch.randn(6, 16, generator=g, dtype=torch.float64)), unit(torch.randn(6, 16, generator=g, dtype=torch.float64))
    crit = MCSoftContrastiveLoss(Munch(init_shift=1.5, init_negative_scale=2.5, num_samples=1)).double()
    with torch.no_grad():
        out.update({'d_q': q2.numpy(), 'd_g': g2.numpy(), 'd_shift': np.float64(1.5), 'd_scale': np.float64(2.5),
                    'd_prob': crit.match_prob(q2, g2, None, None).numpy()})
    np.savez_compressed(OUT / 'match_prob.npz', **out)


def case_uni_lr():
    """ClientTrainer.lr_scheduler (ClientTrainer.py:291-302) called the way run() does (:198, with the round number a
    selected client is handed, MMFL.py:229): learning rate of the SGD optimizer after every call, for a client that
    trains every round and for one that is selected only now and then (the decay flags fire once each, late)."""
    import io
    import contextlib
    import src.algorithms.ClientTrainer as CT
    out = {}
    for tag, epochs in {'every': list(range(30)), 'sparse': [0, 3, 14, 16, 23, 25, 29], 'late': [27, 28],
                        'short': list(range(10))}.items():
        tr = CT.ClientTrainer.__new__(CT.ClientTrainer)
        tr.num_epochs = 10 if tag == 'short' else 30
        tr.init_lr, tr.decay_rate, tr.decay_time = 1e-4, 0.1, [False, False]
        tr.optimizer = torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=tr.init_lr, momentum=0.9, weight_decay=5e-5)
        lrs = []
        for e in epochs:
            with contextlib.redirect_stdout(io.StringIO()):
                tr.lr_scheduler(e)
            lrs.append(tr.optimizer.param_groups[0]['lr'])
        out[f'{tag}_epochs'] = np.asarray(epochs, dtype=np.int64)
        out[f'{tag}_num_epochs'] = np.int64(tr.num_epochs)
        out[f'{tag}_lr'] = np.asarray(lrs, dtype=np.float64)
    np.savez_compressed(OUT / 'uni_lr.npz', **out)


CASES = {'towers': case_towers, 'text_towers': case_text_towers, 'pcme': case_pcme, 'mm_contrast': case_mm_contrast, 'uni_contrast': case_uni_contrast,
         'recall': case_recall, 'partition': case_partition, 'conw': case_conw,
         'match_prob': case_match_prob, 'uni_lr': case_uni_lr}

if __name__ == '__main__':
    install_shims()
    torch.manual_seed(0)
    names = sys.argv[1:] or list(CASES)
    for nme in names:
        print('golden:', nme, flush=True)
        CASES[nme]()
    print('done')
