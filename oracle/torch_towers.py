"""Torch restatement of the reference's encoder towers - TEST INFRASTRUCTURE ONLY (see oracle/creamfl_oracle.py).

The reference builds its towers from third-party modules (torchvision ResNet, `torchvision==0.11.1` in its
requirements.txt:15; HF `BertModel`, unpinned) plus ~60 lines of its own glue.  This file restates the glue on top of
the torchvision / transformers versions present in the image, so that the CUDA towers (creamfl_b200/towers.py) can
be compared with it on identical weights and inputs, in fp32/fp64 torch.

Pinning: tests/golden/make_golden.py (case `towers`) runs the reference's own `PCME` (with import shims) on
deterministic weights (`fill_deterministic`) and stores its outputs; tests/test_oracle_golden.py checks this
restatement against them.
"""
from __future__ import annotations

import zlib

import torch
import torch.nn as nn
import torch.nn.functional as F


def l2_normalize(x, axis=-1):
    """reference src/utils/tensor_utils.py:25-27."""
    return F.normalize(x, p=2, dim=axis)


class RefSelfAttnPool(nn.Module):
    """reference src/networks/models/pie_model.py:11-40 (mask=None path)."""

    def __init__(self, n_head, d_in, d_hidden):
        super().__init__()
        self.w_1 = nn.Linear(d_in, d_hidden, bias=False)
        self.w_2 = nn.Linear(d_hidden, n_head, bias=False)

    def forward(self, x):
        attn = self.w_2(torch.tanh(self.w_1(x)))          # pie_model.py:30
        attn = torch.softmax(attn, dim=1)                 # :35
        output = torch.bmm(attn.transpose(1, 2), x)       # :37
        if output.shape[1] == 1:
            output = output.squeeze(1)
        return output, attn


class RefPIENet(nn.Module):
    """reference pie_model.py:43-67 with dropout = 0."""

    def __init__(self, n_embeds, d_in, d_out, d_h):
        super().__init__()
        self.attention = RefSelfAttnPool(n_embeds, d_in, d_h)
        self.fc = nn.Linear(d_in, d_out)
        self.layer_norm = nn.LayerNorm(d_out)

    def forward(self, out, x):
        residual, attn = self.attention(x)
        residual = torch.sigmoid(self.fc(residual))       # :63
        out = self.layer_norm(out + residual)             # :66
        return out, attn, residual


class RefEncoderImage(nn.Module):
    """reference src/networks/models/image_encoder.py:17-71 (mlp_local False, random init instead of pretrained)."""

    def __init__(self, cnn_type, embed_dim):
        super().__init__()
        import torchvision
        self.cnn = getattr(torchvision.models, cnn_type)(weights=None)
        cnn_dim = self.cnn_dim = self.cnn.fc.in_features
        self.avgpool = self.cnn.avgpool
        self.cnn.avgpool = nn.Sequential()
        self.fc = nn.Linear(cnn_dim, embed_dim)
        self.cnn.fc = nn.Sequential()
        self.pie_net = RefPIENet(1, cnn_dim, embed_dim, cnn_dim // 2)

    def forward(self, images):
        out_7x7 = self.cnn(images).view(-1, self.cnn_dim, 7, 7)          # :55
        pooled = self.avgpool(out_7x7).view(-1, self.cnn_dim)            # :56
        out = self.fc(pooled)                                            # :57
        out_7x7 = out_7x7.view(-1, self.cnn_dim, 7 * 7)                  # :60
        out, attn, residual = self.pie_net(out, out_7x7.transpose(1, 2)) # :62
        return {'embedding': l2_normalize(out)}                          # :67-71


class RefPCME(nn.Module):
    """reference src/networks/models/pcme.py:15-57 with the BERT tower and pre-tokenised inputs (the reference
    tokenises strings on the host inside forward, pcme.py:40-42)."""

    def __init__(self, cnn_type, embed_dim, bert_config=None, bert_dropout=0.0):
        super().__init__()
        from transformers import BertConfig, BertModel
        cfg = bert_config or BertConfig()
        # bert_dropout = 0: dropout frozen.  bert_dropout = 0.1 is what the reference trains with
        # (HF defaults, pcme.py:31); run the forward under `frozen_dropout` to feed it given keep masks.
        cfg.hidden_dropout_prob = float(bert_dropout)
        cfg.attention_probs_dropout_prob = float(bert_dropout)
        if bert_dropout > 0:
            cfg._attn_implementation = 'eager'    # the eager path calls F.dropout on the probabilities (patchable)
        self.embed_dim = embed_dim
        self.img_enc = RefEncoderImage(cnn_type, embed_dim)
        self.txt_enc = BertModel(cfg)
        self.linear = nn.Linear(cfg.hidden_size, embed_dim)

    def forward(self, images, input_ids, attention_mask, token_type_ids=None):
        image_output = self.img_enc(images)
        caption_output = self.txt_enc(input_ids=input_ids, attention_mask=attention_mask,
                                      token_type_ids=token_type_ids)
        cap = l2_normalize(self.linear(caption_output['last_hidden_state'][:, 0, :]))     # pcme.py:44
        return {'image_features': image_output['embedding'], 'caption_features': cap}


class frozen_dropout:
    """Context manager that replaces `torch.nn.functional.dropout` (reached by nn.Dropout and by HF's eager
    attention) with a multiplication by caller-supplied keep masks: `provider(call_index, shape)` returns the keep
    mask (0/1, broadcastable to `shape`) of the call_index-th train-mode dropout of the forward.  For HF BertModel
    the call order is: embeddings (0), then per layer attention probabilities (1 + 3l), attention output (2 + 3l),
    FFN output (3 + 3l) - the site numbering of include/creamfl_b200.h."""

    def __init__(self, provider):
        self.provider, self.calls, self._orig = provider, 0, None

    def __enter__(self):
        self._orig = F.dropout

        def masked(input, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return input
            keep = self.provider(self.calls, tuple(input.shape)).to(input.dtype).to(input.device)
            self.calls += 1
            return input * keep / (1.0 - p)
        torch.nn.functional.dropout = masked
        return self

    def __exit__(self, *exc):
        torch.nn.functional.dropout = self._orig
        return False


def fill_deterministic(module: nn.Module, seed: int = 0) -> None:
    """Overwrite every parameter / buffer with values that depend only on (name, shape, seed), so that two
    independently constructed models (the reference's and a restatement) hold identical weights without shipping a
    checkpoint.  Scales keep activations O(1) through deep stacks."""
    with torch.no_grad():
        for name, t in sorted(module.state_dict().items()):
            if not t.is_floating_point():
                t.zero_()
                continue
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7fffffff)
            shape = tuple(t.shape)
            if name.endswith('running_var'):
                v = 0.5 + torch.rand(shape, generator=g)
            elif name.endswith('running_mean'):
                v = 0.1 * torch.randn(shape, generator=g)
            elif t.dim() == 1:
                if 'bn' in name or 'norm' in name.lower() or 'downsample.1' in name:
                    v = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith('weight') else \
                        0.1 * torch.randn(shape, generator=g)
                else:
                    v = 0.05 * torch.randn(shape, generator=g)
            elif t.dim() == 4:
                fan_in = shape[1] * shape[2] * shape[3]
                v = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
            elif 'embeddings' in name:
                v = 0.05 * torch.randn(shape, generator=g)
            else:
                v = torch.randn(shape, generator=g) * (1.0 / shape[-1]) ** 0.5
            t.copy_(v.to(t.dtype))


class RefImageClient(nn.Module):
    """reference src/networks/resnet_client.py:100-201 (ResNet with BasicBlock [2,2,2,2] = resnet18_client) on
    torchvision's resnet18 trunk (same graph and parameter names)."""

    def __init__(self, num_class=100, embed_dim=256, scale=128, is_train=True, phase='none'):
        super().__init__()
        import torchvision
        tv = torchvision.models.resnet18(weights=None)
        self.conv1, self.bn1, self.relu, self.maxpool = tv.conv1, tv.bn1, nn.ReLU(inplace=False), tv.maxpool
        self.layer1, self.layer2, self.layer3, self.layer4 = tv.layer1, tv.layer2, tv.layer3, tv.layer4
        self.avg_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.embed_dim = embed_dim
        if embed_dim != 512:
            self.linear = nn.Linear(512, embed_dim)
        self.class_fc_2 = nn.Linear(embed_dim, num_class)
        self.class_fc_22 = nn.Linear(embed_dim, 80)
        self.is_train, self.scale, self.phase = bool(is_train), int(scale), str(phase)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))            # resnet_client.py:164-167
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        x = self.avg_pool(x).flatten(1) * self.scale                      # :176-179
        if self.embed_dim != 512:
            x = self.linear(x)
        if self.phase == 'extract_conv_feature':
            return F.normalize(x, p=2, dim=1)                             # :188
        if self.is_train:
            w2 = self.relu(self.class_fc_2.weight)                        # :193-197
            self.class_fc_2.weight.data = w2
            w22 = self.relu(self.class_fc_22.weight)
            self.class_fc_22.weight.data = w22
            return self.class_fc_2(x), self.class_fc_22(x), w2, w22
        return x


def ref_unimodal_supervised_loss(model, inputs, labels, num_class, inter_distance=4.0):
    """reference src/algorithms/ClientTrainer.py:344-355."""
    fvec, _, class_weight, _ = model(inputs)
    onehot = F.one_hot(labels, num_class).to(fvec.dtype)
    fvec = fvec - inter_distance * onehot
    loss = F.cross_entropy(fvec, labels)
    center = F.cross_entropy(class_weight @ class_weight.t(), torch.arange(num_class, device=fvec.device))
    return 0.5 * center + loss, fvec


# ===================================================================================================== GRU text towers
def get_pad_mask(max_length, lengths, set_pad_to_one=True):
    """reference src/networks/models/caption_encoder.py:19-26."""
    ind = torch.arange(0, max_length).unsqueeze(0)
    mask = (ind >= lengths.unsqueeze(1)) if set_pad_to_one else (ind < lengths.unsqueeze(1))
    return mask


class RefMaskedPIENet(nn.Module):
    """reference pie_model.PIENet with the pad-mask path (text variant: d_in = 300, d_h = 150; pie_model.py:28-40,
    61-67, dropout 0)."""

    def __init__(self, d_in, d_out, d_h):
        super().__init__()
        self.attention = nn.Module()
        self.attention.w_1 = nn.Linear(d_in, d_h, bias=False)
        self.attention.w_2 = nn.Linear(d_h, 1, bias=False)
        self.fc = nn.Linear(d_in, d_out)
        self.layer_norm = nn.LayerNorm(d_out)
        nn.init.xavier_uniform_(self.attention.w_1.weight)
        nn.init.xavier_uniform_(self.attention.w_2.weight)
        nn.init.xavier_uniform_(self.fc.weight)
        nn.init.constant_(self.fc.bias, 0.0)

    def forward(self, out, x, pad_mask):
        attn = self.attention.w_2(torch.tanh(self.attention.w_1(x)))            # pie_model.py:30
        attn = attn.masked_fill(pad_mask.unsqueeze(-1), float('-inf'))            # :31-34
        attn = torch.softmax(attn, dim=1)                                         # :35
        residual = torch.bmm(attn.transpose(1, 2), x).squeeze(1)                  # :37-39
        residual = torch.sigmoid(self.fc(residual))                               # :63
        return self.layer_norm(out + residual), attn, residual                    # :66


class RefGRUEncoderText(nn.Module):
    """reference src/networks/models/caption_encoder.py:29-116 (wemb_type None -> xavier init; mlp_local False)."""

    def __init__(self, vocab_size, word_dim, embed_dim):
        super().__init__()
        self.embed_dim = embed_dim
        self.embed = nn.Embedding(vocab_size, word_dim)
        self.rnn = nn.GRU(word_dim, embed_dim // 2, bidirectional=True, batch_first=True)
        self.pie_net = RefMaskedPIENet(word_dim, embed_dim, word_dim // 2)
        nn.init.xavier_uniform_(self.embed.weight)

    def trunk(self, x, lengths):
        from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
        lengths = lengths.cpu() if torch.is_tensor(lengths) else torch.as_tensor(lengths)
        wemb_out = self.embed(x)                                                  # caption_encoder.py:90
        packed = pack_padded_sequence(wemb_out, lengths, batch_first=True)        # :93
        rnn_out, _ = self.rnn(packed)                                             # :96
        padded, _ = pad_packed_sequence(rnn_out, batch_first=True, total_length=wemb_out.shape[1])   # :97
        idx = (lengths - 1).to(x.device).view(-1, 1, 1).expand(-1, 1, self.embed_dim)
        out = torch.gather(padded, 1, idx).squeeze(1)                             # :99-101
        pad_mask = get_pad_mask(wemb_out.shape[1], lengths, True).to(x.device)    # :105
        out, attn, residual = self.pie_net(out, wemb_out, pad_mask)               # :107
        return out

    def forward(self, x, lengths):
        return {'embedding': l2_normalize(self.trunk(x, lengths))}                # :109


class RefTextClient(RefGRUEncoderText):
    """reference src/networks/language_model.py:28-130 (vocabulary size passed in instead of the pickled vocab)."""

    def __init__(self, vocab_size=11755, word_dim=300, embed_dim=256, num_class=4, scale=128):
        super().__init__(vocab_size, word_dim, embed_dim)
        self.class_fc = nn.Linear(embed_dim, num_class)
        self.class_fc_2 = nn.Linear(embed_dim, 80)
        self.is_train, self.phase, self.scale = True, '', scale

    def forward(self, x, lengths):
        out = torch.relu(self.trunk(x, lengths) * self.scale)                     # language_model.py:111-112
        if self.is_train:
            w1 = torch.relu(self.class_fc.weight)                                 # :115-121
            self.class_fc.weight.data = w1.detach().clone()
            w2 = torch.relu(self.class_fc_2.weight)
            self.class_fc_2.weight.data = w2.detach().clone()
            return self.class_fc(out), self.class_fc_2(out), w1, w2
        return F.normalize(out, p=2, dim=1)                                       # :128


def ref_text_supervised_loss(model, captions, lengths, labels, num_class, inter_distance=4.0):
    """reference src/algorithms/ClientTrainer.py:335-355 for the text clients."""
    fvec, _, class_weight, _ = model(captions, lengths)
    onehot = F.one_hot(labels, num_class).to(fvec.dtype)
    fvec = fvec - inter_distance * onehot
    loss = F.cross_entropy(fvec, labels)
    center = F.cross_entropy(class_weight @ class_weight.t(), torch.arange(num_class, device=fvec.device))
    return 0.5 * center + loss, fvec
