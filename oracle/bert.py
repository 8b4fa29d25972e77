"""Oracle: BERT text / fusion encoder with cross-attention, fp32 CPU, functional over a state_dict.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates, op for op,
  * BertEmbeddings                     model/bert.py:101-149 (token_type 0, absolute positions 0..S-1; dropout off)
  * BertSelfAttention (self and cross) model/bert.py:184-283 (scores / sqrt(d) AFTER the matmul :257, + additive mask :260)
  * BertSelfOutput / BertOutput        model/bert.py:286-297, 364-375 (post-LN: LN(dense(x) + residual))
  * BertIntermediate                   model/bert.py:349-361 (exact-erf GELU)
  * BertLayer                          model/bert.py:393-461 (self -> cross if encoder_hidden_states -> FFN)
  * masks                              model/bert.py:697-781: (1-m) * -10000, 2-D -> (b,1,1,S), 3-D -> (b,1,S,S); encoder
                                       mask through HF invert_attention_mask: (1-m) * finfo.min (:872)
  * BertLMPredictionHead + CE          model/bert.py:575-609, 1084-1090 (ignore_index -100, mean)
`p` maps the reference's own state_dict key names (prefix e.g. 'multimodal_encoder.') to tensors.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def drop_mult(p, seed, offset, shape):
    """Host restatement of the product's counter-based dropout (mico_b200/csrc/common.cuh drop_mult): element i of a tensor
    gets the multiplier 1/(1-p) if splitmix64(seed, offset + i) >> 40 scaled to [0,1) is >= p, else 0.  The reference draws
    its masks from torch's Philox stream (bert.py:93,243: nn.Dropout), which no other implementation can reproduce; tests
    therefore feed THESE masks to the fp32 restatement and compare."""
    n = int(np.prod(shape))
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(offset)) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
    u = (z >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    m = np.where(u >= np.float32(p), np.float32(1.0 / (1.0 - p)), np.float32(0.0))
    return torch.from_numpy(m.reshape(shape))


def attn_drop_mult(p, seed, B, H, Sq, Sk):
    """Host restatement of the attention kernels' probability-dropout mask (mico_b200/csrc/common.cuh drop_row_key /
    drop_pair_bits): score row r = (b*H + h)*Sq + i gets the 32-bit key low32(splitmix64(seed + r * golden)); keys 2t, 2t+1
    of the row share x = (key ^ (t * 0x9E3779B1)) * 0x85EBCA6B, x ^= x >> 15, whose low / high half is their 16-bit uniform; an element is kept
    (multiplier 1/(1-p)) when uniform16 >= round(p * 65536).  Returns a (B, H, Sq, Sk) fp32 tensor."""
    rows = B * H * Sq
    with np.errstate(over="ignore"):
        z = np.arange(rows, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
        z ^= z >> np.uint64(30)
        z *= np.uint64(0xBF58476D1CE4E5B9)
        z ^= z >> np.uint64(27)
        z *= np.uint64(0x94D049BB133111EB)
        z ^= z >> np.uint64(31)
        key = (z & np.uint64(0xFFFFFFFF)).astype(np.uint32)[:, None]
        j = np.arange(Sk, dtype=np.uint32)[None, :]
        x = key ^ ((j >> np.uint32(1)) * np.uint32(0x9E3779B1))
        x *= np.uint32(0x85EBCA6B)
        x ^= x >> np.uint32(15)
    u = np.where((j & np.uint32(1)) == 1, x >> np.uint32(16), x & np.uint32(0xFFFF))
    thresh = np.uint32(int(np.float32(p) * np.float32(65536.0) + np.float32(0.5)))
    m = np.where(u >= thresh, np.float32(1.0 / (1.0 - p)), np.float32(0.0)).astype(np.float32)
    return torch.from_numpy(m.reshape(B, H, Sq, Sk))


def _site(li, k):
    return (0 if li < 0 else 3 * li + k) << 40


def _lin(p, k, x):
    return F.linear(x, p[k + ".weight"], p[k + ".bias"])


def _ln(p, k, x, eps):
    return F.layer_norm(x, (x.shape[-1],), p[k + ".weight"], p[k + ".bias"], eps)


def embeddings(p, pre, ids, eps):
    S = ids.shape[1]
    x = p[pre + "word_embeddings.weight"][ids] + p[pre + "token_type_embeddings.weight"][0]
    x = x + p[pre + "position_embeddings.weight"][:S]
    return _ln(p, pre + "LayerNorm", x, eps)


def attention(p, pre, x, kv_src, add_mask, heads, adrop=None):
    b, S, D = x.shape
    d = D // heads
    q = _lin(p, pre + "self.query", x).view(b, S, heads, d).transpose(1, 2)
    k = _lin(p, pre + "self.key", kv_src).view(b, -1, heads, d).transpose(1, 2)
    v = _lin(p, pre + "self.value", kv_src).view(b, -1, heads, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(d)
    if add_mask is not None:
        s = s + add_mask
    pr = s.softmax(-1)
    if adrop is not None:          # attention-probability dropout (bert.py:243-247) with the product's mask
        pr = pr * attn_drop_mult(adrop[0], adrop[1], *pr.shape)
    ctx = (pr @ v).transpose(1, 2).reshape(b, S, D)
    return ctx


def layer(p, pre, h, mask_self, enc, mask_enc, heads, eps, li=0, drop=None):
    ph, pa, seed = drop if drop is not None else (0.0, 0.0, 0)

    def hd(x, k):      # hidden dropout after a dense layer (bert.py:294, 372)
        return x * drop_mult(ph, seed, _site(li, k), tuple(x.shape)) if ph > 0 else x

    ctx = attention(p, pre + "attention.", h, h, mask_self, heads, (pa, seed + 2 * li + 1) if pa > 0 else None)
    h = _ln(p, pre + "attention.output.LayerNorm", hd(_lin(p, pre + "attention.output.dense", ctx), 1) + h, eps)
    if enc is not None:
        ctx = attention(p, pre + "crossattention.", h, enc, mask_enc, heads, (pa, seed + 2 * li + 2) if pa > 0 else None)
        h = _ln(p, pre + "crossattention.output.LayerNorm", hd(_lin(p, pre + "crossattention.output.dense", ctx), 2) + h, eps)
    a = F.gelu(_lin(p, pre + "intermediate.dense", h))
    return _ln(p, pre + "output.LayerNorm", hd(_lin(p, pre + "output.dense", a), 3) + h, eps)


def extended_mask(attention_mask):
    m = attention_mask.float()
    m = m[:, None, :, :] if m.dim() == 3 else m[:, None, None, :]
    return (1.0 - m) * -10000.0


def bert_model(p, ids, attention_mask=None, enc=None, enc_mask=None, prefix="bert.", layers=12, heads=12, eps=1e-12,
               drop=None):
    b, S = ids.shape
    if attention_mask is None:
        attention_mask = torch.ones(b, S, device=ids.device)
    ms = extended_mask(attention_mask)
    me = None
    if enc is not None and enc_mask is not None:
        me = (1.0 - enc_mask.float()[:, None, None, :]) * torch.finfo(torch.float32).min
    h = embeddings(p, prefix + "embeddings.", ids, eps)
    if drop is not None and drop[0] > 0:
        h = h * drop_mult(drop[0], drop[2], _site(-1, 0), tuple(h.shape))
    for i in range(layers):
        h = layer(p, f"{prefix}encoder.layer.{i}.", h, ms, enc, me, heads, eps, i, drop)
    return h


def lm_head(p, h, prefix="cls.predictions.", eps=1e-12):
    t = _ln(p, prefix + "transform.LayerNorm", F.gelu(_lin(p, prefix + "transform.dense", h)), eps)
    return F.linear(t, p[prefix + "decoder.weight"], p[prefix + "bias"])


def masked_lm(p, ids, attention_mask=None, enc=None, enc_mask=None, labels=None, layers=12, heads=12, eps=1e-12, prefix="",
              drop=None):
    seq = bert_model(p, ids, attention_mask, enc, enc_mask, prefix + "bert.", layers, heads, eps, drop)
    logits = lm_head(p, seq, prefix + "cls.predictions.", eps)
    loss = None
    if labels is not None:
        loss = F.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1))
    return loss, logits, seq


def init_params(seed=0, prefix="", layers=12, hidden=768, heads=12, inter=3072, vocab=30522, max_pos=512, types=2, std=0.02):
    """Random-init state_dict of BertForMaskedLM as the reference builds it (bert.py:624-637 _init_weights: Linear / Embedding
    weights ~ N(0, 0.02), biases 0, LayerNorm 1 / 0; decoder tied to the word embeddings as under transformers 4.31), for
    the CPU baseline legs of bench.py -- no checkpoint is available offline."""
    g = torch.Generator().manual_seed(seed)
    n = lambda *s: std * torch.randn(*s, generator=g)
    p = {}

    def lin(k, o, i):
        p[k + ".weight"], p[k + ".bias"] = n(o, i), torch.zeros(o)

    def ln(k):
        p[k + ".weight"], p[k + ".bias"] = torch.ones(hidden), torch.zeros(hidden)
    e = prefix + "bert.embeddings."
    p[e + "word_embeddings.weight"] = n(vocab, hidden)
    p[e + "word_embeddings.weight"][0].zero_()
    p[e + "position_embeddings.weight"] = n(max_pos, hidden)
    p[e + "token_type_embeddings.weight"] = n(types, hidden)
    ln(e + "LayerNorm")
    for i in range(layers):
        l = f"{prefix}bert.encoder.layer.{i}."
        for att in ("attention.", "crossattention."):
            for k in ("query", "key", "value"):
                lin(l + att + "self." + k, hidden, hidden)
            lin(l + att + "output.dense", hidden, hidden)
            ln(l + att + "output.LayerNorm")
        lin(l + "intermediate.dense", inter, hidden)
        lin(l + "output.dense", hidden, inter)
        ln(l + "output.LayerNorm")
    c = prefix + "cls.predictions."
    lin(c + "transform.dense", hidden, hidden)
    ln(c + "transform.LayerNorm")
    p[c + "decoder.weight"] = p[e + "word_embeddings.weight"]
    p[c + "bias"] = torch.zeros(vocab)
    return p
