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

import torch
import torch.nn.functional as F


def _lin(p, k, x):
    return F.linear(x, p[k + ".weight"], p[k + ".bias"])


def _ln(p, k, x, eps):
    return F.layer_norm(x, (x.shape[-1],), p[k + ".weight"], p[k + ".bias"], eps)


def embeddings(p, pre, ids, eps):
    S = ids.shape[1]
    x = p[pre + "word_embeddings.weight"][ids] + p[pre + "token_type_embeddings.weight"][0]
    x = x + p[pre + "position_embeddings.weight"][:S]
    return _ln(p, pre + "LayerNorm", x, eps)


def attention(p, pre, x, kv_src, add_mask, heads):
    b, S, D = x.shape
    d = D // heads
    q = _lin(p, pre + "self.query", x).view(b, S, heads, d).transpose(1, 2)
    k = _lin(p, pre + "self.key", kv_src).view(b, -1, heads, d).transpose(1, 2)
    v = _lin(p, pre + "self.value", kv_src).view(b, -1, heads, d).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(d)
    if add_mask is not None:
        s = s + add_mask
    ctx = (s.softmax(-1) @ v).transpose(1, 2).reshape(b, S, D)
    return ctx


def layer(p, pre, h, mask_self, enc, mask_enc, heads, eps):
    ctx = attention(p, pre + "attention.", h, h, mask_self, heads)
    h = _ln(p, pre + "attention.output.LayerNorm", _lin(p, pre + "attention.output.dense", ctx) + h, eps)
    if enc is not None:
        ctx = attention(p, pre + "crossattention.", h, enc, mask_enc, heads)
        h = _ln(p, pre + "crossattention.output.LayerNorm", _lin(p, pre + "crossattention.output.dense", ctx) + h, eps)
    a = F.gelu(_lin(p, pre + "intermediate.dense", h))
    return _ln(p, pre + "output.LayerNorm", _lin(p, pre + "output.dense", a) + h, eps)


def extended_mask(attention_mask):
    m = attention_mask.float()
    m = m[:, None, :, :] if m.dim() == 3 else m[:, None, None, :]
    return (1.0 - m) * -10000.0


def bert_model(p, ids, attention_mask=None, enc=None, enc_mask=None, prefix="bert.", layers=12, heads=12, eps=1e-12):
    b, S = ids.shape
    if attention_mask is None:
        attention_mask = torch.ones(b, S)
    ms = extended_mask(attention_mask)
    me = None
    if enc is not None and enc_mask is not None:
        me = (1.0 - enc_mask.float()[:, None, None, :]) * torch.finfo(torch.float32).min
    h = embeddings(p, prefix + "embeddings.", ids, eps)
    for i in range(layers):
        h = layer(p, f"{prefix}encoder.layer.{i}.", h, ms, enc, me, heads, eps)
    return h


def lm_head(p, h, prefix="cls.predictions.", eps=1e-12):
    t = _ln(p, prefix + "transform.LayerNorm", F.gelu(_lin(p, prefix + "transform.dense", h)), eps)
    return F.linear(t, p[prefix + "decoder.weight"], p[prefix + "bias"])


def masked_lm(p, ids, attention_mask=None, enc=None, enc_mask=None, labels=None, layers=12, heads=12, eps=1e-12, prefix=""):
    seq = bert_model(p, ids, attention_mask, enc, enc_mask, prefix + "bert.", layers, heads, eps)
    logits = lm_head(p, seq, prefix + "cls.predictions.", eps)
    loss = None
    if labels is not None:
        loss = F.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1))
    return loss, logits, seq
