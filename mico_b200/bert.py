"""BERT text / fusion encoder with cross-attention (bert-base-uncased-crossattn) on the sm_100a kernels.

Host-side mirror of the reference's ``BertForMaskedLM`` / ``BertModel`` (model/bert.py:81-1108) for the
configuration MiCo uses (model/bert-base-uncased-crossattn/config.json: 12 post-LN layers, hidden 768, 12 heads x 64,
FFN 3072, ``is_decoder`` + ``add_cross_attention``, absolute positions).  Same ``state_dict`` keys; same call
contract ``model(input_ids, attention_mask, encoder_hidden_states=, labels=)`` -> object with
``loss / logits / sequence_output`` (bert.py:1093-1097); ``model.bert(...)`` -> ``.last_hidden_state``.

The encoder stack is ONE autograd node (embeddings + all layers): forward and backward are explicit launch
sequences over the C-ABI kernels --
    K6  word + position + type gather, LayerNorm                          (bert.py:139-148)
    K3  fused q|k|v projection (three nn.Linear as one GEMM), dense, FFN   (bert.py:196-209, 293, 357, 370)
    K4  softmax(q k^T / sqrt(64) + mask) v, self and cross                 (bert.py:233-277; masks :697-781)
    K2  post-LN  LayerNorm(dense(x) + residual)                            (bert.py:286-297, 364-375)
the LM head (dense + GELU + LN + vocab GEMM) and the cross-entropy are separate small nodes (bert.py:592-609,
1084-1090).  Masks follow the reference exactly: 2-D ``(b,S)`` or 3-D ``(b,S,S)`` attention masks become additive
``(1-m) * -10000`` and there is NO automatic causal mask even though ``is_decoder`` is set (bert.py:716-763).

Dropout (hidden 0.1 after the embeddings and every dense output, attention-probability 0.1: bert.py:148, 243-247, 294,
372) runs in training mode with counter-based masks (splitmix64 of a per-call seed and the element index): the
attention kernels and `mico_dropout` regenerate the same mask in the backward pass, nothing is stored.  The masks are
not the reference's Philox stream (no implementation could be, across devices), so parity with dropout on is checked
against an fp32 reference driven by the same masks (tests/test_gpu_bert.py); eval mode and p = 0 are exact.
"""
import json
import math
import os

import torch
import torch.nn as nn

from . import ops
from .functional import cross_entropy
from .ops import ACT_GELU, ACT_GELU_SAVE_GRAD, ACT_MUL_AUX, BF16, F32, MicoError


class BertConfig:
    """The subset of transformers.BertConfig the path reads; from_pretrained reads the reference's config.json."""

    def __init__(self, **kw):
        self.vocab_size = kw.get("vocab_size", 30522)
        self.hidden_size = kw.get("hidden_size", 768)
        self.num_hidden_layers = kw.get("num_hidden_layers", 12)
        self.num_attention_heads = kw.get("num_attention_heads", 12)
        self.intermediate_size = kw.get("intermediate_size", 3072)
        self.hidden_act = kw.get("hidden_act", "gelu")
        self.hidden_dropout_prob = kw.get("hidden_dropout_prob", 0.1)
        self.attention_probs_dropout_prob = kw.get("attention_probs_dropout_prob", 0.1)
        self.max_position_embeddings = kw.get("max_position_embeddings", 512)
        self.type_vocab_size = kw.get("type_vocab_size", 2)
        self.initializer_range = kw.get("initializer_range", 0.02)
        self.layer_norm_eps = kw.get("layer_norm_eps", 1e-12)
        self.pad_token_id = kw.get("pad_token_id", 0)
        self.is_decoder = kw.get("is_decoder", True)
        self.add_cross_attention = kw.get("add_cross_attention", True)
        self.position_embedding_type = kw.get("position_embedding_type", "absolute")
        self.tie_word_embeddings = kw.get("tie_word_embeddings", True)
        if self.hidden_act != "gelu" or self.position_embedding_type != "absolute":
            raise NotImplementedError("only the bert-base-uncased-crossattn configuration family is supported")

    @classmethod
    def from_pretrained(cls, path):
        f = path if path.endswith(".json") else os.path.join(path, "config.json")
        with open(f) as fh:
            return cls(**json.load(fh))


class _Lin(nn.Module):
    def __init__(self, i, o, std):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i).normal_(0.0, std))
        self.bias = nn.Parameter(torch.zeros(o))


class _LN(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class _Emb(nn.Module):
    def __init__(self, n, d, std, padding_idx=None):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n, d).normal_(0.0, std))
        if padding_idx is not None:
            with torch.no_grad():
                self.weight[padding_idx].zero_()


class BertEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = _Emb(c.vocab_size, c.hidden_size, c.initializer_range, c.pad_token_id)
        self.position_embeddings = _Emb(c.max_position_embeddings, c.hidden_size, c.initializer_range)
        self.token_type_embeddings = _Emb(c.type_vocab_size, c.hidden_size, c.initializer_range)
        self.LayerNorm = _LN(c.hidden_size, c.layer_norm_eps)
        self.register_buffer("position_ids", torch.arange(c.max_position_embeddings).expand((1, -1)))


class _SelfAttn(nn.Module):
    def __init__(self, c):
        super().__init__()
        h, s = c.hidden_size, c.initializer_range
        self.query, self.key, self.value = _Lin(h, h, s), _Lin(h, h, s), _Lin(h, h, s)


class _SelfOut(nn.Module):
    def __init__(self, c, i):
        super().__init__()
        self.dense = _Lin(i, c.hidden_size, c.initializer_range)
        self.LayerNorm = _LN(c.hidden_size, c.layer_norm_eps)


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.self = _SelfAttn(c)
        self.output = _SelfOut(c, c.hidden_size)


class _Inter(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = _Lin(c.hidden_size, c.intermediate_size, c.initializer_range)


class BertLayer(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.attention = _Attn(c)
        if c.add_cross_attention:
            self.crossattention = _Attn(c)
        self.intermediate = _Inter(c)
        self.output = _SelfOut(c, c.intermediate_size)


class BertEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(c) for _ in range(c.num_hidden_layers)])
        self.gradient_checkpointing = False


class _Cache:
    """bf16 GEMM operands (q|k|v and k|v weights concatenated) refreshed when a source parameter changes."""

    def __init__(self):
        self._c = {}

    def cat_w(self, key, params):
        ver = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._c.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        rows = sum(p.shape[0] for p in params)
        buf = torch.empty((rows, params[0].shape[1]), device=params[0].device, dtype=BF16)
        r = 0
        for p in params:
            ops.cast_bf16(p.detach().contiguous(), dst=buf[r:r + p.shape[0]])
            r += p.shape[0]
        self._c[key] = (ver, buf)
        return buf

    def cat_b(self, key, params):
        ver = tuple((p.data_ptr(), p._version) for p in params)
        hit = self._c.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        buf = torch.cat([p.detach() for p in params])
        self._c[key] = (ver, buf)
        return buf


# flat parameter order: embeddings (word, pos, type, ln.w, ln.b) then per layer:
_SELF = ("q.w", "q.b", "k.w", "k.b", "v.w", "v.b", "o.w", "o.b", "ln.w", "ln.b")
_L_SQ, _L_SO, _L_SLN = 0, 6, 8
_L_CQ, _L_CO, _L_CLN = 10, 16, 18
_L_IW, _L_IB, _L_OW, _L_OB, _L_OLN = 20, 21, 22, 23, 24
_PER_LAYER = 26


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, keep, drop, ids, mask_self, enc, mask_enc, enc_index, *params):
        out, saved = model._launch_forward(ids, mask_self, enc, mask_enc, params, keep, drop, enc_index)
        ctx.model, ctx.saved, ctx.params = model, saved, params
        ctx.has_enc = enc is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        if ctx.saved is None:
            raise MicoError("BERT backward called but activations were not kept")
        denc, grads = ctx.model._launch_backward(dout, ctx.saved, ctx.params)
        ctx.saved = None
        return (None, None, None, None, None, denc if ctx.needs_input_grad[5] else None, None, None) + tuple(grads)


class _Out:
    """Attribute bag standing in for the reference's edict / BaseModelOutput (bert.py:1093-1097, 909-916)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def __getitem__(self, k):
        return self.__dict__[k] if isinstance(k, str) else list(self.__dict__.values())[k]


class _LazyLogitsOut(_Out):
    """Output of BertForMaskedLM.forward(labels=...) when the loss came from the chunked LM-head kernel path (K7): `.logits`
    runs the full LM head the first time a caller reads it (the training loops read `.loss` only)."""

    def __init__(self, make_logits, **kw):
        super().__init__(**kw)
        self.__dict__["_make_logits"] = make_logits

    def __getattr__(self, k):          # only reached when the attribute is not in __dict__ yet
        if k == "logits":
            self.__dict__["logits"] = self.__dict__["_make_logits"]()
            return self.__dict__["logits"]
        raise AttributeError(k)

    def __getitem__(self, k):
        if k == "logits" or k == 1:
            return self.logits
        return self.__dict__[k] if isinstance(k, str) else [self.__dict__["loss"], None, self.__dict__["sequence_output"]][k]


class BertModel(nn.Module):
    def __init__(self, config, add_pooling_layer=False):
        super().__init__()
        if add_pooling_layer:
            raise NotImplementedError("MiCo builds BertModel(add_pooling_layer=False) (bert.py:1034)")
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.pooler = None
        self._cache = _Cache()
        self._dropout_calls = 0
        self.dropout_seed = None       # set to pin the dropout masks of the next training-mode calls (tests)

    # ------------------------------------------------------------------ parameters in launch order
    def _flat_params(self):
        e = self.embeddings
        ps = [e.word_embeddings.weight, e.position_embeddings.weight, e.token_type_embeddings.weight, e.LayerNorm.weight,
              e.LayerNorm.bias]
        for l in self.encoder.layer:
            for att in (l.attention, getattr(l, "crossattention", None)):
                if att is None:
                    ps += [None] * 10
                    continue
                s, o = att.self, att.output
                ps += [s.query.weight, s.query.bias, s.key.weight, s.key.bias, s.value.weight, s.value.bias, o.dense.weight,
                       o.dense.bias, o.LayerNorm.weight, o.LayerNorm.bias]
            ps += [l.intermediate.dense.weight, l.intermediate.dense.bias, l.output.dense.weight, l.output.dense.bias,
                   l.output.LayerNorm.weight, l.output.LayerNorm.bias]
        return ps

    def invalidate_weight_cache(self):
        self._cache = _Cache()

    def _dropout_cfg(self):
        """(p_hidden, p_attention, seed) for a training-mode call, else None.  Masks are counter-based (splitmix64 of
        seed and element index): the backward pass regenerates them, and a test can reproduce them on the host."""
        c = self.config
        if not self.training or (c.hidden_dropout_prob <= 0 and c.attention_probs_dropout_prob <= 0):
            return None
        self._dropout_calls += 1
        seed = self.dropout_seed if self.dropout_seed is not None else \
            (torch.initial_seed() * 1000003 + self._dropout_calls * 7919) & (2 ** 62 - 1)
        return (float(c.hidden_dropout_prob), float(c.attention_probs_dropout_prob), int(seed))

    # ------------------------------------------------------------------ launch sequences
    @staticmethod
    def _site(li, k):
        """hidden-dropout site -> element-counter offset (sites: 0 embeddings; per layer 1 self-out, 2 cross-out, 3 ffn-out)"""
        return (0 if li < 0 else 3 * li + k) << 40

    def _launch_forward(self, ids, mask_self, enc, mask_enc, params, keep, drop=None, enc_index=None):
        c = self.config
        Dh, H = c.hidden_size, c.num_attention_heads
        d = Dh // H
        b, S = ids.shape
        M = b * S
        eps = c.layer_norm_eps
        scale = 1.0 / math.sqrt(d)
        cache = self._cache
        det = lambda i: params[i].detach()
        x0 = ops.embedding_gather(ids.reshape(-1).contiguous(), det(0), det(1), det(2), S)
        # LayerNorm output in both formats: bf16 = next GEMM operand, fp32 = residual of the next sub-layer
        hb, h, mean, rstd = ops.layernorm_fwd(x0, det(3), det(4), eps, out_bf16=True, out_f32=True, save_stats=keep)
        ph, pa, seed = drop if drop is not None else (0.0, 0.0, 0)
        if ph > 0:       # bert.py:148 embedding dropout
            h, hb = ops.dropout(h, ph, seed, self._site(-1, 0), out_f32=True, out_bf16=True)
        saved = dict(ids=ids, x0=x0, stats0=(mean, rstd), layers=[], b=b, S=S, mask_self=mask_self, mask_enc=mask_enc,
                     drop=drop) if keep else None

        def dense_res(x_b, w, bias, res, li, k):
            """LayerNorm input of a post-LN sub-layer: dropout(dense(x)) + residual (bert.py:293-296, 370-373)"""
            if ph > 0:
                y = ops.gemm(x_b, w, out_dtype=F32, bias=bias)
                return ops.dropout(y, ph, seed, self._site(li, k), res=res)[0]
            return ops.gemm(x_b, w, out_dtype=F32, bias=bias, residual=res)

        def adrop(li, cross):
            return (pa, seed + 2 * li + (2 if cross else 1)) if pa > 0 else None
        encb = None
        Sk = E = 0
        groups = None
        if enc is not None:
            # enc_index (int32 [b]): text sequence i cross-attends to encoder entry enc_index[i] -- several sequences
            # (ITM positive / negative-text, caption) share one sample's visual tokens, whose K / V projection then runs once
            E, Sk = enc.shape[0], enc.shape[1]
            if enc_index is None and E != b:
                raise MicoError("encoder_hidden_states batch differs from input_ids batch and no encoder_index was given")
            encb = ops.scale_cast_bf16(enc.reshape(E * Sk, Dh).contiguous().float())
            if keep:
                saved["encb"] = encb
                saved["E"] = E
                if enc_index is not None:
                    groups = ops.kv_groups(enc_index, E)
                saved["enc_index"], saved["groups"] = enc_index, groups

        def ln2(y, wi, bi):
            return ops.layernorm_fwd(y, det(wi), det(bi), eps, out_bf16=True, out_f32=True, save_stats=keep)

        for li in range(c.num_hidden_layers):
            base = 5 + li * _PER_LAYER
            P = lambda k: params[base + k]
            # ---- self attention
            wqkv = cache.cat_w(("sqkv", li), [P(0), P(2), P(4)])
            bqkv = cache.cat_b(("sqkvb", li), [P(1), P(3), P(5)])
            qkv = ops.gemm(hb, wqkv, bias=bqkv)
            q5 = qkv.view(b, S, 3, H, d)
            if isinstance(mask_self, list):
                # consecutive groups of sequences with their own mask layout (the grouped fusion-encoder call: ITM sequences
                # carry a per-key padding mask, which every query row shares, caption sequences a per-row causal mask): one
                # launch per group keeps the cheap broadcast mask path for the former
                ctx = torch.empty((b, S, H, d), device=ids.device, dtype=BF16)
                lse = []
                for si, (lo, hi, m) in enumerate(mask_self):
                    dr = adrop(li, False)
                    _, l_ = ops.attention_fwd(q5[lo:hi, :, 0], q5[lo:hi, :, 1], q5[lo:hi, :, 2], scale, mask=m, out=ctx[lo:hi],
                                              need_lse=keep, dropout=(dr[0], dr[1] + 7919 * si) if dr else None)
                    lse.append(l_)
            else:
                ctx, lse = ops.attention_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], scale, mask=mask_self, need_lse=keep,
                                             dropout=adrop(li, False))
            y1 = dense_res(ctx.view(M, Dh), cache.cat_w(("so", li), [P(6)]), P(7).detach(), h, li, 1)
            h1b, h1, m1, r1 = ln2(y1, base + 8, base + 9)
            rec = dict(hb=hb, qkv=qkv, ctx=ctx, lse=lse, y1=y1, st1=(m1, r1), h1b=h1b) if keep else None
            hb_in, h_in = h1b, h1
            # ---- cross attention
            if enc is not None:
                qc = ops.gemm(h1b, cache.cat_w(("cq", li), [P(10)]), bias=P(11).detach())
                wkv = cache.cat_w(("ckv", li), [P(12), P(14)])
                bkv = cache.cat_b(("ckvb", li), [P(13), P(15)])
                kv = ops.gemm(encb, wkv, bias=bkv)
                kv5 = kv.view(E, Sk, 2, H, d)
                ctx2, lse2 = ops.attention_fwd(qc.view(b, S, H, d), kv5[:, :, 0], kv5[:, :, 1], scale, mask=mask_enc,
                                               need_lse=keep, dropout=adrop(li, True), kv_index=enc_index)
                y2 = dense_res(ctx2.view(M, Dh), cache.cat_w(("co", li), [P(16)]), P(17).detach(), h1, li, 2)
                h2b, h2, m2, r2 = ln2(y2, base + 18, base + 19)
                if keep:
                    rec.update(qc=qc, kv=kv, ctx2=ctx2, lse2=lse2, y2=y2, st2=(m2, r2), h2b=h2b)
                hb_in, h_in = h2b, h2
            # ---- feed forward
            wi = cache.cat_w(("i", li), [P(20)])
            pre = torch.empty((M, wi.shape[0]), device=ids.device, dtype=BF16) if keep else None
            a = ops.gemm(hb_in, wi, bias=P(21).detach(), act=ACT_GELU_SAVE_GRAD, aux_out=pre)     # pre holds gelu'(x)
            y3 = dense_res(a, cache.cat_w(("o", li), [P(22)]), P(23).detach(), h_in, li, 3)
            hb, h, m3, r3 = ln2(y3, base + 24, base + 25)
            if keep:
                rec.update(pre=pre, a=a, y3=y3, st3=(m3, r3))
                saved["layers"].append(rec)
        return h.view(b, S, Dh), saved

    def _launch_backward(self, dout, saved, params):
        c = self.config
        Dh, H = c.hidden_size, c.num_attention_heads
        d = Dh // H
        b, S = saved["b"], saved["S"]
        M = b * S
        scale = 1.0 / math.sqrt(d)
        cache = self._cache
        dev = dout.device
        grads = [None] * len(params)
        det = lambda i: params[i].detach()

        def pg(i):
            g = torch.empty_like(params[i], dtype=F32)
            grads[i] = g
            return g

        encb = saved.get("encb")
        has_enc = encb is not None
        denc = torch.zeros((encb.shape[0], Dh), device=dev, dtype=F32) if has_enc else None
        E = saved.get("E", b)
        Sk = encb.shape[0] // E if has_enc else 0
        g32 = dout.contiguous().view(M, Dh).float()
        g16 = None
        drop = saved.get("drop")
        ph, pa, seed = drop if drop is not None else (0.0, 0.0, 0)

        def dmask(gb, li, k):
            """gradient entering a dense layer whose output was dropped out: the same mask again"""
            return ops.dropout(gb, ph, seed, self._site(li, k), out_f32=False, out_bf16=True)[1] if ph > 0 else gb

        def ln_bwd_dense(g32_, g16_, y, st, i_w, i_b, i_bias, li, k):
            """Backward of LN(dropout(dense(x)) + residual): (fp32 residual gradient, bf16 gradient of the dense output) and
            the dense layer's bias gradient -> pg(i_bias).  One launch when the LayerNorm kernel can carry the dropout mask and
            the column sums; else LayerNorm backward, then the mask, then the column sums."""
            if ops.layernorm_bwd_fuses_dropout(Dh, g32_):
                return ops.layernorm_bwd(g32_, y, *st, det(i_w), pg(i_w), pg(i_b), want_bf16=True, dy2=g16_,
                                         colsum_out=pg(i_bias), dropout=(ph, seed, self._site(li, k)) if ph > 0 else None)
            d32, d16 = ops.layernorm_bwd(g32_, y, *st, det(i_w), pg(i_w), pg(i_b), want_bf16=True, dy2=g16_)
            d16 = dmask(d16, li, k)
            ops.colsum(d16, out=pg(i_bias))
            return d32, d16

        def adrop(li, cross):
            return (pa, seed + 2 * li + (2 if cross else 1)) if pa > 0 else None
        for li in range(c.num_hidden_layers - 1, -1, -1):
            rec = saved["layers"].pop()
            base = 5 + li * _PER_LAYER
            P = lambda k: params[base + k]
            hb_in = rec["h2b"] if has_enc else rec["h1b"]
            # ---- FFN: h3 = LN(a Wo^T + bo + h_in)
            dy3, dy3b = ln_bwd_dense(g32, g16, rec["y3"], rec["st3"], base + 24, base + 25, base + 23, li, 3)
            ops.gemm(dy3b, rec["a"], a_mn=True, b_mn=True, out=pg(base + 22))
            dpre = ops.gemm(dy3b, cache.cat_w(("o", li), [P(22)]), b_mn=True, act=ACT_MUL_AUX, aux_in=rec["pre"])
            ops.gemm(dpre, hb_in, a_mn=True, b_mn=True, out=pg(base + 20))
            ops.colsum(dpre, out=pg(base + 21))
            g16 = ops.gemm(dpre, cache.cat_w(("i", li), [P(20)]), b_mn=True)
            g32 = dy3
            # ---- cross attention: h2 = LN(ctx2 Wo^T + bo + h1)
            if has_enc:
                dy2_, dy2b = ln_bwd_dense(g32, g16, rec["y2"], rec["st2"], base + 18, base + 19, base + 17, li, 2)
                ops.gemm(dy2b, rec["ctx2"].view(M, Dh), a_mn=True, b_mn=True, out=pg(base + 16))
                dctx = ops.gemm(dy2b, cache.cat_w(("co", li), [P(16)]), b_mn=True)
                kv5 = rec["kv"].view(E, Sk, 2, H, d)
                dqc = torch.empty_like(rec["qc"])
                dkv = torch.empty_like(rec["kv"])
                dkv5 = dkv.view(E, Sk, 2, H, d)
                ops.attention_bwd(rec["qc"].view(b, S, H, d), kv5[:, :, 0], kv5[:, :, 1], rec["ctx2"], rec["lse2"],
                                  dctx.view(b, S, H, d), scale, mask=saved["mask_enc"], dq=dqc.view(b, S, H, d),
                                  dk=dkv5[:, :, 0], dv=dkv5[:, :, 1], dropout=adrop(li, True),
                                  kv_index=saved.get("enc_index"), groups=saved.get("groups"))
                ops.gemm(dqc, rec["h1b"], a_mn=True, b_mn=True, out=pg(base + 10))
                ops.colsum(dqc, out=pg(base + 11))
                g16 = ops.gemm(dqc, cache.cat_w(("cq", li), [P(10)]), b_mn=True)
                gkv = torch.empty((2 * Dh, Dh), device=dev, dtype=F32)
                ops.gemm(dkv, encb, a_mn=True, b_mn=True, out=gkv)
                grads[base + 12], grads[base + 14] = gkv[:Dh], gkv[Dh:]
                bkv = ops.colsum(dkv)
                grads[base + 13], grads[base + 15] = bkv[:Dh], bkv[Dh:]
                ops.gemm(dkv, cache.cat_w(("ckv", li), [P(12), P(14)]), b_mn=True, out=denc, accumulate=True)
                g32 = dy2_
            # ---- self attention: h1 = LN(ctx Wo^T + bo + h)
            dy1, dy1b = ln_bwd_dense(g32, g16, rec["y1"], rec["st1"], base + 8, base + 9, base + 7, li, 1)
            ops.gemm(dy1b, rec["ctx"].view(M, Dh), a_mn=True, b_mn=True, out=pg(base + 6))
            dctx = ops.gemm(dy1b, cache.cat_w(("so", li), [P(6)]), b_mn=True)
            q5 = rec["qkv"].view(b, S, 3, H, d)
            dqkv = torch.empty_like(rec["qkv"])
            g5 = dqkv.view(b, S, 3, H, d)
            if isinstance(saved["mask_self"], list):
                d4 = dctx.view(b, S, H, d)
                for si, (lo, hi, m) in enumerate(saved["mask_self"]):
                    dr = adrop(li, False)
                    ops.attention_bwd(q5[lo:hi, :, 0], q5[lo:hi, :, 1], q5[lo:hi, :, 2], rec["ctx"][lo:hi], rec["lse"][si], d4[lo:hi],
                                      scale, mask=m, dq=g5[lo:hi, :, 0], dk=g5[lo:hi, :, 1], dv=g5[lo:hi, :, 2],
                                      dropout=(dr[0], dr[1] + 7919 * si) if dr else None)
            else:
                ops.attention_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], rec["ctx"], rec["lse"], dctx.view(b, S, H, d), scale,
                                  mask=saved["mask_self"], dq=g5[:, :, 0], dk=g5[:, :, 1], dv=g5[:, :, 2], dropout=adrop(li, False))
            gw = torch.empty((3 * Dh, Dh), device=dev, dtype=F32)
            ops.gemm(dqkv, rec["hb"], a_mn=True, b_mn=True, out=gw)
            grads[base + 0], grads[base + 2], grads[base + 4] = gw[:Dh], gw[Dh:2 * Dh], gw[2 * Dh:]
            gb = ops.colsum(dqkv)
            grads[base + 1], grads[base + 3], grads[base + 5] = gb[:Dh], gb[Dh:2 * Dh], gb[2 * Dh:]
            g16 = ops.gemm(dqkv, cache.cat_w(("sqkv", li), [P(0), P(2), P(4)]), b_mn=True)
            g32 = dy1
        # ---- embeddings: h0 = LN(word + type + pos)
        if ph > 0:       # embedding dropout: mask both halves of the incoming gradient
            g32 = ops.dropout(g32.contiguous(), ph, seed, self._site(-1, 0))[0]
            if g16 is not None:
                g16 = ops.dropout(g16, ph, seed, self._site(-1, 0), out_f32=False, out_bf16=True)[1]
        dx0, _ = ops.layernorm_bwd(g32, saved["x0"], *saved["stats0"], det(3), pg(3), pg(4), dy2=g16)
        gword = torch.zeros_like(params[0], dtype=F32)
        ops.embedding_scatter_add(dx0, saved["ids"].reshape(-1).contiguous(), gword)
        grads[0] = gword
        gpos = torch.zeros_like(params[1], dtype=F32)
        ops.batch_sum(dx0, b, out=gpos.view(-1)[:S * Dh])
        grads[1] = gpos
        gtype = torch.zeros_like(params[2], dtype=F32)
        ops.batch_sum(gpos[:S].contiguous(), S, out=gtype[0])       # token_type_ids are all zero (bert.py:121-127)
        grads[2] = gtype
        if has_enc:
            denc = denc.view(E, Sk, Dh)
        return denc, grads

    # ------------------------------------------------------------------ masks (bert.py:697-781, :872)
    @staticmethod
    def _additive(mask, neg):
        if mask is None:
            return None
        if mask.dim() not in (2, 3):
            raise ValueError(f"Wrong shape for attention_mask (shape {tuple(mask.shape)})")
        return ((1.0 - mask.float()) * neg).contiguous()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None, past_key_values=None,
                use_cache=None, output_attentions=None, output_hidden_states=None, return_dict=None, encoder_index=None):
        """encoder_index (not in the reference signature): int tensor [batch] -- sequence i cross-attends to
        encoder_hidden_states[encoder_index[i]], so that sequences sharing a sample's visual tokens share its K / V."""
        if input_ids is None or inputs_embeds is not None or token_type_ids is not None or position_ids is not None \
                or past_key_values is not None or head_mask is not None:
            raise NotImplementedError("BertModel: only (input_ids, attention_mask, encoder_hidden_states, "
                                      "encoder_attention_mask) are on the MiCo path")
        if not input_ids.is_cuda:
            raise MicoError("mico_b200 runs on CUDA (sm_100a) only")
        drop = self._dropout_cfg()
        b, S = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones((b, S), device=input_ids.device)
        if isinstance(attention_mask, (list, tuple)):
            # several masks, each for the next mask.shape[0] sequences (2-D and 3-D layouts may be mixed; not in the reference
            # signature: used by the grouped fusion-encoder call of mico_b200/train_step.py)
            mask_self, lo = [], 0
            for m in attention_mask:
                mask_self.append((lo, lo + m.shape[0], self._additive(m, -10000.0)))
                lo += m.shape[0]
            if lo != b:
                raise ValueError("attention_mask list does not cover the batch")
        else:
            mask_self = self._additive(attention_mask, -10000.0)
        mask_enc = None
        if encoder_hidden_states is not None and encoder_attention_mask is not None:
            mask_enc = self._additive(encoder_attention_mask, torch.finfo(torch.float32).min)   # invert_attention_mask
        flat = self._flat_params()
        live = [p for p in flat if p is not None]
        enc = encoder_hidden_states
        keep = torch.is_grad_enabled() and (any(p.requires_grad for p in live) or (enc is not None and enc.requires_grad))
        # placeholders keep the flat indexing when the config has no cross-attention
        args = [p if p is not None else torch.empty(0, device=input_ids.device) for p in flat]
        if encoder_index is not None:
            encoder_index = encoder_index.to(device=input_ids.device, dtype=torch.int32).contiguous()
        out = _EncoderFn.apply(self, keep, drop, input_ids.long(), mask_self, enc, mask_enc, encoder_index, *args)
        return _Out(last_hidden_state=out, pooler_output=None)


class _Transform(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.dense = _Lin(c.hidden_size, c.hidden_size, c.initializer_range)
        self.LayerNorm = _LN(c.hidden_size, c.layer_norm_eps)


class _Decoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(c.vocab_size, c.hidden_size).normal_(0.0, c.initializer_range))


class _Predictions(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.transform = _Transform(c)
        self.decoder = _Decoder(c)
        self.bias = nn.Parameter(torch.zeros(c.vocab_size))
        self.decoder.bias = self.bias          # same Parameter under both names (bert.py:604-607)


class _MLMHead(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.predictions = _Predictions(c)


class _LMHeadFn(torch.autograd.Function):
    """logits = decoder(LN(gelu(dense(h)))) + bias   (bert.py:575-609): two tcgen05 GEMMs and one LayerNorm."""

    @staticmethod
    def forward(ctx, h, wd, bd, lnw, lnb, wdec, bdec, eps, keep):
        shp = h.shape
        h2 = h.reshape(-1, shp[-1]).contiguous().float()
        hb = ops.scale_cast_bf16(h2)
        wdb = ops.cast_bf16(wd.detach().contiguous())
        pre = torch.empty((hb.shape[0], wd.shape[0]), device=h.device, dtype=BF16)
        t = ops.gemm(hb, wdb, bias=bd.detach(), act=ACT_GELU, aux_out=pre, out_dtype=F32)
        tb, _, mean, rstd = ops.layernorm_fwd(t, lnw.detach(), lnb.detach(), eps, save_stats=True)
        wdecb = ops.cast_bf16(wdec.detach().contiguous())
        V = wdec.shape[0]
        ldo = (V + 3) // 4 * 4
        buf = torch.empty((hb.shape[0], ldo), device=h.device, dtype=F32)
        logits = buf[:, :V]
        ops.gemm(tb, wdecb, bias=bdec.detach(), out=logits)
        if keep:
            ctx.save_for_backward(hb, wdb, pre, t, mean, rstd, tb, wdecb, lnw)
        ctx.shp = shp
        return logits.view(*shp[:-1], V)      # row pitch padded to 16 bytes: a strided view, never a copy

    @staticmethod
    def backward(ctx, dlogits):
        hb, wdb, pre, t, mean, rstd, tb, wdecb, lnw = ctx.saved_tensors
        V = wdecb.shape[0]
        dl = dlogits.reshape(-1, V)
        if dl.dtype != BF16 or dl.stride(0) % 8 or dl.stride(1) != 1:
            ldd = (V + 7) // 8 * 8          # generic consumer of .logits: repack the gradient as a 16-byte-pitch bf16 operand
            buf = torch.zeros((dl.shape[0], ldd), device=dl.device, dtype=BF16)
            buf[:, :V].copy_(dl)
            dlb = buf[:, :V]
        else:
            dlb = dl
        gdec = ops.gemm(dlb, tb, a_mn=True, b_mn=True, out_dtype=F32)
        gbdec = ops.colsum(dlb)
        dtb = ops.gemm(dlb, wdecb, b_mn=True)
        glnw, glnb = torch.empty_like(lnw), torch.empty_like(lnw)
        _, dtb2 = ops.layernorm_bwd(dtb, t, mean, rstd, lnw.detach(), glnw, glnb, want_f32=False, want_bf16=True)
        # dpre = dt * gelu'(pre): identity-free -- fold it into the dense dgrad/wgrad operand with the fp32 helper
        dpre = ops.gelu_f32(pre.float(), dtb2.float())
        dpreb = ops.scale_cast_bf16(dpre)
        gwd = ops.gemm(dpreb, hb, a_mn=True, b_mn=True, out_dtype=F32)
        gbd = ops.colsum(dpreb)
        dh = ops.gemm(dpreb, wdb, b_mn=True, out_dtype=F32).view(ctx.shp)
        return dh, gwd, gbd, glnw, glnb, gdec, gbdec, None, None


class _LMHeadLossFn(torch.autograd.Function):
    """K7: mean cross-entropy of decoder(LN(gelu(dense(h)))) + bias against `labels` (bert.py:575-609 + 1084-1090) WITHOUT the
    [rows, vocab] logits: the vocabulary is walked in chunks of `chunk` columns -- decoder GEMM into one reusable fp32 buffer,
    online log-sum-exp per row (mico_ce_chunk_update) -- and the backward pass recomputes each chunk with the same GEMM, turns
    it into bf16 dlogits (mico_ce_chunk_grad) and feeds the decoder's weight / bias / input gradients chunk by chunk.  Same
    arithmetic as _LMHeadFn + cross_entropy (bf16 dlogits there too); peak memory rows x chunk x 6 B instead of rows x vocab x 6 B."""

    @staticmethod
    def forward(ctx, h, wd, bd, lnw, lnb, wdec, bdec, labels, eps, ignore_index, chunk):
        shp = h.shape
        h2 = h.reshape(-1, shp[-1]).contiguous().float()
        hb = ops.scale_cast_bf16(h2)
        wdb = ops.cast_bf16(wd.detach().contiguous())
        pre = torch.empty((hb.shape[0], wd.shape[0]), device=h.device, dtype=BF16)
        t = ops.gemm(hb, wdb, bias=bd.detach(), act=ACT_GELU, aux_out=pre, out_dtype=F32)
        tb, _, mean, rstd = ops.layernorm_fwd(t, lnw.detach(), lnb.detach(), eps, save_stats=True)
        wdecb = ops.cast_bf16(wdec.detach().contiguous())
        V, M = wdec.shape[0], hb.shape[0]
        lab = labels.reshape(-1).contiguous()
        buf = torch.empty((M, chunk), device=h.device, dtype=F32)
        run_max, run_sum = torch.empty(M, device=h.device, dtype=F32), torch.empty(M, device=h.device, dtype=F32)
        lab_logit = torch.zeros(M, device=h.device, dtype=F32)
        bias = bdec.detach()
        for c0 in range(0, V, chunk):
            c1 = min(V, c0 + chunk)
            lg = ops.gemm(tb, wdecb[c0:c1], bias=bias[c0:c1], out=buf[:, :c1 - c0])
            ops.ce_chunk_update(lg, c0, lab, run_max, run_sum, lab_logit, first=(c0 == 0))
        stats, lse = ops.ce_chunk_finalize(run_max, run_sum, lab_logit, lab, V, ignore_index)
        ctx.save_for_backward(hb, wdb, pre, t, mean, rstd, tb, wdecb, lnw, lab, lse, stats, bias)
        ctx.meta = (shp, ignore_index, chunk)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, g):
        hb, wdb, pre, t, mean, rstd, tb, wdecb, lnw, lab, lse, stats, bias = ctx.saved_tensors
        shp, ignore_index, chunk = ctx.meta
        V, M, H = wdecb.shape[0], hb.shape[0], tb.shape[1]
        dev = hb.device
        g = g.detach().reshape(1).float().contiguous()
        buf = torch.empty((M, chunk), device=dev, dtype=F32)
        dl = torch.empty((M, chunk), device=dev, dtype=BF16)
        gdec = torch.empty((V, H), device=dev, dtype=F32)
        gbdec = torch.empty(V, device=dev, dtype=F32)
        dt = torch.zeros((M, H), device=dev, dtype=F32)
        for c0 in range(0, V, chunk):
            c1 = min(V, c0 + chunk)
            n = c1 - c0
            lg = ops.gemm(tb, wdecb[c0:c1], bias=bias[c0:c1], out=buf[:, :n])          # the forward pass's own launch
            ops.ce_chunk_grad(lg, c0, lab, lse, g, stats, V, dl, ignore_index)
            dlc = dl[:, :n]
            ops.gemm(dlc, tb, a_mn=True, b_mn=True, out=gdec[c0:c1])                     # d decoder.weight rows
            ops.colsum(dlc, out=gbdec[c0:c1])
            ops.gemm(dlc, wdecb[c0:c1], b_mn=True, out=dt, accumulate=(c0 > 0))         # dt += dlogits_c W_c
        dtb = ops.scale_cast_bf16(dt)
        glnw, glnb = torch.empty_like(lnw), torch.empty_like(lnw)
        _, dtb2 = ops.layernorm_bwd(dtb, t, mean, rstd, lnw.detach(), glnw, glnb, want_f32=False, want_bf16=True)
        dpre = ops.gelu_f32(pre.float(), dtb2.float())
        dpreb = ops.scale_cast_bf16(dpre)
        gwd = ops.gemm(dpreb, hb, a_mn=True, b_mn=True, out_dtype=F32)
        gbd = ops.colsum(dpreb)
        dh = ops.gemm(dpreb, wdb, b_mn=True, out_dtype=F32).view(shp)
        return dh, gwd, gbd, glnw, glnb, gdec, gbdec, None, None, None, None


class BertForMaskedLM(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = _MLMHead(config)
        if config.tie_word_embeddings:     # transformers 4.31 post_init() ties decoder and word embeddings
            self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight
        # lm_loss(): LM head + cross-entropy chunk by chunk over the vocabulary (K7); forward(labels=...) still returns .logits
        self.fused_lm_loss = True
        self.lm_loss_chunk = 4096

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def _set_gradient_checkpointing(self, module, value=False):
        if isinstance(module, BertEncoder):
            module.gradient_checkpointing = value

    def mask_position_logits(self, input_ids, attention_mask, encoder_hidden_states=None):
        """fp32 logits (rows, vocab) at the LAST position only -- one decode step of generation reads the prediction at the
        appended [MASK] token (bert.py:1126-1143), so the 30522-wide LM-head GEMM runs on one row per sequence."""
        seq = self.bert(input_ids, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states).last_hidden_state
        pr = self.cls.predictions
        logits = _LMHeadFn.apply(seq[:, -1:, :].contiguous(), pr.transform.dense.weight, pr.transform.dense.bias,
                                 pr.transform.LayerNorm.weight, pr.transform.LayerNorm.bias, pr.decoder.weight, pr.bias,
                                 self.config.layer_norm_eps, False)
        return logits[:, 0, :]

    def generate(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, max_new_tokens=20, num_beams=1,
                 eos_token_id=None, pad_token_id=None, length_penalty=1.0, do_sample=False, top_k=None, **kwargs):
        """The reference's `multimodal_encoder.generate(...)` call (inference_demo.py:164-171, vast.py:527-545): beam search /
        greedy / top-k sampling over the [MASK]-append decode step; see mico_b200/generation.py."""
        from .generation import generate as _generate
        return _generate(self, input_ids, attention_mask, encoder_hidden_states=encoder_hidden_states,
                         max_new_tokens=max_new_tokens, num_beams=num_beams, eos_token_id=eos_token_id,
                         pad_token_id=pad_token_id, length_penalty=length_penalty, do_sample=do_sample, top_k=top_k,
                         mask_token_id=kwargs.get("mask_token_id"), generator=kwargs.get("generator"))

    def lm_loss(self, seq, labels):
        """LM head + cross-entropy (bert.py:1084-1090) on an encoder output computed elsewhere (a slice of a larger call)."""
        pr = self.cls.predictions
        if self.fused_lm_loss:     # K7: the [rows, vocab] logits are never materialised
            return _LMHeadLossFn.apply(seq, pr.transform.dense.weight, pr.transform.dense.bias, pr.transform.LayerNorm.weight,
                                       pr.transform.LayerNorm.bias, pr.decoder.weight, pr.bias, labels,
                                       self.config.layer_norm_eps, -100, self.lm_loss_chunk)
        logits = _LMHeadFn.apply(seq, pr.transform.dense.weight, pr.transform.dense.bias, pr.transform.LayerNorm.weight,
                                 pr.transform.LayerNorm.bias, pr.decoder.weight, pr.bias, self.config.layer_norm_eps,
                                 torch.is_grad_enabled())
        return cross_entropy(logits.view(-1, self.config.vocab_size), labels.reshape(-1), ignore_index=-100, grad_dtype=BF16)

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, encoder_hidden_states=None, encoder_attention_mask=None, labels=None,
                output_attentions=None, output_hidden_states=None, return_dict=None):
        seq = self.bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids, position_ids=position_ids,
                        head_mask=head_mask, inputs_embeds=inputs_embeds, encoder_hidden_states=encoder_hidden_states,
                        encoder_attention_mask=encoder_attention_mask).last_hidden_state
        pr = self.cls.predictions
        keep = torch.is_grad_enabled()

        def make_logits():
            return _LMHeadFn.apply(seq, pr.transform.dense.weight, pr.transform.dense.bias, pr.transform.LayerNorm.weight,
                                   pr.transform.LayerNorm.bias, pr.decoder.weight, pr.bias, self.config.layer_norm_eps, keep)
        if labels is not None and self.fused_lm_loss:
            return _LazyLogitsOut(make_logits, loss=self.lm_loss(seq, labels), sequence_output=seq)
        logits = make_logits()
        loss = None
        if labels is not None:
            loss = cross_entropy(logits.view(-1, self.config.vocab_size), labels.view(-1), ignore_index=-100,
                                 grad_dtype=BF16)
        return _Out(loss=loss, logits=logits, sequence_output=seq)
