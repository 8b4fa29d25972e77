"""MiCo omni-modal module on the sm_100a kernels: the drop-in for the reference's ``model/mico.py`` ``MiCo`` class.

Same constructor (``MiCo(config)``, ``MiCo.from_pretrained(opts, state_dict)``), same ``state_dict`` schema
(SURVEY.md 8b), same encoder / pooling / fusion-input / head methods that ``inference_demo.py:132-171`` calls
(reference model/mico.py:115-248, 374-423).  The released MiCo class has no ``forward``: the training step the loop
calls (``model(batch, task, compute_loss=True)``, data/utils/pipeline.py:44) is stated only in the sibling VAST tree,
so ``forward`` / ``forward_ret`` / ``forward_cap`` here follow data/model/vast.py:317-348, 383-464, 485-512 with the
collectives of data/utils/distributed.py:12-66.

Compute placement: vision / audio / depth towers = ``eva_vit.EVAVisionTransformer`` (one autograd node, tcgen05
GEMMs + fused attention); text / fusion = ``bert.BertForMaskedLM`` (one autograd node + LM head + CE kernel);
heads, normalisation, contrastive logits and cross-entropies = fp32 kernels via ``functional``.  torch ops that remain
are glue on tiny tensors (cls select / frame mean, broadcast adds of frame / type embeddings, torch.cat of pooled
features, multinomial sampling of hard negatives) and the NCCL collectives.
"""
import random

import torch
import torch.distributed as dist
import torch.nn as nn

from . import functional as MF
from .bert import BertConfig, BertForMaskedLM, _Out
from .eva_vit import EVAVisionTransformer
from .ops import F32, MicoError

# EVA-CLIP vision towers (model/evaclip/model_configs/*.json): name -> constructor arguments
_EVA_CFG = {
    "evaclip01_giant": dict(patch_size=14, embed_dim=1408, depth=40, num_heads=16, mlp_ratio=4.3637, drop_path_rate=0.4,
                            num_classes=1024),
}


# model/evaclip/model_configs/EVA02-CLIP-B-16.json / EVA02-CLIP-L-14.json (mico.py:326-339); drop_path_rate 0
_EVA02_CFG = {
    "evaclip02_base": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=2.6667, num_classes=512),
    "evaclip02_base_self": dict(patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=2.6667, num_classes=512),
    "evaclip02_large": dict(patch_size=14, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=2.6667, num_classes=768),
}


_CLIP_CFG = {
    "clip_vit_base_16": dict(patch_size=16, width=768, layers=12, heads=12, output_dim=512),
    "clip_vit_large_14_336px": dict(patch_size=14, width=1024, layers=24, heads=16, output_dim=768),
}


class _AttrDict(dict):
    """attribute access over a dict (stands in for easydict.EasyDict, which the reference wraps batches/configs in)"""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = dict.__setitem__


class _Holder(nn.Module):
    """Parameter holder for nn.Linear-shaped heads (weight [out,in], optional bias)."""

    def __init__(self, i, o, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(o, i).normal_(0.0, 0.02))
        self.bias = nn.Parameter(torch.zeros(o)) if bias else None

    def forward(self, x):
        return MF.linear(x, self.weight, self.bias)


class _LayerNorm(nn.Module):
    def __init__(self, d, eps=1e-12):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))

    def forward(self, x):
        return MF.layer_norm(x, self.weight, self.bias, self.eps)


class Contra_head(nn.Module):
    """mico.py:36-41: bias-free Linear(input_dim -> contra_dim)."""

    def __init__(self, input_dim, contra_dim):
        super().__init__()
        self.linear = _Holder(input_dim, contra_dim, bias=False)

    def forward(self, cls_token):
        return MF.linear_f32(cls_token, self.linear.weight, None)


class Match_head(nn.Module):
    """mico.py:44-52: Linear -> GELU(erf) -> LayerNorm(1e-12) -> Linear(hidden -> 2)."""

    def __init__(self, hidden_size):
        super().__init__()
        self.linear1 = _Holder(hidden_size, hidden_size)
        self.layernorm = _LayerNorm(hidden_size, eps=1e-12)
        self.linear2 = _Holder(hidden_size, 2)

    def forward(self, cls_token):
        x = MF.gelu(MF.linear_f32(cls_token.float(), self.linear1.weight, self.linear1.bias))
        return MF.linear_f32(self.layernorm(x), self.linear2.weight, self.linear2.bias)


class _VisionEncoder(nn.Module):
    """Stands in for the reference's CustomCLIP (evaclip/model.py:272-314): ``.visual`` tower + ``logit_scale``;
    ``.text`` exists only so that ``del model.vision_encoder.text`` (mico.py:419) works."""

    def __init__(self, visual):
        super().__init__()
        self.visual = visual
        self.text = nn.Identity()
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592600369)    # log(1 / 0.07)

    def set_grad_checkpointing(self, enable=True):
        self.visual.set_grad_checkpointing(enable)


class TokenMasker:
    """general_module.py:52-97: host-side MLM masking (80 % [MASK], 10 % random token, 10 % kept), python RNG."""

    def __init__(self, mask_token=-1, range_start=-1, range_end=-1):
        self.mask_token, self.range = mask_token, [range_start, range_end]

    def __call__(self, tokens, mask_prob):
        dev = tokens.device
        tok = tokens.clone().cpu().tolist()
        labels = [[-100] * len(r) for r in tok]
        for i, row in enumerate(tok):
            ind = [0] * len(row)
            if not any(t != 0 for t in row[1:]):
                continue
            while not any(ind):
                for j in range(1, len(row)):
                    if row[j] != 0 and random.random() < mask_prob:
                        ind[j] = 1
            for j in range(len(row)):
                if ind[j]:
                    src, prob = row[j], random.random()
                    if prob < 0.8:
                        row[j] = self.mask_token
                    elif prob < 0.9:
                        row[j] = random.choice(range(*self.range))
                    labels[i][j] = src
        return torch.tensor(tok, dtype=torch.long, device=dev), torch.tensor(labels, dtype=torch.long, device=dev)


# ---------------------------------------------------------------------------------------------- collectives
_GROUP = None      # process group of the data-parallel world (None = the default group)


def set_process_group(group):
    """Run the loss collectives over `group` instead of the default process group (a sub-world, e.g. a 2-rank replay of a
    reference fixture inside a larger job).  Returns the previous setting."""
    global _GROUP
    prev, _GROUP = _GROUP, group
    return prev


def _world():
    return dist.get_world_size(_GROUP) if dist.is_available() and dist.is_initialized() else 1


def _rank():
    return dist.get_rank(_GROUP) if dist.is_available() and dist.is_initialized() else 0


@torch.no_grad()
def concat_all_gather(t):
    """distributed.py:53-66 without the list-of-tensors copy: one all_gather_into_tensor into a contiguous buffer."""
    w = _world()
    if w == 1:
        return t
    t = t.contiguous()
    out = torch.empty((w * t.shape[0],) + tuple(t.shape[1:]), device=t.device, dtype=t.dtype)
    dist.all_gather_into_tensor(out, t, group=_GROUP)
    return out


class _GatherWithGrad(torch.autograd.Function):
    """distributed.py:12-30 GatherLayer: forward all-gather; backward all-reduce(SUM) of the stacked gradients and
    take this rank's slice == reduce-scatter(SUM)."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        out = torch.empty((_world() * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
        dist.all_gather_into_tensor(out, x, group=_GROUP)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        n = g.shape[0] // _world()
        if dist.get_backend(_GROUP) == "gloo":      # CPU test path: gloo has no reduce_scatter; the reference's own formulation
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=_GROUP)
            return g[_rank() * n:(_rank() + 1) * n].clone()
        out = torch.empty((n,) + tuple(g.shape[1:]), device=g.device, dtype=g.dtype)
        dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=_GROUP)
        return out


def all_gather_with_grad(t):
    return t if _world() == 1 else _GatherWithGrad.apply(t)


class MiCo(nn.Module):
    """VLP pretraining module (reference model/mico.py:374-423)."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.construct_vision_encoder()
        self.audio_dim = self.vision_dim      # mico.py:92-96: audio and depth reuse the vision tower
        self.depth_dim = self.vision_dim
        self.construct_multimodal_encoder()
        cd, md, vd = self.config.contra_dim, self.multimodal_dim, self.vision_dim
        self.contra_head_t = Contra_head(md, cd)
        self.contra_head_s = Contra_head(md, cd)
        self.contra_head_v = Contra_head(vd, cd)
        self.contra_head_a = Contra_head(self.audio_dim, cd)
        self.contra_head_d = Contra_head(self.depth_dim, cd)
        self.contra_head_va = _Holder(vd + self.audio_dim, cd)
        self.contra_head_id = _Holder(vd + self.depth_dim, cd)
        self.contra_head_vs = _Holder(vd + md, cd)
        self.contra_head_vas = _Holder(vd + self.audio_dim + md, cd)
        self.contra_temp = nn.Parameter(torch.tensor(0.07))
        self.itm_head = Match_head(md)
        c = self.config
        self.vision_frame_embedding = nn.Parameter(0.02 * torch.randn(1, c.max_vision_sample_num, md))
        self.audio_frame_embedding = nn.Parameter(0.02 * torch.randn(1, c.max_audio_sample_num, md))
        self.depth_frame_embedding = nn.Parameter(0.02 * torch.randn(1, c.max_depth_sample_num, md))
        self.hidden_trans_vision_multimodal = nn.Sequential(_Holder(vd, md), _LayerNorm(md, 1e-12))
        self.hidden_trans_audio_multimodal = nn.Sequential(_Holder(self.audio_dim, md), _LayerNorm(md, 1e-12))
        self.hidden_trans_depth_multimodal = nn.Sequential(_Holder(self.depth_dim, md), _LayerNorm(md, 1e-12))
        self.hidden_trans_subtitle_multimodal = nn.Sequential(_Holder(md, md), _LayerNorm(md, 1e-12))
        self.vision_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, md))
        self.audio_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, md))
        self.depth_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, md))
        self.subtitle_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, md))
        self.beam_size = c.beam_size
        self.itm_ratio = c.itm_ratio
        self.max_omni_caption_len = c.max_omni_caption_len
        self.max_caption_len = c.max_caption_len
        self.max_subtitle_len = c.max_subtitle_len
        self.text_masker = TokenMasker(mask_token=103, range_start=106, range_end=30522)   # vast.py:79 ([MASK] = 103)

    # ------------------------------------------------------------------ construction (mico.py:74-113, 329-352)
    def construct_vision_encoder(self):
        t = self.config.vision_encoder_type
        if t in _EVA_CFG:
            kw = dict(_EVA_CFG[t])
            kw.update(getattr(self.config, "vision_tower_kwargs", None) or {})   # test hook: smaller towers
            self.vision_dim = kw["embed_dim"]
            tower = EVAVisionTransformer(img_size=self.config.vision_resolution, qkv_bias=True, use_mean_pooling=False,
                                         grad_checkpointing=bool(self.config.checkpointing), **kw)
            self.vision_encoder = _VisionEncoder(tower)
        elif t in _EVA02_CFG:     # RoPE / SwiGLU / sub-LN towers (eva_vit_model.py with rope, naiveswiglu, subln)
            from .eva02_vit import EVA02VisionTransformer
            kw = dict(_EVA02_CFG[t])
            kw.update(getattr(self.config, "vision_tower_kwargs", None) or {})
            self.vision_dim = kw["embed_dim"]
            tower = EVA02VisionTransformer(img_size=self.config.vision_resolution, qkv_bias=True, use_mean_pooling=False,
                                           grad_checkpointing=bool(self.config.checkpointing), **kw)
            self.vision_encoder = _VisionEncoder(tower)
        elif t in _CLIP_CFG:      # OpenAI CLIP towers (mico.py:354-371; reference needs JIT weights at a hard-coded path)
            from .clip_vit import VisionTransformer
            kw = dict(_CLIP_CFG[t])
            kw.update(getattr(self.config, "vision_tower_kwargs", None) or {})
            self.vision_dim = kw["width"]
            kw["input_resolution"] = self.config.vision_resolution
            self.vision_encoder = _VisionEncoder(VisionTransformer(checkpointing=bool(self.config.checkpointing), **kw))
        elif t.startswith("swin"):     # mico.py:86 calls an undefined load_swin_model(); general_module.py:230-241 builds Swin-B 22k
            from .swin import SwinTransformer
            kw = dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=7, drop_path_rate=0.2,
                      num_classes=0)
            kw.update(getattr(self.config, "vision_tower_kwargs", None) or {})
            self.vision_encoder = SwinTransformer(img_size=self.config.vision_resolution, **kw)
            self.vision_dim = self.vision_encoder.num_features
        else:
            raise NotImplementedError(f"vision_encoder_type {t!r}: EVA01-g, EVA02-B / -L, the OpenAI CLIP ViTs and Swin are built; "
                                      "EVA02-bigE (head_dim 112, post-norm) and VideoSwin towers are not (DESIGN.md)")

    def construct_multimodal_encoder(self):
        bert_kw = dict(getattr(self.config, "bert_config", None) or {})
        self.multimodal_encoder = BertForMaskedLM(BertConfig(**bert_kw))
        self.multimodal_dim = self.multimodal_encoder.config.hidden_size
        self.multimodal_encoder.tokenizer = None     # the caller attaches a HF BertTokenizer (mico.py:109-113)

    @classmethod
    def from_pretrained(cls, opts, state_dict, *inputs, **kwargs):
        model = cls(opts, *inputs, **kwargs)
        missing_keys, unexpected_keys = model.load_state_dict(state_dict, strict=False)
        if hasattr(model.vision_encoder, "text"):
            del model.vision_encoder.text
        if state_dict != {}:
            print(f"Unexpected keys {unexpected_keys}")
            print(f"missing_keys  {missing_keys}")
        return model

    # ------------------------------------------------------------------ encoders (mico.py:115-155)
    def forward_vision_encoder(self, vision_pixels):
        b, n, _, h, w = vision_pixels.shape
        if self.config.vision_encoder_type.startswith("swin"):          # mico.py:125-127
            out = self.vision_encoder(vision_pixels.reshape(b * n, 3, h, w))
        else:
            out = self.vision_encoder.visual(vision_pixels.reshape(b * n, 3, h, w), return_all_features=True)
        return out.reshape(b, -1, *out.shape[-2:])

    def forward_audio_encoder(self, audio_spectrograms):
        # reference: unsqueeze(2).repeat(1,1,3,1,1) then the vision tower (mico.py:139-143)
        if self.config.vision_encoder_type.startswith("swin"):
            return self.forward_vision_encoder(audio_spectrograms.unsqueeze(2).repeat(1, 1, 3, 1, 1))
        b, n, h, w = audio_spectrograms.shape
        out = self.vision_encoder.visual(audio_spectrograms.reshape(b * n, h, w), return_all_features=True)
        return out.reshape(b, -1, *out.shape[-2:])

    def forward_depth_encoder(self, depth_pixels):
        return self.forward_vision_encoder(depth_pixels)

    def forward_multimodal_encoder(self, input_ids, attention_mask, condition_feat=None, labels=None, position_ids=None,
                                   preprocess=True):
        return self.multimodal_encoder(input_ids=input_ids, attention_mask=attention_mask,
                                       encoder_hidden_states=condition_feat, labels=labels)

    # ------------------------------------------------------------------ pooling (mico.py:157-185)
    def pool_vision_for_contra(self, feature):
        if self.config.vision_encoder_type.startswith("swin"):          # mico.py:161-162: token mean, then frame mean
            return feature.mean(dim=2).mean(dim=1)
        return feature[:, :, 0].mean(dim=1)

    pool_audio_for_contra = pool_vision_for_contra
    pool_depth_for_contra = pool_vision_for_contra

    def pool_text_for_contra(self, feature):
        return feature[:, 0]

    # ------------------------------------------------------------------ fusion inputs (mico.py:187-248)
    def _fusion_input(self, out, trans, frame_emb, type_emb, adaptive=True):
        b, n, x, c = out.shape
        if self.config.pool_video:
            out = torch.cat([out[:, :, 0:1], out[:, :, 1:].mean(2, keepdim=True)], dim=2)
        out = trans(out)
        if adaptive and frame_emb is not None:
            fe = frame_emb
            if n != fe.shape[1]:      # nearest-neighbour resize of the frame table (mico.py:196-199)
                fe = torch.nn.functional.interpolate(fe.float().permute(0, 2, 1), n, mode="nearest").permute(0, 2, 1)
            out = out + fe.unsqueeze(-2)
        out = out.reshape(b, -1, self.multimodal_dim)
        return out + type_emb

    def get_multimodal_forward_input_vision(self, vision_output):
        return self._fusion_input(vision_output, self.hidden_trans_vision_multimodal, self.vision_frame_embedding,
                                  self.vision_type_embeddings, self.config.frame_embedding_type == "adaptive")

    def get_multimodal_forward_input_audio(self, audio_output):
        return self._fusion_input(audio_output, self.hidden_trans_audio_multimodal, self.audio_frame_embedding,
                                  self.audio_type_embeddings)

    def get_multimodal_forward_input_depth(self, depth_output):
        return self._fusion_input(depth_output, self.hidden_trans_depth_multimodal, self.depth_frame_embedding,
                                  self.depth_type_embeddings)

    def get_multimodal_forward_input_subtitle(self, subtitle_output):
        return self.hidden_trans_subtitle_multimodal(subtitle_output) + self.subtitle_type_embeddings

    # ------------------------------------------------------------------ lazy feature cache (vast.py:81-314)
    def _tokens(self, batch, key, texts_key, max_len, on_host=False):
        """on_host: leave the tokenizer's CPU tensors where they are (the fused step masks them on the host before any
        kernel is queued, mico_b200/train_step.py)."""
        if key in batch:
            return batch[key]
        tok = self.multimodal_encoder.tokenizer
        if tok is None:
            raise MicoError(f"batch has no {key!r} and no tokenizer is attached to model.multimodal_encoder.tokenizer")
        t = tok(batch[texts_key], padding="max_length", truncation=True, max_length=max_len, return_tensors="pt")
        dev = torch.device("cpu") if on_host else self.contra_temp.device
        batch[key] = _AttrDict(input_ids=t["input_ids"].to(dev), attention_mask=t["attention_mask"].to(dev))
        return batch[key]

    def batch_get(self, batch, key):
        if key in batch:
            return batch[key]
        g = lambda k: self.batch_get(batch, k)
        if key == "caption_tokens":
            return self._tokens(batch, key, "raw_captions", self.max_caption_len)
        if key == "subtitle_tokens":
            return self._tokens(batch, key, "raw_subtitles", self.max_subtitle_len)
        if key in ("caption_output", "subtitle_output"):
            t = g(key.replace("output", "tokens"))
            v = self.multimodal_encoder.bert(input_ids=t.input_ids, attention_mask=t.attention_mask).last_hidden_state
        elif key == "vision_output":
            v = self.forward_vision_encoder(batch["vision_pixels"])
        elif key == "audio_output":
            v = self.forward_audio_encoder(batch["audio_spectrograms"])
        elif key == "depth_output":
            v = self.forward_depth_encoder(batch["depth_pixels"])
        elif key == "condition_feats_v":
            v = self.get_multimodal_forward_input_vision(g("vision_output"))
        elif key == "condition_feats_a":
            v = self.get_multimodal_forward_input_audio(g("audio_output"))
        elif key == "condition_feats_d":
            v = self.get_multimodal_forward_input_depth(g("depth_output"))
        elif key == "condition_feats_s":
            v = self.get_multimodal_forward_input_subtitle(g("subtitle_output"))
        elif key.startswith("condition_feats_"):        # va, vs, vas, id: concatenation along the token axis
            parts = {"va": "va", "vs": "vs", "vas": "vas", "id": "vd"}.get(key[len("condition_feats_"):])
            if parts is None:
                raise NotImplementedError(key)
            v = torch.cat([g(f"condition_feats_{m}") for m in parts], dim=1)
        elif key == "feat_t":
            v = MF.normalize(self.contra_head_t(self.pool_text_for_contra(g("caption_output"))))
        elif key == "feat_s":
            v = MF.normalize(self.contra_head_s(self.pool_text_for_contra(g("subtitle_output"))))
        elif key == "feat_v":
            v = MF.normalize(self.contra_head_v(self.pool_vision_for_contra(g("vision_output"))))
        elif key == "feat_a":
            v = MF.normalize(self.contra_head_a(self.pool_audio_for_contra(g("audio_output"))))
        elif key == "feat_d":
            v = MF.normalize(self.contra_head_d(self.pool_depth_for_contra(g("depth_output"))))
        elif key in ("feat_va", "feat_vs", "feat_vas", "feat_id"):
            pools = {"v": lambda: self.pool_vision_for_contra(g("vision_output")),
                     "a": lambda: self.pool_audio_for_contra(g("audio_output")),
                     "d": lambda: self.pool_depth_for_contra(g("depth_output")),
                     "s": lambda: self.pool_text_for_contra(g("subtitle_output"))}
            combo = {"feat_va": "va", "feat_vs": "vs", "feat_vas": "vas", "feat_id": "vd"}[key]
            head = getattr(self, "contra_head_" + key[5:])
            x = torch.cat([pools[m]() for m in combo], dim=1)
            v = MF.normalize(MF.linear_f32(x, head.weight, head.bias))
        else:
            raise NotImplementedError(key)
        batch[key] = v
        return v

    # ------------------------------------------------------------------ training step (vast.py:317-348)
    _SUBTASKS = ("tv", "ta", "tva", "tvs", "tvas", "td", "tid")   # td / tid: depth combos (unpinned upstream)

    def forward(self, batch, task, compute_loss=True):
        batch = _AttrDict(batch)
        # Training (gradients on): the same losses on the schedule of mico_b200/train_step.py -- one tower pass for all
        # modalities, ITM / caption sub-tasks differentiated group by group so that their activations never coexist.
        # config.step_schedule = "reference" keeps the reference's lazy one-sub-task-at-a-time order below.
        if compute_loss and torch.is_grad_enabled() and getattr(self.config, "step_schedule", "fused") == "fused" \
                and hasattr(getattr(self.vision_encoder, "visual", None), "forward_multi"):
            from .train_step import fused_train_forward
            if "caption_tokens" not in batch and "raw_captions" in batch:
                self._tokens(batch, "caption_tokens", "raw_captions", self.max_caption_len, on_host=True)
            return fused_train_forward(self, batch, task)
        tok = batch.get("caption_tokens")
        if tok is not None and not tok.input_ids.is_cuda:        # a tokenizer's host tensors
            dev = self.contra_temp.device
            batch["caption_tokens"] = _AttrDict(input_ids=tok.input_ids.to(dev), attention_mask=tok.attention_mask.to(dev))
        out = {}
        for t in task.split("_"):
            if t.startswith("ret"):
                out.update(self.forward_ret(batch, t, compute_loss=compute_loss))
            elif t.startswith("cap"):
                out.update(self.forward_cap(batch, t, compute_loss=compute_loss))
            else:
                raise NotImplementedError(t)
        return out

    def forward_ret(self, batch, task, compute_loss=True):
        subtasks = task.split("%")[1:]
        feat_t = self.batch_get(batch, "feat_t")
        tokens = self.batch_get(batch, "caption_tokens")
        input_ids, attention_mask = tokens.input_ids, tokens.attention_mask
        if not compute_loss:
            ev = dict(feat_t=feat_t, input_ids=input_ids, attention_mask=attention_mask)
            for st in subtasks:
                assert st in self._SUBTASKS
                ev[f"feat_cond_{st}"] = self.batch_get(batch, f"feat_{st[1:]}")
                ev[f"condition_feats_{st}"] = self.batch_get(batch, f"condition_feats_{st[1:]}")
            return ev
        loss_itc, loss_itm = [], []
        feat_t_all = concat_all_gather(feat_t)
        ids_all = concat_all_gather(input_ids)
        att_all = concat_all_gather(attention_mask)
        rank, bs = _rank(), feat_t.shape[0]
        targets = torch.arange(rank * bs, rank * bs + bs, device=feat_t.device)
        for st in subtasks:
            assert st in self._SUBTASKS
            # ---- ITC (vast.py:402-417): both directions against the gathered negatives, label smoothing 0.1
            feat_c = self.batch_get(batch, f"feat_{st[1:]}")
            feat_c_all = concat_all_gather(feat_c)
            sim_c2t = MF.contrastive_logits(feat_c, feat_t_all, self.contra_temp)
            sim_t2c = MF.contrastive_logits(feat_t, feat_c_all, self.contra_temp)
            loss_itc.append((MF.cross_entropy(sim_c2t, targets, label_smoothing=0.1)
                             + MF.cross_entropy(sim_t2c, targets, label_smoothing=0.1)) / 2)
            # ---- ITM (vast.py:419-457): one hard negative per sample and direction, sampled from the ITC similarities
            cond = self.batch_get(batch, f"condition_feats_{st[1:]}")
            cond_all = all_gather_with_grad(cond)
            with torch.no_grad():
                w_t2c = torch.softmax(sim_t2c.detach(), dim=1) + 1e-4
                w_t2c[:, rank * bs:rank * bs + bs].fill_diagonal_(0)
                w_c2t = torch.softmax(sim_c2t.detach(), dim=1) + 1e-4
                w_c2t[:, rank * bs:rank * bs + bs].fill_diagonal_(0)
                # one batched draw per direction instead of the reference's 2*bs multinomial(...).item() host syncs
                neg_c = batch.get(f"itm_neg_cond_{st}")
                neg_t = batch.get(f"itm_neg_text_{st}")
                neg_c = torch.multinomial(w_t2c, 1).view(-1) if neg_c is None else neg_c.to(cond.device)
                neg_t = torch.multinomial(w_c2t, 1).view(-1) if neg_t is None else neg_t.to(cond.device)
            cond_neg = cond_all[neg_c]
            ids_1 = torch.cat((input_ids, input_ids, ids_all[neg_t]), dim=0)
            att_1 = torch.cat((attention_mask, attention_mask, att_all[neg_t]), dim=0)
            cond_3 = torch.cat((cond, cond_neg, cond), dim=0)
            output = self.multimodal_encoder.bert(input_ids=ids_1, attention_mask=att_1,
                                                  encoder_hidden_states=cond_3).last_hidden_state
            logits = self.itm_head(output[:, 0])
            truth = torch.zeros(bs * 3, dtype=torch.long, device=logits.device)
            truth[:bs] = 1
            loss_itm.append(self.itm_ratio * MF.cross_entropy(logits, truth))
        return dict(loss_itc=sum(loss_itc) / len(loss_itc), loss_itm=sum(loss_itm) / len(loss_itm))

    def _generate_captions(self, batch, subtasks):
        """Evaluation branch of forward_cap (data/model/vast.py:514-553): beam search (or top-k sampling in captioner_mode)
        from [CLS] over the fusion inputs of every modality combination; returns token ids, and decoded strings when a
        tokenizer is attached."""
        enc = self.multimodal_encoder
        tok = enc.tokenizer
        ids_of = lambda name, default: getattr(tok, name, None) if tok is not None and getattr(tok, name, None) is not None else default
        bos, sep, pad, msk = ids_of("bos_token_id", 101), ids_of("sep_token_id", 102), ids_of("pad_token_id", 0), ids_of("mask_token_id", 103)
        out = {}
        for st in subtasks:
            assert st in self._SUBTASKS
            cond = self.batch_get(batch, f"condition_feats_{st[1:]}")
            b = cond.shape[0]
            captioner = bool(getattr(self.config, "captioner_mode", False))
            if captioner:
                n = int(self.config.generate_nums)
                cond = cond.unsqueeze(1).expand(-1, n, -1, -1).reshape(-1, *cond.shape[1:])
                b *= n
            init_ids = torch.full((b, 1), bos, dtype=torch.long, device=cond.device)
            init_mask = init_ids.new_ones(b, 1, 1)
            if captioner:
                ids = enc.generate(input_ids=init_ids, attention_mask=init_mask, do_sample=True, top_k=10,
                                   encoder_hidden_states=cond, max_new_tokens=self.max_caption_len, eos_token_id=sep,
                                   pad_token_id=pad, mask_token_id=msk)
            else:
                ids = enc.generate(input_ids=init_ids, attention_mask=init_mask, encoder_hidden_states=cond,
                                   max_new_tokens=self.max_caption_len, num_beams=self.beam_size, eos_token_id=sep,
                                   pad_token_id=pad, length_penalty=0.6, mask_token_id=msk)
            new = ids[:, 1:]
            out[f"generated_captions_{st}"] = tok.batch_decode(new, skip_special_tokens=True) if tok is not None else new
        return out

    def forward_cap(self, batch, task, compute_loss=True):
        subtasks = task.split("%")[1:]
        if not compute_loss:
            return self._generate_captions(batch, subtasks)
        tokens = self.batch_get(batch, "caption_tokens")
        input_ids, attention_mask = tokens.input_ids, tokens.attention_mask
        if "cap_input_ids" in batch:       # parity hook: masked ids / labels supplied by the caller
            input_ids, labels = batch["cap_input_ids"], batch["cap_labels"]
        else:
            input_ids, labels = self.text_masker(input_ids, 0.6)
        S = attention_mask.shape[1]
        att3 = torch.tril(attention_mask.unsqueeze(1).expand(-1, S, -1).clone())     # vast.py:497-499
        losses = []
        for st in subtasks:
            assert st in self._SUBTASKS
            cond = self.batch_get(batch, f"condition_feats_{st[1:]}")
            losses.append(self.multimodal_encoder(input_ids=input_ids, attention_mask=att3, encoder_hidden_states=cond,
                                                  labels=labels).loss)
        return dict(loss_cap=sum(losses) / len(losses))
