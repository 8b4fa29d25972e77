"""Generate golden fixtures by running the UNMODIFIED reference on CPU (build container only).

    python oracle/make_golden.py [--only vit|bert|mico|loss|fbank|swin|clip|transformer]

Writes tests/golden/*.pt (+ MANIFEST.json with library versions).  Each fixture holds the
reference module's state_dict, seeded inputs, and the reference's outputs (and gradients),
so that tests can re-check the oracle (CPU) and the CUDA product (GPU) without
/root/reference being present.  Fixtures are small configurations of the same code path
(same head_dim 88 / patch 14 / token count 257 for the ViT) because full ViT-g weights are
4 GB; full-size parity is covered by oracle-vs-CUDA tests on seeded random weights.
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
GOLD = os.path.join(REPO, "tests", "golden")

import torch  # noqa: E402

from oracle import ref_shims  # noqa: E402


def _save(name, obj):
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, name)
    torch.save(obj, path)
    print(f"wrote {path}  ({os.path.getsize(path) / 1e6:.2f} MB)")


def _randomize(module, seed):
    """Perturb every parameter so that zero-initialised biases etc. carry signal."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for _, prm in sorted(module.named_parameters()):
            prm.add_(0.02 * torch.randn(prm.shape, generator=g))


def gen_vit():
    """Reference EVAVisionTransformer (model/evaclip/eva_vit_model.py:488) at width 176
    (2 heads x 88), depth 2, mlp 352, patch 14, 224x224 -> 257 tokens; eval fwd, and train
    fwd+bwd with DropPath rate 0.4 under a fixed torch seed."""
    from functools import partial
    from model.evaclip.eva_vit_model import EVAVisionTransformer
    from model.evaclip.transformer import LayerNorm
    cfg = dict(width=176, depth=2, heads=2, mlp=352, patch=14, image=224, eps=1e-6)
    torch.manual_seed(0)
    m = EVAVisionTransformer(img_size=224, patch_size=14, num_classes=8, use_mean_pooling=False,
                             embed_dim=176, depth=2, num_heads=2, mlp_ratio=2.0, qkv_bias=True,
                             drop_path_rate=0.4, norm_layer=partial(LayerNorm, eps=1e-6), xattn=True)
    _randomize(m, 1)
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(2, 3, 224, 224, generator=g)
    m.eval()
    with torch.no_grad():
        y_eval = m(x, return_all_features=True)
    # train mode: DropPath draws x.new_empty((B,1,1)).bernoulli_(keep) per call (eva:121-138)
    m.train()
    torch.manual_seed(77)
    y_train = m(x, return_all_features=True)
    loss = y_train.float().pow(2).mean()
    loss.backward()
    grads = {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}
    # replay the RNG stream to recover the masks (two DropPath calls per block with p>0)
    torch.manual_seed(77)
    rates = [v.item() for v in torch.linspace(0, 0.4, 2)]
    dp = torch.ones(2, 2, 2)
    for i, r in enumerate(rates):
        if r > 0.0:
            for j in range(2):
                keep = 1.0 - r
                mask = torch.empty(2, 1, 1).bernoulli_(keep)
                dp[i, j] = (mask / keep).view(-1)
    _save("eva_vit_tiny.pt", dict(cfg=cfg, state_dict={k: v.detach().clone() for k, v in m.state_dict().items()},
                                  x=x, y_eval=y_eval, y_train=y_train.detach(), loss=loss.detach(),
                                  dp_scales=dp, grads=grads))


def gen_bert():
    """Reference BertForMaskedLM (model/bert.py:1021) at hidden 128 = 2 heads x 64, 2 layers, FFN 256, vocab 1000,
    is_decoder + add_cross_attention, dropout 0 (so train-mode gradients are deterministic):
      (a) text-only, 2-D padding mask            -> last_hidden_state        (vast.py:150-156 'caption_output')
      (b) cross-attention to 37 encoder tokens, 3-D causal mask, labels -> loss, logits, every gradient (vast.py:493-507)"""
    from transformers.models.bert.configuration_bert import BertConfig
    from model.bert import BertForMaskedLM
    cfg = BertConfig(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, max_position_embeddings=64,
                     is_decoder=True, add_cross_attention=True, layer_norm_eps=1e-12, pad_token_id=0)
    torch.manual_seed(0)
    m = BertForMaskedLM(cfg)
    _randomize(m, 3)
    # transformers 4.31 ties decoder.weight to the word embeddings; 5.x does not -> tie explicitly
    m.cls.predictions.decoder.weight = m.bert.embeddings.word_embeddings.weight
    g = torch.Generator().manual_seed(99)
    b, S, Sk = 3, 24, 37
    ids = torch.randint(1, 1000, (b, S), generator=g)
    lens = torch.tensor([24, 17, 9])
    att = (torch.arange(S)[None] < lens[:, None]).long()
    ids = ids * att
    m.eval()
    with torch.no_grad():
        text_only = m.bert(input_ids=ids, attention_mask=att).last_hidden_state
    enc = torch.randn(b, Sk, 128, generator=g, requires_grad=True)
    att3 = att.unsqueeze(1).expand(-1, S, -1).clone()
    att3[:, :S, :S] = torch.tril(att3[:, :S, :S])
    labels = torch.full((b, S), -100, dtype=torch.long)
    pick = torch.rand(b, S, generator=g) < 0.5
    labels[pick & (att > 0)] = ids[pick & (att > 0)]
    m.train()
    out = m(input_ids=ids, attention_mask=att3, encoder_hidden_states=enc, labels=labels)
    out.loss.backward()
    grads = {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    _save("bert_tiny.pt", dict(cfg=dict(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                        intermediate_size=256, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                                        max_position_embeddings=64, layer_norm_eps=1e-12, pad_token_id=0),
                               state_dict=sd, ids=ids, att=att, att3=att3, enc=enc.detach(), labels=labels,
                               text_only=text_only, loss=out.loss.detach(), logits=out.logits.detach(),
                               sequence_output=out.sequence_output.detach(), grads=grads, d_enc=enc.grad.clone()))


def gen_mico_parts():
    """Reference MiCo pieces that run standalone at small sizes (the full class hard-wires the 1.2 B-parameter towers):
    Contra_head / Match_head (model/mico.py:36-52), and the unbound MMGeneralModule methods pool_vision_for_contra,
    get_multimodal_forward_input_{vision,audio} (model/mico.py:157-227) driven through a stub `self` that carries the
    same attributes the real module has; plus the reference collectives' single-process semantics."""
    import torch.nn as nn
    from model import mico as RM
    torch.manual_seed(0)
    vd, md, cd = 176, 128, 64
    ch, mh = RM.Contra_head(vd, cd), RM.Match_head(md)
    _randomize(ch, 1)
    _randomize(mh, 2)

    class Stub(RM.MMGeneralModule):
        pass

    st = Stub()
    st.config = ref_shims.default_model_cfg(pool_video=False)
    st.multimodal_dim = md
    st.hidden_trans_vision_multimodal = nn.Sequential(nn.Linear(vd, md), RM.LayerNorm(md, eps=1e-12))
    st.hidden_trans_audio_multimodal = nn.Sequential(nn.Linear(vd, md), RM.LayerNorm(md, eps=1e-12))
    st.vision_frame_embedding = nn.Parameter(0.02 * torch.randn(1, 8, md))
    st.audio_frame_embedding = nn.Parameter(0.02 * torch.randn(1, 3, md))
    st.vision_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, md))
    st.audio_type_embeddings = nn.Parameter(0.02 * torch.randn(1, 1, md))
    _randomize(st, 3)
    g = torch.Generator().manual_seed(5)
    feat8 = torch.randn(2, 8, 9, vd, generator=g)     # 8-frame video
    feat2 = torch.randn(2, 2, 9, vd, generator=g)     # 2 frames -> nearest-interpolated frame table
    aud3 = torch.randn(2, 3, 9, vd, generator=g)
    with torch.no_grad():
        out = dict(
            pooled=st.pool_vision_for_contra(feat8), contra=ch(st.pool_vision_for_contra(feat8)),
            match=mh(torch.randn(6, md, generator=torch.Generator().manual_seed(6))),
            fuse_v8=st.get_multimodal_forward_input_vision(feat8), fuse_v2=st.get_multimodal_forward_input_vision(feat2),
            fuse_a3=st.get_multimodal_forward_input_audio(aud3))
        st.config.pool_video = True
        out["fuse_v8_pool"] = st.get_multimodal_forward_input_vision(feat8)
    sd = {"contra_head_v." + k: v.detach().clone() for k, v in ch.state_dict().items()}
    sd.update({"itm_head." + k: v.detach().clone() for k, v in mh.state_dict().items()})
    sd.update({k: v.detach().clone() for k, v in st.state_dict().items()})
    _save("mico_parts_tiny.pt", dict(state_dict=sd, feat8=feat8, feat2=feat2, aud3=aud3,
                                     match_in=torch.randn(6, md, generator=torch.Generator().manual_seed(6)), **out))


def gen_transformer():
    """Reference model/transformer.py TransformerEncoder, pre-norm and post-norm, hidden 128 = 2 heads x 64, FFN 256,
    2 layers, dropout 0, additive padding mask (b,1,1,S); forward + gradients of sum(y^2)."""
    from model.transformer import TransformerEncoder
    out = {}
    for mode in ("prenorm", "postnorm"):
        cfg = ref_shims._AttrDict(hidden_size=128, num_attention_heads=2, intermediate_size=256, num_hidden_layers=2,
                                  hidden_dropout=0.0, attention_dropout=0.0, checkpointing=False)
        torch.manual_seed(0)
        m = TransformerEncoder(cfg, mode)
        _randomize(m, 4)
        g = torch.Generator().manual_seed(8)
        x = torch.randn(3, 40, 128, generator=g, requires_grad=True)
        lens = torch.tensor([40, 31, 12])
        mask = ((torch.arange(40)[None] >= lens[:, None]).float() * -10000.0)[:, None, None, :]
        m.train()
        y, _ = m(x, mask)
        y.pow(2).sum().backward()
        out[mode] = dict(state_dict={k: v.detach().clone() for k, v in m.state_dict().items()}, x=x.detach().clone(),
                         mask=mask, y=y.detach().clone(), dx=x.grad.clone(),
                         grads={k: v.grad.clone() for k, v in m.named_parameters()})
    _save("transformer_tiny.pt", out)


def gen_clip():
    """Reference OpenAI-CLIP VisionTransformer (model/clip/clip.py:228) at width 128 = 2 heads x 64, 2 layers, patch 16,
    224x224 -> 197 tokens: forward (all features) and gradients of mean(y^2)."""
    from model.clip.clip import VisionTransformer
    torch.manual_seed(0)
    m = VisionTransformer(input_resolution=224, patch_size=16, width=128, layers=2, heads=2, output_dim=32)
    _randomize(m, 6)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 3, 224, 224, generator=g)
    m.train()
    y = m(x, return_all_features=True)
    y.pow(2).mean().backward()
    with torch.no_grad():
        pooled = m(x)
    _save("clip_vit_tiny.pt", dict(state_dict={k: v.detach().clone() for k, v in m.state_dict().items()}, x=x,
                                   y=y.detach().clone(), pooled=pooled,
                                   grads={k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}))


def gen_fbank():
    """torchaudio.compliance.kaldi.fbank exactly as model/audioprocessor.py:39-40 calls it (the reference's own
    AudioProcessor needs torchaudio.load, which needs torchcodec here -- SURVEY.md 8c -- so the call is made directly on a
    synthetic 16 kHz waveform), for 224 and 64 mel bins."""
    import torchaudio
    g = torch.Generator().manual_seed(4242)
    wave = 0.1 * torch.randn(1, 16000 * 3 + 123, generator=g)        # 3 s clip, ragged tail
    t = torch.arange(wave.shape[1]) / 16000.0
    wave = wave + 0.3 * torch.sin(2 * 3.14159265 * 440.0 * t) + 0.05                # tone + DC offset
    out = {"wave": wave, "torchaudio": torchaudio.__version__}
    for bins in (224, 64):
        out[f"fbank_{bins}"] = torchaudio.compliance.kaldi.fbank(wave * 2 ** 15, num_mel_bins=bins, sample_frequency=16000,
                                                                 frame_length=25, frame_shift=10)
    _save("fbank.pt", out)


def gen_swin():
    """Reference SwinTransformer (model/swin.py:485) at embed_dim 32, depths [2,2], heads [1,2] (head_dim 32 like Swin-B),
    window 7, 56x56 input with patch 4 -> 14x14 tokens -> shifted windows -> patch merge -> 7x7: forward_features and
    gradients of mean(y^2), drop_path 0."""
    from model.swin import SwinTransformer
    torch.manual_seed(0)
    m = SwinTransformer(img_size=56, patch_size=4, in_chans=3, num_classes=0, embed_dim=32, depths=[2, 2], num_heads=[1, 2],
                        window_size=7, mlp_ratio=4., drop_path_rate=0.0)
    _randomize(m, 8)
    with torch.no_grad():
        for n_, prm in m.named_parameters():
            if "relative_position_bias_table" in n_:
                prm.add_(0.5 * torch.randn(prm.shape, generator=torch.Generator().manual_seed(9)))
    g = torch.Generator().manual_seed(33)
    x = torch.randn(3, 3, 56, 56, generator=g)
    m.train()
    y = m.forward_features(x)
    y.pow(2).mean().backward()
    _save("swin_tiny.pt", dict(state_dict={k: v.detach().clone() for k, v in m.state_dict().items()}, x=x,
                               y=y.detach().clone(),
                               grads={k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}))


def gen_adamw():
    """The reference's own optimizer (data/utils/build_optimizer.py: build_optimizer grouping + AdamW.step) and LR schedule
    (data/utils/sched.py:get_lr_sched, applied as in data/utils/pipeline.py:75-78) run UNMODIFIED on CPU for 4 steps over
    a small module whose parameter names hit all of the grouping rules; stores initial parameters, per-step gradients,
    per-step learning rates and the parameters / moments after every step."""
    sys.path.insert(0, os.path.join(ref_shims.REF_ROOT, "data"))
    from utils.build_optimizer import build_optimizer
    from utils.sched import get_lr_sched
    from easydict import EasyDict as edict

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vision_encoder = torch.nn.Module()
            self.vision_encoder.visual = torch.nn.Module()
            self.vision_encoder.visual.proj = torch.nn.Linear(24, 40)            # 'visual' -> clip groups
            self.vision_encoder.visual.LayerNorm = torch.nn.LayerNorm(40)
            self.multimodal_encoder = torch.nn.Linear(40, 523)                   # basic groups (two chunks, ragged tail)
            self.LayerNorm = torch.nn.LayerNorm(7)                               # basic no-decay
            self.contra_head_new = torch.nn.Linear(40, 8)                        # 'new' groups

    torch.manual_seed(3)
    model = Tiny()
    args = edict(model_cfg=edict(vision_encoder_type='evaclip01_giant'),
                 run_cfg=edict(optim='adamw', learning_rate=1e-3, clip_lr=5e-4, new_lr=2e-3, betas=[0.9, 0.98],
                               weight_decay=0.01, new_params_name=['contra_head_new'], warmup_ratio=0.25,
                               num_train_steps=8, scheduler='warmup_linear'))
    opt = build_optimizer(model, args, None)
    names = [k for k, _ in model.named_parameters()]
    init = {k: v.detach().clone() for k, v in model.named_parameters()}
    gen = torch.Generator().manual_seed(11)
    steps = []
    for step in range(1, 5):
        lr_ratio = get_lr_sched(step, args.run_cfg)
        for pg in opt.param_groups:
            pg['lr'] = pg['init_lr'] * lr_ratio
        grads = {}
        for k, v in model.named_parameters():
            g = torch.randn(v.shape, generator=gen) * (0.1 if step != 3 else 10.0)
            v.grad = g.clone()
            grads[k] = g
        opt.step()
        steps.append(dict(lr_ratio=lr_ratio, lrs=[pg['lr'] for pg in opt.param_groups], grads=grads,
                          params={k: v.detach().clone() for k, v in model.named_parameters()},
                          exp_avg={k: opt.state[v]['exp_avg'].clone() for k, v in model.named_parameters()},
                          exp_avg_sq={k: opt.state[v]['exp_avg_sq'].clone() for k, v in model.named_parameters()}))
    groups = {}
    for gi, pg in enumerate(opt.param_groups):
        for prm in pg['params']:
            for k, v in model.named_parameters():
                if v is prm:
                    groups[k] = gi
    _save("adamw.pt", dict(names=names, init=init, steps=steps, groups=groups, run_cfg=dict(args.run_cfg),
                           group_cfg=[dict(weight_decay=pg['weight_decay'], init_lr=pg['init_lr'], betas=tuple(pg['betas']),
                                           eps=pg['eps'], correct_bias=pg['correct_bias']) for pg in opt.param_groups]))


def gen_checkpoint():
    """MMGeneralModule.modify_checkpoint (model/mico.py:250-321) run UNMODIFIED through a stub `self` on synthetic
    checkpoints: legacy key names ('video', 'evaclip_model', 'clip_model'), fp16 leaves, frame embeddings of the wrong
    length (nearest resize) and position embeddings of another grid (bilinear resize), for the evaclip and the clip
    branch."""
    import types
    from model.mico import MMGeneralModule
    from easydict import EasyDict as edict
    g = torch.Generator().manual_seed(5)
    r = lambda *s: torch.randn(*s, generator=g)
    cases = {}
    ck_eva = {"evaclip_model.visual.pos_embed": r(1, 1 + 16, 8), "evaclip_model.visual.patch_embed.proj.weight": r(8, 3, 14, 14),
              "video_frame_embedding": r(1, 4, 6), "audio_frame_embedding": r(1, 2, 6), "contra_head_t.linear.weight": r(4, 6).half(),
              "multimodal_encoder.bert.embeddings.word_embeddings.weight": r(10, 6).half(), "video_type_embeddings": r(1, 1, 6)}
    cfg_eva = dict(frame_embedding_type='adaptive', max_vision_sample_num=8, max_audio_sample_num=3,
                   vision_encoder_type='evaclip01_giant', vision_resolution=98)
    ck_clip = {"clip_model.visual.positional_embedding": r(1 + 9, 8), "clip_model.visual.conv1.weight": r(8, 3, 16, 16),
               "vision_perceiver.video_frame_embedding": r(1, 8, 6), "itm_head.linear1.bias": r(6).half()}
    cfg_clip = dict(frame_embedding_type='adaptive', max_vision_sample_num=3, max_audio_sample_num=1,
                    vision_encoder_type='clip_vit_base_16', vision_resolution=80)
    for name, ck, cfg in (("evaclip", ck_eva, cfg_eva), ("clip", ck_clip, cfg_clip)):
        stub = types.SimpleNamespace(config=edict(cfg))
        out = MMGeneralModule.modify_checkpoint(stub, {k: v.clone() for k, v in ck.items()})
        cases[name] = dict(inp=ck, cfg=cfg, out=dict(out))
    _save("modify_checkpoint.pt", cases)


def _loss_worker(rank, world, port, q):
    """Body of gen_losses on one rank (in-process for world 1, spawned for world 2)."""
    _data = os.path.join(ref_shims.REF_ROOT, "data")
    while _data in sys.path:          # a spawned child inherits the parent's sys.path: keep `model` = <reference>/model
        sys.path.remove(_data)
    ref_shims.install()

    import importlib
    import importlib.util
    import random
    import types
    import numpy as np
    import torch.distributed as dist
    from transformers.models.bert.configuration_bert import BertConfig
    from model.bert import BertForMaskedLM
    DATA = os.path.join(ref_shims.REF_ROOT, "data")
    if DATA not in sys.path:
        sys.path.append(DATA)          # behind the reference root: only `utils.*` resolves here
    spec = importlib.util.spec_from_file_location("refdata_model", os.path.join(DATA, "model", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(DATA, "model")])
    sys.modules["refdata_model"] = importlib.util.module_from_spec(spec)       # package shell: data/model/ without its __init__
    V = importlib.import_module("refdata_model.vast")
    GM = importlib.import_module("refdata_model.general_module")
    edict = sys.modules["easydict"].EasyDict
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    cfg = BertConfig(vocab_size=1000, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                     hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, max_position_embeddings=64,
                     is_decoder=True, add_cross_attention=True, layer_norm_eps=1e-12, pad_token_id=0)
    torch.manual_seed(0)
    bert = BertForMaskedLM(cfg)
    _randomize(bert, 3)
    bert.cls.predictions.decoder.weight = bert.bert.embeddings.word_embeddings.weight
    bert.train()
    itm_head = GM.Match_head(128)
    _randomize(itm_head, 4)
    contra_temp = torch.nn.Parameter(torch.tensor(0.07))
    masker = GM.TokenMasker(mask_token=103, range_start=106, range_end=1000)

    g = torch.Generator().manual_seed(21 + rank)
    b, S, Sk, cd = 4, 16, 9, 32
    lens = torch.tensor([16, 11, 7, 13]) - 2 * rank
    att = (torch.arange(S)[None] < lens[:, None]).long()
    ids = torch.randint(106, 1000, (b, S), generator=g) * att
    ids[:, 0] = 101
    raw_t = torch.randn(b, cd, generator=g, requires_grad=True)
    raw_v = torch.randn(b, cd, generator=g, requires_grad=True)
    cond = (0.5 * torch.randn(b, Sk, 128, generator=g)).requires_grad_(True)

    recorded = dict(multinomial=[], masked=None)
    real_multinomial, real_cuda, real_half = torch.multinomial, torch.Tensor.cuda, torch.Tensor.half

    def rec_multinomial(w, n, *a, **k):
        out = real_multinomial(w, n, *a, **k)
        recorded["multinomial"].append(int(out.item()))
        return out

    def rec_masker(tokens, prob):
        out = masker(tokens, prob)
        recorded["masked"] = (out[0].clone(), out[1].clone())
        return out

    stub = types.SimpleNamespace(contra_temp=contra_temp, itm_ratio=0.1, itm_head=itm_head, multimodal_encoder=bert,
                                 text_masker=rec_masker, config=edict(captioner_mode=False))
    stub.batch_get = lambda batch, key: batch[key]
    torch.multinomial = rec_multinomial
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.half = lambda self, *a, **k: self
    try:
        batch = edict(raw_captions=["a"] * b)
        batch["feat_t"] = torch.nn.functional.normalize(raw_t, dim=-1)
        batch["feat_v"] = torch.nn.functional.normalize(raw_v, dim=-1)
        batch["condition_feats_v"] = cond
        batch["caption_tokens"] = edict(input_ids=ids, attention_mask=att)
        torch.manual_seed(5 + rank)
        random.seed(5 + rank)
        np.random.seed(5 + rank)
        ret = V.VAST.forward_ret(stub, batch, "ret%tv", compute_loss=True)
        cap = V.VAST.forward_cap(stub, batch, "cap%tv", compute_loss=True)
    finally:
        torch.multinomial, torch.Tensor.cuda, torch.Tensor.half = real_multinomial, real_cuda, real_half
    total = ret["loss_itc"] + ret["loss_itm"] + cap["loss_cap"]
    total.backward()
    neg = recorded["multinomial"]
    sd = {"multimodal_encoder." + k: v.detach().clone() for k, v in bert.state_dict().items()}
    sd.update({"itm_head." + k: v.detach().clone() for k, v in itm_head.state_dict().items()})
    sd["contra_temp"] = contra_temp.detach().clone()
    keep = ("embeddings.word_embeddings.weight", "layer.0.attention.self.query.weight", "layer.1.crossattention.self.key.weight",
            "layer.1.output.dense.bias", "cls.predictions.transform.dense.weight", "cls.predictions.bias")
    grads = {"multimodal_encoder." + k: v.grad.clone() for k, v in bert.named_parameters()
             if v.grad is not None and k.endswith(keep)}
    grads.update({"itm_head." + k: v.grad.clone() for k, v in itm_head.named_parameters()})
    grads["contra_temp"] = contra_temp.grad.clone()
    res = dict(state_dict=sd, ids=ids, att=att, raw_t=raw_t.detach(), raw_v=raw_v.detach(), cond=cond.detach(),
               neg_c=torch.tensor(neg[:b]), neg_t=torch.tensor(neg[b:2 * b]),
               cap_ids=recorded["masked"][0], cap_labels=recorded["masked"][1],
               loss_itc=ret["loss_itc"].detach(), loss_itm=ret["loss_itm"].detach(),
               loss_cap=cap["loss_cap"].detach(), grads=grads, d_raw_t=raw_t.grad.clone(),
               d_raw_v=raw_v.grad.clone(), d_cond=cond.grad.clone(), layers=2, heads=2, itm_ratio=0.1)
    dist.barrier()
    dist.destroy_process_group()
    if q is not None:                      # spawned: hand the result over through a file (q carries the directory)
        res.pop("state_dict")              # identical on every rank and to losses_tiny.pt (same seeds)
        torch.save(res, os.path.join(q, f"rank{rank}.pt"))
    return res


def gen_losses():
    """data/model/vast.py forward_ret (:383-462: ITC with label smoothing and the learned temperature, hard-negative ITM through
    the cross-attention BERT) and forward_cap (:485-512: TokenMasker, causal 3-D mask, masked-LM loss) run UNMODIFIED as unbound
    methods over a stub `self` that carries the attributes the real VAST module has (reference BertForMaskedLM, reference
    Match_head, contra_temp, itm_ratio, text_masker), on one gloo rank.  The VAST class itself cannot be constructed here
    (weight files); its loss code can.  Process-wide patches for this generator only: Tensor.cuda() and
    Tensor.half() are identities (CPU fp32 run; the reference casts the [CLS] state to fp16 before the ITM head, vast.py:452),
    torch.multinomial is wrapped to RECORD the sampled hard negatives, and the masker's output is recorded -- the fixture
    stores those discrete choices so that oracle and product replay them."""
    _save("losses_tiny.pt", _loss_worker(0, 1, 29577, None))
    # the same on TWO gloo ranks: gathered negatives, all_gather_with_grad of the fusion inputs, per-rank batches
    import tempfile
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    with tempfile.TemporaryDirectory() as tmp:
        procs = [ctx.Process(target=_loss_worker, args=(r, 2, 29578, tmp)) for r in range(2)]
        for p_ in procs:
            p_.start()
        for p_ in procs:
            p_.join(timeout=900)
        ranks = [torch.load(os.path.join(tmp, f"rank{r}.pt"), weights_only=False) for r in range(2)]
    _save("losses_2rank.pt", dict(ranks=ranks))


def gen_imageproc():
    """The reference's own ImageProcessor (model/imageprocessor.py, image_transforms='none') on synthetic PNG files: a 300x400
    photo-like image down to 224 (anti-aliased by the installed torchvision's Resize default) with the evaclip and the swin
    normalisation, and a 97x150 image UP to 224.  Also stores torchvision's Resize(antialias=False) output for the first
    image: the behaviour of the reference's pinned torchvision 0.15.2."""
    import tempfile
    import numpy as np
    from PIL import Image
    from torchvision import transforms
    from model.imageprocessor import ImageProcessor
    rng = np.random.RandomState(7)
    yy, xx = np.mgrid[0:300, 0:400]
    base = (127 + 90 * np.sin(xx / 23.0)[..., None] * np.cos(yy / 17.0)[..., None] * np.array([1.0, 0.6, -0.8])).clip(0, 255)
    big = (base + rng.randint(-30, 30, (300, 400, 3))).clip(0, 255).astype(np.uint8)
    small = rng.randint(0, 256, (97, 150, 3)).astype(np.uint8)
    out = dict(big=torch.from_numpy(big), small=torch.from_numpy(small))
    with tempfile.TemporaryDirectory() as tmp:
        for name, arr in (("big", big), ("small", small)):
            Image.fromarray(arr).save(os.path.join(tmp, name + ".png"))
        out["big_evaclip"] = ImageProcessor(224, "evaclip01_giant", training=True)(os.path.join(tmp, "big.png"))
        out["big_swin"] = ImageProcessor(224, "swin_base_22k_224", training=False)(os.path.join(tmp, "big.png"))
        out["small_evaclip"] = ImageProcessor(224, "evaclip01_giant", training=True)(os.path.join(tmp, "small.png"))
    proc = ImageProcessor(224, "evaclip01_giant")
    t = transforms.ToTensor()(Image.fromarray(big))
    out["big_evaclip_noaa"] = transforms.Normalize(proc.mean, proc.std)(transforms.Resize((224, 224), antialias=False)(t)).unsqueeze(0)
    import torchvision
    out["torchvision"] = torchvision.__version__
    # model/videoprocessor.py: the frame-segment sampler (the module imports decord at the top; only that import is stubbed)
    import types
    sys.modules.setdefault("decord", types.SimpleNamespace(VideoReader=None))
    from model.videoprocessor import split as ref_split
    out["video_split"] = {(n, k): ref_split(list(range(n)), k) for n in (1, 3, 7, 8, 30, 31, 257) for k in (1, 3, 4, 8)}
    _save("imageproc.pt", out)


def gen_eva02():
    """Reference EVAVisionTransformer in its EVA02 configuration (rope, naiveswiglu, subln; the xformers branch is not installable
    here, so xattn=False: the reference's own plain-attention path) at width 128 = 2 heads x 64, depth 2, mlp_ratio 2.6667
    (hidden 341), patch 14, 224x224 -> 257 tokens, pt_hw_seq_len 16 with interpolated frequencies: eval forward and every
    gradient of mean(y^2).  Parity target for the EVA02 CUDA tower of a later round (oracle/eva02.py)."""
    from functools import partial
    from model.evaclip.eva_vit_model import EVAVisionTransformer
    from model.evaclip.transformer import LayerNorm
    cfg = dict(width=128, depth=2, heads=2, patch=14, image=224, eps=1e-6, mlp_ratio=2.6667, pt_hw_seq_len=16)
    torch.manual_seed(0)
    m = EVAVisionTransformer(img_size=224, patch_size=14, num_classes=8, use_mean_pooling=False, embed_dim=128, depth=2,
                             num_heads=2, mlp_ratio=2.6667, qkv_bias=True, drop_path_rate=0.0,
                             norm_layer=partial(LayerNorm, eps=1e-6), xattn=False, rope=True, pt_hw_seq_len=16, intp_freq=True,
                             naiveswiglu=True, subln=True)
    _randomize(m, 2)
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 3, 224, 224, generator=g)
    m.eval()
    y = m(x, return_all_features=True)
    y.pow(2).mean().backward()
    keep = ("blocks.0.attn.q_proj.weight", "blocks.0.attn.k_proj.weight", "blocks.1.attn.q_bias", "blocks.0.attn.inner_attn_ln.bias",
            "blocks.1.mlp.w1.weight", "blocks.1.mlp.w2.weight", "blocks.1.mlp.ffn_ln.weight", "blocks.0.mlp.w3.bias", "pos_embed",
            "cls_token", "norm.weight")
    grads = {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None and k in keep}
    sd = {k: v.detach().clone() for k, v in m.state_dict().items() if ".rope." not in k and not k.startswith("head.")}
    _save("eva02_tiny.pt", dict(cfg=cfg, state_dict=sd, x=x, y=y.detach(), grads=grads))


def _dist_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ref_shims.REF_ROOT, "data"))
    from utils.distributed import all_gather_with_grad, concat_all_gather   # the reference's own collectives
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(3, 5, generator=g, requires_grad=True)
    ids = torch.randint(0, 50, (3, 4), generator=g)
    gathered = all_gather_with_grad(x)
    ids_all = concat_all_gather(ids)
    w = torch.arange(1, gathered.numel() + 1, dtype=torch.float32).view_as(gathered) * (rank + 1)
    (gathered * w).sum().backward()
    q.put((rank, x.detach(), ids, gathered.detach(), ids_all, x.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def gen_dist():
    """data/utils/distributed.py:12-66 run UNMODIFIED on 2 gloo ranks: all_gather_with_grad (forward all-gather, backward
    all-reduce(SUM) + own slice) and concat_all_gather."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, 29613, q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p_ in procs:
        p_.join()
    _save("dist_gather_2rank.pt", dict(x=[r[1] for r in res], ids=[r[2] for r in res], gathered=[r[3] for r in res],
                                       ids_all=[r[4] for r in res], x_grad=[r[5] for r in res]))


def gen_swin_b():
    """Reference SwinTransformer at Swin-B size (swin_base_patch4_window7_224_22k.yaml: embed 128, depths [2,2,18,2], heads
    [4,8,16,32], window 7, 224 x 224), batch 2, drop_path 0: forward_features, a selection of small gradient tensors and the
    norm of EVERY parameter gradient of mean(y^2).  Parameters come from oracle.fullsize.seeded_params_ (regenerated, not
    stored: 88 M values)."""
    from model.swin import SwinTransformer
    from oracle import fullsize as FS
    m = FS.seeded_params_(SwinTransformer(**FS.SWIN_B), seed=5)
    x = FS.swin_b_input()
    m.train()
    y = m.forward_features(x)
    y.pow(2).mean().backward()
    named = dict(m.named_parameters())
    _save("swin_b.pt", dict(y=y.detach().clone(), grads={k: named[k].grad.clone() for k in FS.SWIN_B_GRAD_KEYS},
                            grad_norms={k: float(v.grad.norm()) for k, v in named.items() if v.grad is not None},
                            n_params=sum(v.numel() for v in named.values())))


def gen_generation():
    """The reference's own decode-step hooks (model/bert.py:1110-1143 update_attention_mask / update_position_ids /
    prepare_inputs_for_generation, :1145-1190 _update_model_kwargs_for_generation), called UNBOUND over a stub `self`
    (HF GenerationMixin of transformers 4.31 is not importable here, the hooks themselves are plain functions): inputs and
    outputs of one and of several consecutive decode steps -> tests/golden/generation_steps.pt."""
    import types
    from model.bert import BertForMaskedLM as R
    stub = types.SimpleNamespace(tokenizer=types.SimpleNamespace(mask_token_id=103))
    stub.update_attention_mask = lambda m: R.update_attention_mask(stub, m)
    stub.update_position_ids = lambda p: R.update_position_ids(stub, p)
    stub._extract_past_from_model_output = lambda outputs, standardize_cache_format=False: None
    g = torch.Generator().manual_seed(5)
    cases = []
    for b, n in ((3, 1), (2, 4), (4, 7)):
        lens = torch.randint(1, n + 1, (b,), generator=g)
        att = (torch.arange(n)[None] < lens[:, None]).long()
        mask = torch.tril(att.unsqueeze(1).expand(-1, n, -1).clone())
        ids = torch.randint(5, 1000, (b, n), generator=g)
        pos = torch.arange(n)[None].expand(b, -1).clone()
        enc = torch.randn(b, 5, 8, generator=g)
        prep = R.prepare_inputs_for_generation(stub, ids, attention_mask=mask, position_ids=pos, encoder_hidden_states=enc)
        steps = []
        kw = {"attention_mask": mask, "position_ids": pos}
        for _ in range(3):      # the kwargs mask after each generated token (greedy / beam loops call this once per step)
            kw = R._update_model_kwargs_for_generation(stub, None, dict(kw))
            steps.append(dict(attention_mask=kw["attention_mask"].clone(), position_ids=kw["position_ids"].clone()))
        cases.append(dict(ids=ids, mask=mask, pos=pos, enc=enc, prep_input_ids=prep["input_ids"], prep_mask=prep["attention_mask"],
                          prep_pos=prep["position_ids"], kwargs_steps=steps, mask_token_id=103))
    _save("generation_steps.pt", dict(cases=cases))


GENERATORS = {"generation": gen_generation, "swin_b": gen_swin_b, "vit": gen_vit, "bert": gen_bert, "mico_parts": gen_mico_parts, "dist": gen_dist,
              "transformer": gen_transformer, "clip": gen_clip, "fbank": gen_fbank, "swin": gen_swin,
              "adamw": gen_adamw, "checkpoint": gen_checkpoint, "losses": gen_losses, "imageproc": gen_imageproc, "eva02": gen_eva02}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    ref_shims.install()
    import transformers
    for name, fn in GENERATORS.items():
        if args.only and args.only != name:
            continue
        print(f"== {name}")
        fn()
    manifest = dict(torch=torch.__version__, transformers=transformers.__version__,
                    reference="invictus717/MiCo @ 831847f (/root/reference)",
                    note="fixtures generated by oracle/make_golden.py running the unmodified reference on CPU")
    with open(os.path.join(GOLD, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
