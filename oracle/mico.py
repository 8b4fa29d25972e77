"""Oracle: MiCo heads, pooling, fusion inputs and the retrieval / caption training step, fp32 CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
  * Contra_head / Match_head                               model/mico.py:36-52
  * pool_{vision,audio,depth,text}_for_contra               model/mico.py:157-185
  * get_multimodal_forward_input_{vision,audio,depth}       model/mico.py:187-241 (Linear+LN(1e-12), + frame emb, + type emb)
  * forward_{vision,audio}_encoder                          model/mico.py:115-148 (frames folded into the batch; audio = 3x channel repeat)
  * batch_get features                                      data/model/vast.py:209-272 (feat_* = normalize(head(pool(.))))
  * ITC loss                                                data/model/vast.py:394-417
  * ITM loss                                                data/model/vast.py:419-457 (negatives passed in explicitly)
  * caption loss                                            data/model/vast.py:493-510
  * concat_all_gather / all_gather_with_grad                data/utils/distributed.py:12-66 (world given as lists of per-rank tensors)
over the reference's own state_dict key names.  PINNED: tests/golden/losses_tiny.pt and losses_2rank.pt hold the outputs of the
reference's own data/model/vast.py forward_ret / forward_cap (run unmodified over a stub module on 1 and 2 gloo ranks,
oracle/make_golden.py:gen_losses); tests/test_oracle_golden.py checks itc_loss / itm_loss / caption_loss against them.  The multi-rank step is simulated in ONE process by evaluating every rank's
tensors and concatenating them (all_gather) -- autograd through the concatenation reproduces GatherLayer's backward
(all-reduce(SUM) of the stacked gradients, own slice) when the per-rank losses are SUMMED, which is what a DDP-less loop
computes (pipeline.py:93-99 sums gradients without dividing).
"""
import torch
import torch.nn.functional as F

from . import bert as OB
from . import eva_vit as OV


def contra_head(p, name, x):
    return F.linear(x, p[name + ".linear.weight"])


def match_head(p, x, name="itm_head"):
    h = F.gelu(F.linear(x, p[name + ".linear1.weight"], p[name + ".linear1.bias"]))
    h = F.layer_norm(h, (h.shape[-1],), p[name + ".layernorm.weight"], p[name + ".layernorm.bias"], 1e-12)
    return F.linear(h, p[name + ".linear2.weight"], p[name + ".linear2.bias"])


def vision_encoder(p, pixels, vit_cfg, dp_scales=None):
    b, n = pixels.shape[:2]
    y = OV.forward_features(p, pixels.reshape(b * n, *pixels.shape[2:]), vit_cfg, prefix="vision_encoder.visual.",
                            dp_scales=dp_scales)
    return y.reshape(b, n, *y.shape[-2:])


def audio_encoder(p, spec, vit_cfg):
    return vision_encoder(p, spec.unsqueeze(2).repeat(1, 1, 3, 1, 1), vit_cfg)


def pool_tower(feature):
    return feature[:, :, 0].mean(dim=1)


def fusion_input(p, out, kind, pool_video=False):
    """kind in {vision, audio, depth}"""
    b, n, x, c = out.shape
    if pool_video:
        out = torch.cat([out[:, :, 0:1], out[:, :, 1:].mean(2, keepdim=True)], dim=2)
    t = f"hidden_trans_{kind}_multimodal."
    out = F.linear(out, p[t + "0.weight"], p[t + "0.bias"])
    out = F.layer_norm(out, (out.shape[-1],), p[t + "1.weight"], p[t + "1.bias"], 1e-12)
    fe = p[f"{kind}_frame_embedding"]
    if n != fe.shape[1]:
        fe = F.interpolate(fe.permute(0, 2, 1), n, mode="nearest").permute(0, 2, 1)
    out = out + fe.unsqueeze(-2)
    out = out.reshape(b, -1, out.shape[-1])
    return out + p[f"{kind}_type_embeddings"]


def text_feature(p, ids, att, layers, heads):
    h = OB.bert_model(p, ids, att, prefix="multimodal_encoder.bert.", layers=layers, heads=heads)
    return F.normalize(contra_head(p, "contra_head_t", h[:, 0]), dim=-1)


def itc_loss(feat_c, feat_t, feat_c_all, feat_t_all, temp, rank):
    bs = feat_t.shape[0]
    sim_c2t = feat_c @ feat_t_all.t() / temp
    sim_t2c = feat_t @ feat_c_all.t() / temp
    tgt = torch.arange(rank * bs, rank * bs + bs, device=feat_t.device)
    loss = (F.cross_entropy(sim_c2t, tgt, label_smoothing=0.1) + F.cross_entropy(sim_t2c, tgt, label_smoothing=0.1)) / 2
    return loss, sim_c2t, sim_t2c


def itm_loss(p, cond, cond_all, ids, att, ids_all, att_all, neg_c, neg_t, itm_ratio, layers, heads):
    bs = cond.shape[0]
    ids_1 = torch.cat((ids, ids, ids_all[neg_t]), 0)
    att_1 = torch.cat((att, att, att_all[neg_t]), 0)
    cond_3 = torch.cat((cond, cond_all[neg_c], cond), 0)
    out = OB.bert_model(p, ids_1, att_1, cond_3, None, prefix="multimodal_encoder.bert.", layers=layers, heads=heads)
    logits = match_head(p, out[:, 0])
    truth = torch.zeros(bs * 3, dtype=torch.long, device=logits.device)
    truth[:bs] = 1
    return itm_ratio * F.cross_entropy(logits, truth)


def caption_loss(p, cond, ids_masked, att, labels, layers, heads):
    S = att.shape[1]
    att3 = torch.tril(att.unsqueeze(1).expand(-1, S, -1).clone())
    loss, _, _ = OB.masked_lm(p, ids_masked, att3, cond, None, labels, layers=layers, heads=heads, prefix="multimodal_encoder.")
    return loss


def retrieval_caption_step(p, ranks, vit_cfg, layers, heads, itm_ratio=0.1, task="ret%tv_cap%tv"):
    """One 'ret%tv[_cap%tv]' step for every rank of a simulated world.  `ranks` is a list of per-rank dicts with
    pixels (b,n,3,H,W), ids, att, neg_c, neg_t [, cap_ids, cap_labels].  Returns per-rank dicts of losses."""
    feats_t = [text_feature(p, r["ids"], r["att"], layers, heads) for r in ranks]
    vis = [vision_encoder(p, r["pixels"], vit_cfg) for r in ranks]
    feats_v = [F.normalize(contra_head(p, "contra_head_v", pool_tower(v)), dim=-1) for v in vis]
    conds = [fusion_input(p, v, "vision") for v in vis]
    ft_all = torch.cat(feats_t).detach()          # concat_all_gather: no gradient (distributed.py:53-66)
    fv_all = torch.cat(feats_v).detach()
    ids_all = torch.cat([r["ids"] for r in ranks])
    att_all = torch.cat([r["att"] for r in ranks])
    cond_all = torch.cat(conds)                    # all_gather_with_grad
    out = []
    for k, r in enumerate(ranks):
        d = {}
        l_itc, _, _ = itc_loss(feats_v[k], feats_t[k], fv_all, ft_all, p["contra_temp"], k)
        d["loss_itc"] = l_itc
        d["loss_itm"] = itm_loss(p, conds[k], cond_all, r["ids"], r["att"], ids_all, att_all, r["neg_c"], r["neg_t"],
                                 itm_ratio, layers, heads)
        if "cap" in task:
            d["loss_cap"] = caption_loss(p, conds[k], r["cap_ids"], r["att"], r["cap_labels"], layers, heads)
        out.append(d)
    return out


# ---------------------------------------------------------------------------------------------- omni-modal step (one rank)
_COMBO = {"v": "v", "a": "a", "d": "d", "va": "va", "id": "vd"}
_KIND = {"v": "vision", "a": "audio", "d": "depth"}


def omni_features(p, batch, vit_cfg, need="vad", dp_scales=None):
    """Tower outputs, pooled features and fusion inputs of every modality in `need` (mico.py:115-148, 157-248):
    vision_pixels (b,n,3,H,W); audio_spectrograms (b,n,T,mel) -> 3x channel repeat; depth_pixels (b,n,3,H,W)."""
    outs = {}
    if "v" in need:
        outs["v"] = vision_encoder(p, batch["vision_pixels"], vit_cfg, dp_scales=dp_scales)
    if "a" in need:
        outs["a"] = audio_encoder(p, batch["audio_spectrograms"], vit_cfg)
    if "d" in need:
        outs["d"] = vision_encoder(p, batch["depth_pixels"], vit_cfg)
    pools = {m: pool_tower(o) for m, o in outs.items()}
    conds = {m: fusion_input(p, o, _KIND[m]) for m, o in outs.items()}
    return outs, pools, conds


def combo_feature(p, pools, combo):
    """feat_<combo> of data/model/vast.py:221-272: single modality -> Contra_head; fused -> biased Linear on the concatenation
    (contra_head_va / contra_head_id, mico.py:386-394)."""
    parts = _COMBO[combo]
    if len(parts) == 1:
        return F.normalize(contra_head(p, f"contra_head_{combo}", pools[parts]), dim=-1)
    x = torch.cat([pools[m] for m in parts], dim=1)
    return F.normalize(F.linear(x, p[f"contra_head_{combo}.weight"], p[f"contra_head_{combo}.bias"]), dim=-1)


def omni_step(p, batch, vit_cfg, layers, heads, task, itm_ratio=0.1, negs=None, generator=None):
    """vast.py:317-348 (task dispatch) -> forward_ret :383-464 + forward_cap :485-512 for ONE rank (world size 1) over the
    modality combinations tv / ta / tva / td / tid.  batch: vision_pixels, audio_spectrograms, depth_pixels, ids, att
    [, cap_ids, cap_labels]; negs: {subtask: (neg_cond, neg_text)} to pin the hard negatives, else they are drawn like
    vast.py:430-440 (torch.multinomial on softmax(sim) + 1e-4 with the own-sample column zeroed)."""
    ret_st, cap_st = [], []
    for t in task.split("_"):
        (ret_st if t.startswith("ret") else cap_st).extend(t.split("%")[1:])
    need = set("".join(_COMBO[s[1:]] for s in ret_st + cap_st))
    _, pools, conds = omni_features(p, batch, vit_cfg, need)
    cond_of = lambda c: torch.cat([conds[m] for m in _COMBO[c]], dim=1)
    ids, att = batch["ids"], batch["att"]
    out = {}
    if ret_st:
        feat_t = text_feature(p, ids, att, layers, heads)
        bs = feat_t.shape[0]
        l_itc, l_itm = [], []
        for st in ret_st:
            c = st[1:]
            feat_c = combo_feature(p, pools, c)
            l, sim_c2t, sim_t2c = itc_loss(feat_c, feat_t, feat_c.detach(), feat_t.detach(), p["contra_temp"], 0)
            l_itc.append(l)
            if negs is not None and st in negs:
                neg_c, neg_t = negs[st]
            else:
                with torch.no_grad():
                    w_t2c = F.softmax(sim_t2c, dim=1) + 1e-4
                    w_t2c.fill_diagonal_(0)
                    w_c2t = F.softmax(sim_c2t, dim=1) + 1e-4
                    w_c2t.fill_diagonal_(0)
                    neg_c = torch.multinomial(w_t2c, 1, generator=generator).view(-1)
                    neg_t = torch.multinomial(w_c2t, 1, generator=generator).view(-1)
            cond = cond_of(c)
            l_itm.append(itm_loss(p, cond, cond, ids, att, ids, att, neg_c, neg_t, itm_ratio, layers, heads))
        out["loss_itc"] = sum(l_itc) / len(l_itc)
        out["loss_itm"] = sum(l_itm) / len(l_itm)
    if cap_st:
        l_cap = [caption_loss(p, cond_of(st[1:]), batch["cap_ids"], att, batch["cap_labels"], layers, heads) for st in cap_st]
        out["loss_cap"] = sum(l_cap) / len(l_cap)
    return out
