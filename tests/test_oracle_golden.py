"""CPU: the oracle restatement (oracle/eva_vit.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py ran model/evaclip/eva_vit_model.py on CPU in the build container).
fp32 vs fp32 on the same weights/inputs: tolerance 1e-5 rel-L2 (summation order only)."""
import os

import torch

from conftest import rel_l2
from oracle import eva_vit as O


def _load(golden_dir):
    return torch.load(os.path.join(golden_dir, "eva_vit_tiny.pt"), weights_only=False)


def test_vit_eval_forward_matches_reference(golden_dir):
    g = _load(golden_dir)
    y = O.forward_features(g["state_dict"], g["x"], g["cfg"])
    assert y.shape == g["y_eval"].shape == (2, 257, 176)
    assert rel_l2(y, g["y_eval"]) < 1e-5


def test_vit_train_forward_backward_matches_reference(golden_dir):
    g = _load(golden_dir)
    p = {k: v.clone().requires_grad_(True) for k, v in g["state_dict"].items()}
    y = O.forward_features(p, g["x"], g["cfg"], dp_scales=g["dp_scales"])
    assert rel_l2(y, g["y_train"]) < 1e-5
    loss = y.float().pow(2).mean()
    assert abs(loss.item() - g["loss"].item()) <= 1e-5 * abs(g["loss"].item())
    loss.backward()
    for k, ref in g["grads"].items():
        if k.startswith("head."):
            continue   # the classifier head is not on the return_all_features path (mico.py:120)
        assert p[k].grad is not None, k
        assert rel_l2(p[k].grad, ref) < 2e-5, k


def test_drop_path_rates_match_linspace():
    assert O.drop_path_rates(40)[0] == 0.0
    assert abs(O.drop_path_rates(40)[-1] - 0.4) < 1e-7


# ---------------------------------------------------------------------------------------------- BERT
def _load_bert(golden_dir):
    return torch.load(os.path.join(golden_dir, "bert_tiny.pt"), weights_only=False)


def test_bert_text_only_matches_reference(golden_dir):
    from oracle import bert as OB
    g = _load_bert(golden_dir)
    h = OB.bert_model(g["state_dict"], g["ids"], g["att"], layers=2, heads=2)
    assert rel_l2(h, g["text_only"]) < 1e-5


def test_bert_cross_attention_caption_loss_and_grads_match_reference(golden_dir):
    from oracle import bert as OB
    g = _load_bert(golden_dir)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in g["state_dict"].items()}
    # tied decoder / word embeddings (transformers 4.31 behaviour): one leaf under both names
    p["cls.predictions.decoder.weight"] = p["bert.embeddings.word_embeddings.weight"]
    enc = g["enc"].clone().requires_grad_(True)
    loss, logits, seq = OB.masked_lm(p, g["ids"], g["att3"], enc, None, g["labels"], layers=2, heads=2)
    assert rel_l2(seq, g["sequence_output"]) < 1e-5
    assert rel_l2(logits, g["logits"]) < 1e-5
    assert abs(loss.item() - g["loss"].item()) < 1e-5 * abs(g["loss"].item())
    loss.backward()
    assert rel_l2(enc.grad, g["d_enc"]) < 2e-5
    for k, ref in g["grads"].items():
        if k == "cls.predictions.decoder.weight":
            continue
        assert rel_l2(p[k].grad, ref) < 5e-5, k


# ---------------------------------------------------------------------------------------------- MiCo heads / fusion inputs
def test_mico_parts_match_reference(golden_dir):
    from oracle import mico as OM
    g = torch.load(os.path.join(golden_dir, "mico_parts_tiny.pt"), weights_only=False)
    p = g["state_dict"]
    pooled = OM.pool_tower(g["feat8"])
    assert rel_l2(pooled, g["pooled"]) < 1e-6
    assert rel_l2(OM.contra_head(p, "contra_head_v", pooled), g["contra"]) < 1e-5
    assert rel_l2(OM.match_head(p, g["match_in"]), g["match"]) < 1e-5
    assert rel_l2(OM.fusion_input(p, g["feat8"], "vision"), g["fuse_v8"]) < 1e-5
    assert rel_l2(OM.fusion_input(p, g["feat2"], "vision"), g["fuse_v2"]) < 1e-5      # frame table nearest-resized 8 -> 2
    assert rel_l2(OM.fusion_input(p, g["aud3"], "audio"), g["fuse_a3"]) < 1e-5
    assert rel_l2(OM.fusion_input(p, g["feat8"], "vision", pool_video=True), g["fuse_v8_pool"]) < 1e-5


# ---------------------------------------------------------------------------------------------- Kaldi fbank
def test_fbank_oracle_matches_torchaudio(golden_dir):
    from oracle import fbank as OF
    g = torch.load(os.path.join(golden_dir, "fbank.pt"), weights_only=False)
    for bins in (224, 64):
        fb = OF.kaldi_fbank(g["wave"] * 2 ** 15, bins)
        assert fb.shape == g[f"fbank_{bins}"].shape == (299, bins)
        assert (fb - g[f"fbank_{bins}"]).abs().max().item() < 2e-4      # log-mel values are O(10)
    out = OF.audio_processor(g["wave"], 224, 224, 3)
    assert out.shape == (3, 224, 224)


def test_loss_oracle_vs_reference_forward_ret_and_forward_cap():
    """oracle/mico.py ITC / ITM / caption losses against the fixture produced by the reference's OWN data/model/vast.py
    forward_ret / forward_cap (run unmodified over a stub module, oracle/make_golden.py:gen_losses): values and gradients,
    with the hard negatives and MLM masks the reference drew."""
    import torch.nn.functional as F
    from oracle import mico as OM
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "losses_tiny.pt"), weights_only=False)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in g["state_dict"].items()}
    p["multimodal_encoder.cls.predictions.decoder.weight"] = p["multimodal_encoder.bert.embeddings.word_embeddings.weight"]
    raw_t, raw_v, cond = (g[k].clone().requires_grad_(True) for k in ("raw_t", "raw_v", "cond"))
    ft, fv = F.normalize(raw_t, dim=-1), F.normalize(raw_v, dim=-1)
    l_itc, _, _ = OM.itc_loss(fv, ft, fv.detach(), ft.detach(), p["contra_temp"], 0)
    l_itm = OM.itm_loss(p, cond, cond, g["ids"], g["att"], g["ids"], g["att"], g["neg_c"], g["neg_t"], g["itm_ratio"],
                        g["layers"], g["heads"])
    l_cap = OM.caption_loss(p, cond, g["cap_ids"], g["att"], g["cap_labels"], g["layers"], g["heads"])
    for got, key in ((l_itc, "loss_itc"), (l_itm, "loss_itm"), (l_cap, "loss_cap")):
        assert abs(got.item() - g[key].item()) < 1e-5 * max(1.0, abs(g[key].item())), (key, got.item(), g[key].item())
    (l_itc + l_itm + l_cap).backward()
    assert rel_l2(raw_t.grad, g["d_raw_t"]) < 1e-4 and rel_l2(raw_v.grad, g["d_raw_v"]) < 1e-4 and rel_l2(cond.grad, g["d_cond"]) < 1e-4
    for k, want in g["grads"].items():
        assert rel_l2(p[k].grad, want) < 1e-4, k


def test_loss_oracle_two_rank_world_vs_reference():
    """The oracle's simulated 2-rank world (per-rank losses summed; gathers as concatenations) against the reference's
    forward_ret / forward_cap run on TWO gloo ranks (tests/golden/losses_2rank.pt): per-rank loss values, the gradient of every
    rank's fusion input (all_gather_with_grad: all-reduce(SUM) of the gathered gradient, own slice) and features, and the SUM
    over ranks of the reference's local parameter gradients (pipeline.py:93-99 sums without dividing)."""
    import torch.nn.functional as F
    from oracle import mico as OM
    gd = os.path.join(os.path.dirname(__file__), "golden")
    one = torch.load(os.path.join(gd, "losses_tiny.pt"), weights_only=False)
    ranks = torch.load(os.path.join(gd, "losses_2rank.pt"), weights_only=False)["ranks"]
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in one["state_dict"].items()}
    p["multimodal_encoder.cls.predictions.decoder.weight"] = p["multimodal_encoder.bert.embeddings.word_embeddings.weight"]
    leaves = [{k: r[k].clone().requires_grad_(True) for k in ("raw_t", "raw_v", "cond")} for r in ranks]
    ft = [F.normalize(l["raw_t"], dim=-1) for l in leaves]
    fv = [F.normalize(l["raw_v"], dim=-1) for l in leaves]
    ft_all, fv_all = torch.cat(ft).detach(), torch.cat(fv).detach()
    cond_all = torch.cat([l["cond"] for l in leaves])
    ids_all, att_all = torch.cat([r["ids"] for r in ranks]), torch.cat([r["att"] for r in ranks])
    total = 0.0
    for k, r in enumerate(ranks):
        l_itc, _, _ = OM.itc_loss(fv[k], ft[k], fv_all, ft_all, p["contra_temp"], k)
        l_itm = OM.itm_loss(p, leaves[k]["cond"], cond_all, r["ids"], r["att"], ids_all, att_all, r["neg_c"], r["neg_t"],
                            r["itm_ratio"], r["layers"], r["heads"])
        l_cap = OM.caption_loss(p, leaves[k]["cond"], r["cap_ids"], r["att"], r["cap_labels"], r["layers"], r["heads"])
        for got, key in ((l_itc, "loss_itc"), (l_itm, "loss_itm"), (l_cap, "loss_cap")):
            assert abs(got.item() - r[key].item()) < 1e-5 * max(1.0, abs(r[key].item())), (k, key, got.item(), r[key].item())
        total = total + l_itc + l_itm + l_cap
    total.backward()
    for k, r in enumerate(ranks):
        assert rel_l2(leaves[k]["cond"].grad, r["d_cond"]) < 1e-4, k
        assert rel_l2(leaves[k]["raw_t"].grad, r["d_raw_t"]) < 1e-4 and rel_l2(leaves[k]["raw_v"].grad, r["d_raw_v"]) < 1e-4
    for key in ranks[0]["grads"]:
        want = ranks[0]["grads"][key] + ranks[1]["grads"][key]
        assert rel_l2(p[key].grad, want) < 1e-4, key


def test_eva02_oracle_matches_reference(golden_dir):
    """oracle/eva02.py (RoPE on the patch tokens, separate q/k/v projections, sub-LN, SwiGLU) against the reference
    EVAVisionTransformer in its EVA02 configuration: forward and gradients.  Parity target for the EVA02 CUDA tower."""
    from oracle import eva02 as O2
    g = torch.load(os.path.join(golden_dir, "eva02_tiny.pt"), weights_only=False)
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in g["state_dict"].items()}
    y = O2.forward_features(p, g["x"], g["cfg"])
    assert y.shape == g["y"].shape == (2, 257, 128)
    assert rel_l2(y, g["y"]) < 1e-5
    y.pow(2).mean().backward()
    for k, want in g["grads"].items():
        assert rel_l2(p[k].grad, want) < 1e-4, k
