"""The omni-modal training step scheduled for one B200 (BASELINE configs[4]).

``MiCo.forward(batch, task, compute_loss=True)`` states the step the way the reference does (data/model/vast.py:317-348 ->
forward_ret :383-464 + forward_cap :485-512): lazily, one modality / sub-task at a time.  Run that way at the BASELINE
shape -- 64 samples x (8 video frames + 3 spectrogram slices + 1 depth map) + 128 text tokens, seven ITC / ITM / caption
sub-tasks -- the fusion encoder alone keeps ~110 GB of activations alive until ``loss.backward()`` (every sub-task's
cross-attention K / V over up to 2827 visual tokens, 12 layers), on top of the tower's checkpoints.  This module computes
the SAME losses and gradients in an order that fits the machine:

  1. ONE tower pass over every frame of every modality (``EVAVisionTransformer.forward_multi``): one weight read, one
     backward pass, one contiguous gradient buffer whose block buckets a data-parallel caller can reduce while the
     backward runs (mico_b200/dp.py).
  2. ITC on the pooled features (tiny; ordinary autograd).
  3. Per modality combination, the ITM and caption sub-tasks that share its fusion input run as one *eager loss group*
     (``_EagerLosses``): the group's sub-graph is built, differentiated and freed inside the forward pass; the node keeps only
     the gradients w.r.t. the fusion input and the parameters.  When the caller's ``backward()`` reaches the node those
     gradients are scaled by the upstream scalar (1, or GradScaler's loss scale) -- exact by linearity.  Peak memory is the
     largest single group instead of the sum of all sub-tasks.

Nothing here changes a value: DropPath / dropout draws are per-sample either way, the negatives are sampled from the same
distributions, and eval / injected-choice runs are bit-comparable with the reference-ordered path
(tests/test_gpu_mico.py::test_fused_step_matches_reference_order).
"""
import torch

from . import functional as MF
from . import ops
from .ops import BF16, F32, MicoError

_COMBO_PARTS = {"v": "v", "a": "a", "d": "d", "s": "s", "va": "va", "vs": "vs", "vas": "vas", "id": "vd"}


def _scale_grads_(grads, g):
    """grads[i] *= g (0-dim device tensor), in place, with as few launches as the storage layout allows: slices of one
    arena that are adjacent in memory are scaled by one launch."""
    g1 = g.detach().reshape(1).to(F32)
    items = sorted(((t.data_ptr(), t) for t in grads if t is not None and t.numel() > 0), key=lambda x: x[0])
    runs = []
    for ptr, t in items:
        if t.dtype != F32 or not t.is_contiguous() or ptr % 16:
            t.mul_(g1.to(t.dtype))           # odd layouts (none on the MiCo path): plain torch
            continue
        nbytes = t.numel() * 4
        if runs and runs[-1][2] is not None and t.untyped_storage().data_ptr() == runs[-1][2] \
                and 0 <= ptr - (runs[-1][0] + runs[-1][1]) < 16:
            runs[-1][1] = ptr + nbytes - runs[-1][0]
        else:
            runs.append([ptr, nbytes, t.untyped_storage().data_ptr(), t])
    for ptr, nbytes, _, t in runs:
        # a run is addressed through its first tensor's storage
        base = t.untyped_storage()
        off = (ptr - base.data_ptr()) // 4
        flat = torch.empty(0, device=t.device, dtype=F32).set_(base, off, (nbytes // 4,), (1,))
        ops.scale_(flat, scale_dev=g1)


class _EagerLosses(torch.autograd.Function):
    """losses = fn(*leaves): built, differentiated and freed inside forward (see the module docstring).

    inputs = n_leaf activation tensors followed by the parameters `fn` uses; outputs = the scalar losses of the group.
    backward: every stored gradient times the upstream scalar.  The group was differentiated as ONE sum, so all of its
    outputs must receive the SAME upstream gradient (they do under ``sum(losses.values()).backward()``, scaled or not,
    pipeline.py:44-47,86-88); unequal weights poison the gradients with NaN instead of silently mis-weighting them."""

    @staticmethod
    def forward(ctx, fn, n_leaf, *inputs):
        leaves_in, params = inputs[:n_leaf], inputs[n_leaf:]
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(t.requires_grad) for t in leaves_in]
            losses = fn(*leaves)
            if not isinstance(losses, (tuple, list)):
                losses = (losses,)
            total = losses[0]
            for l in losses[1:]:
                total = total + l
            wrt = [l for l in leaves if l.requires_grad] + [p for p in params if p.requires_grad]
            grads = torch.autograd.grad(total, wrt, allow_unused=True) if wrt else ()
        it = iter(grads)
        ctx.grads = [next(it) if l.requires_grad else None for l in leaves] + \
                    [next(it) if p.requires_grad else None for p in params]
        ctx.n_out = len(losses)
        out = tuple(l.detach().clone() for l in losses)
        return out if len(out) > 1 else out[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *gs):
        g = gs[0]
        for h in gs[1:]:
            g = torch.where(g == h, g, torch.full_like(g, float("nan")))
        grads, ctx.grads = ctx.grads, None
        if grads is None:
            raise MicoError("an eager loss group can be differentiated once")
        _scale_grads_(grads, g)
        return (None, None) + tuple(grads)


def eager_losses(fn, leaves, params):
    """Run `fn(*leaves) -> tuple of scalar losses` as an eager loss group; `params`: every Parameter `fn` touches."""
    return _EagerLosses.apply(fn, len(leaves), *leaves, *params)


def _group_params(model):
    ps, seen = [], set()
    for mod in (model.multimodal_encoder, model.itm_head):
        for p in mod.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                ps.append(p)
    return ps


def prefetch_towers(model, batch, combos):
    """Every tower input the sub-tasks need, through ONE pass (step 1 of the module docstring)."""
    need = set("".join(_COMBO_PARTS[c] for c in combos)) & set("vad")
    tower = getattr(model.vision_encoder, "visual", None)
    jobs = []
    for m, key, src in (("v", "vision_output", "vision_pixels"), ("a", "audio_output", "audio_spectrograms"),
                        ("d", "depth_output", "depth_pixels")):
        if m in need and key not in batch and src in batch:      # (features may be supplied precomputed)
            jobs.append((key, batch[src]))
    if len(jobs) < 2 or not hasattr(tower, "forward_multi"):
        return               # a single modality (or a tower without the multi-input pass): the lazy path does the same work
    xs = []
    for key, x in jobs:
        if key == "audio_output":        # (b, n, T, mel): one plane per slice, replicated to 3 channels inside the im2col
            xs.append(x.reshape(-1, *x.shape[-2:]))
        else:
            xs.append(x.reshape(-1, *x.shape[-3:]))
    outs = tower.forward_multi(xs)
    for (key, x), y in zip(jobs, outs):
        batch[key] = y.reshape(x.shape[0], x.shape[1], *y.shape[-2:])


def fused_train_forward(model, batch, task):
    """Losses of ``MiCo.forward(batch, task, compute_loss=True)`` on the schedule above.  Returns the same dict."""
    from .mico import _rank, _world, all_gather_with_grad, concat_all_gather
    ret_st, cap_st = [], []
    for t in task.split("_"):
        subs = t.split("%")[1:]
        if t.startswith("ret"):
            ret_st += subs
        elif t.startswith("cap"):
            cap_st += subs
        else:
            raise NotImplementedError(t)
    for st in ret_st + cap_st:
        assert st in model._SUBTASKS, st
    combos = []
    for st in ret_st + cap_st:
        if st[1:] not in combos:
            combos.append(st[1:])

    tokens = model.batch_get(batch, "caption_tokens")
    dev = model.contra_temp.device
    # MLM masking happens on the host (general_module.py:52-97 runs a python loop and ends with .cuda()): do it before any
    # kernel is queued so that reading the ids back never waits for the GPU
    cap_ids = cap_labels = None
    if cap_st:
        if "cap_input_ids" in batch:
            cap_ids, cap_labels = batch["cap_input_ids"], batch["cap_labels"]
        else:
            cap_ids, cap_labels = model.text_masker(tokens.input_ids, 0.6)
        cap_ids, cap_labels = cap_ids.to(dev, non_blocking=True), cap_labels.to(dev, non_blocking=True)
    input_ids = tokens.input_ids.to(dev, non_blocking=True)
    attention_mask = tokens.attention_mask.to(dev, non_blocking=True)
    batch["caption_tokens"] = type(tokens)(input_ids=input_ids, attention_mask=attention_mask)

    prefetch_towers(model, batch, combos)

    out = {}
    rank = _rank()
    sims = {}
    if ret_st:
        feat_t = model.batch_get(batch, "feat_t")
        bs = feat_t.shape[0]
        feat_t_all = concat_all_gather(feat_t)
        ids_all = concat_all_gather(input_ids)
        att_all = concat_all_gather(attention_mask)
        targets = torch.arange(rank * bs, rank * bs + bs, device=dev)
        loss_itc = []
        for st in ret_st:       # ---- ITC (vast.py:402-417)
            feat_c = model.batch_get(batch, f"feat_{st[1:]}")
            feat_c_all = concat_all_gather(feat_c)
            sim_c2t = MF.contrastive_logits(feat_c, feat_t_all, model.contra_temp)
            sim_t2c = MF.contrastive_logits(feat_t, feat_c_all, model.contra_temp)
            loss_itc.append((MF.cross_entropy(sim_c2t, targets, label_smoothing=0.1)
                             + MF.cross_entropy(sim_t2c, targets, label_smoothing=0.1)) / 2)
            sims[st[1:]] = (sim_c2t.detach(), sim_t2c.detach(), st)
        out["loss_itc"] = sum(loss_itc) / len(loss_itc)

    S = attention_mask.shape[1]
    att3 = None
    if cap_st:
        att3 = torch.tril(attention_mask.unsqueeze(1).expand(-1, S, -1).clone())      # vast.py:497-499
    params = _group_params(model)
    n_ret, n_cap = len(ret_st), len(cap_st)
    ret_combos = [st[1:] for st in ret_st]
    cap_combos = [st[1:] for st in cap_st]
    itm_terms, cap_terms = [], []
    for c in combos:
        cond = model.batch_get(batch, f"condition_feats_{c}")
        do_itm, do_cap = c in ret_combos, c in cap_combos
        neg = None
        if do_itm:       # hard negatives from the ITC similarities (vast.py:423-440), one batched draw per direction
            sim_c2t, sim_t2c, st = sims[c]
            bs = cond.shape[0]
            with torch.no_grad():
                w_t2c = torch.softmax(sim_t2c, dim=1) + 1e-4
                w_t2c[:, rank * bs:rank * bs + bs].fill_diagonal_(0)
                w_c2t = torch.softmax(sim_c2t, dim=1) + 1e-4
                w_c2t[:, rank * bs:rank * bs + bs].fill_diagonal_(0)
                neg_c = batch.get(f"itm_neg_cond_{st}")
                neg_t = batch.get(f"itm_neg_text_{st}")
                neg_c = torch.multinomial(w_t2c, 1).view(-1) if neg_c is None else neg_c.to(dev)
                neg_t = torch.multinomial(w_c2t, 1).view(-1) if neg_t is None else neg_t.to(dev)
            neg = (neg_c, neg_t)

        def group(cond_leaf, do_itm=do_itm, do_cap=do_cap, neg=neg):
            """ONE fusion-encoder call for every text sequence that reads this combination's visual tokens: ITM positive,
            ITM negative-condition, ITM negative-text (vast.py:445-451) and the masked caption (vast.py:497-507).  The
            sequences of sample i all cross-attend to encoder entry i (the negative-condition ones to the sampled entry):
            `encoder_index` shares each entry's K / V projection between them."""
            bs = cond_leaf.shape[0]
            ar = torch.arange(bs, device=dev)
            ids, masks, index = [], [], []
            enc = [cond_leaf]
            if do_itm:
                ids += [input_ids, input_ids, ids_all[neg[1]]]
                att_itm = torch.cat((attention_mask, attention_mask, att_all[neg[1]]), dim=0)
                if _world() == 1:        # the negatives are local samples: read their entry in place
                    index += [ar, neg[0], ar]
                else:                    # negatives gathered from every rank (vast.py:422), with gradient
                    cond_all = all_gather_with_grad(cond_leaf)
                    enc.append(cond_all[neg[0]])
                    index += [ar, bs + ar, ar]
                masks.append(att_itm)
            if do_cap:
                ids.append(cap_ids)
                masks.append(att3)
                index.append(ar)
            h = model.multimodal_encoder.bert(input_ids=torch.cat(ids, dim=0),
                                              attention_mask=masks if len(masks) > 1 else masks[0],
                                              encoder_hidden_states=enc[0] if len(enc) == 1 else torch.cat(enc, dim=0),
                                              encoder_index=torch.cat(index)).last_hidden_state
            losses = []
            if do_itm:       # ---- ITM head + CE (vast.py:453-457)
                logits = model.itm_head(h[:3 * bs, 0])
                truth = torch.zeros(bs * 3, dtype=torch.long, device=dev)
                truth[:bs] = 1
                losses.append(model.itm_ratio * MF.cross_entropy(logits, truth) / n_ret)
            if do_cap:       # ---- LM head + CE on the caption sequences (vast.py:504-510)
                hc = h[3 * bs:] if do_itm else h
                losses.append(model.multimodal_encoder.lm_loss(hc, cap_labels) / n_cap)
            return tuple(losses)

        res = eager_losses(group, [cond], params)
        res = res if isinstance(res, tuple) else (res,)
        k = 0
        if do_itm:
            itm_terms.append(res[k]); k += 1
        if do_cap:
            cap_terms.append(res[k])
    if itm_terms:
        out["loss_itm"] = sum(itm_terms)
    if cap_terms:
        out["loss_cap"] = sum(cap_terms)
    return out
