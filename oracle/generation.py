"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the caption decode loop.

The reference calls HF `GenerationMixin.generate` (transformers==4.31.0, set_env.sh:12; not vendored under /root/reference and
not runnable under the installed transformers 5.x) with its own hooks: model/bert.py:1110-1117
(update_attention_mask), :1126-1143 (prepare_inputs_for_generation: append [MASK], grow the 3-D mask, full forward, no KV
cache), :1145-1190 (_update_model_kwargs_for_generation).  Call sites: inference_demo.py:164-171 and
data/model/vast.py:535-545 (num_beams = beam_size, length_penalty 0.6, eos = [SEP]).  PARITY UNPINNED for the search
procedure: the published beam-search algorithm of 4.31 (generation/utils.py beam_search + generation/beam_search.py
BeamSearchScorer.process / finalize, BeamHypotheses.add / is_done with early_stopping False) is restated here from its
documentation; the per-step model call is the pinned oracle BERT (oracle/bert.py).
"""
import torch

from . import bert as OB


def grow_mask(m):
    """bert.py:1110-1117."""
    b, n, _ = m.shape
    out = torch.zeros(b, n + 1, n + 1, dtype=m.dtype)
    out[:, :n, :n] = m
    out[:, n, :n] = m[:, n - 1, :n]
    out[:, n, n] = 1
    return out


def mask_logits(p, ids, mask, enc, mask_token_id, **cfg):
    """prepare_inputs_for_generation + forward: logits at the appended [MASK] position (fp32)."""
    ids2 = torch.cat([ids, torch.full((ids.shape[0], 1), mask_token_id, dtype=torch.long)], 1)
    _, logits, _ = OB.masked_lm(p, ids2, grow_mask(mask), enc, None, **cfg)
    return logits[:, -1, :]


def beam_search(logits_of, ids, mask, enc, max_new_tokens, num_beams, eos, pad, length_penalty):
    """logits_of(ids, mask, enc) -> (rows, V).  Returns the best hypothesis per sample, HF output layout."""
    B, L0 = ids.shape
    nb = num_beams
    max_length = L0 + max_new_tokens
    ids = ids.repeat_interleave(nb, 0)
    mask = mask.repeat_interleave(nb, 0)
    enc = enc.repeat_interleave(nb, 0) if enc is not None else None
    running = torch.zeros(B, nb)
    running[:, 1:] = -1e9
    finished = [[] for _ in range(B)]        # (normalised score, tokens) kept best-nb
    worst = [1e9] * B
    done = [False] * B

    def add(b, toks, s):
        score = s / (len(toks) ** length_penalty)
        if len(finished[b]) < nb or score > worst[b]:
            finished[b].append((score, toks))
            if len(finished[b]) > nb:
                finished[b].sort(key=lambda t: t[0])
                finished[b].pop(0)
                worst[b] = finished[b][0][0]
            else:
                worst[b] = min(worst[b], score)

    cur = L0
    while True:
        lp = torch.log_softmax(logits_of(ids, mask, enc).float(), -1)
        V = lp.shape[-1]
        cand = (lp + running.view(-1, 1)).view(B, nb * V)
        vals, idx = cand.topk(2 * nb, dim=1)
        new_run = torch.zeros(B, nb)
        new_tok = torch.full((B, nb), pad, dtype=torch.long)
        new_src = torch.zeros(B, nb, dtype=torch.long)
        for b in range(B):
            if done[b]:
                continue
            k = 0
            for r in range(2 * nb):
                tok, src, s = int(idx[b, r]) % V, b * nb + int(idx[b, r]) // V, float(vals[b, r])
                if eos is not None and tok == eos:
                    if r < nb:
                        add(b, ids[src].clone(), s)
                    continue
                new_run[b, k], new_tok[b, k], new_src[b, k] = s, tok, src
                k += 1
                if k == nb:
                    break
            if len(finished[b]) >= nb and worst[b] >= float(vals[b].max()) / (cur + 1) ** length_penalty:
                done[b] = True
        running = new_run
        ids = torch.cat([ids[new_src.view(-1)], new_tok.view(-1, 1)], 1)
        mask = grow_mask(mask)
        cur += 1
        if all(done) or cur >= max_length:
            break
    best = []
    for b in range(B):
        if not done[b]:
            for j in range(nb):
                add(b, ids[b * nb + j], float(running[b, j]))
        best.append(max(finished[b], key=lambda t: t[0])[1])
    width = min(max(len(t) for t in best) + 1, max_length)
    out = torch.full((B, width), pad, dtype=torch.long)
    for b, t in enumerate(best):
        out[b, :len(t)] = t
        if len(t) < width and eos is not None:
            out[b, len(t)] = eos
    return out


def greedy(logits_of, ids, mask, enc, max_new_tokens, eos, pad):
    B, L0 = ids.shape
    alive = torch.ones(B, dtype=torch.long)
    while True:
        nxt = logits_of(ids, mask, enc).float().argmax(-1)
        if eos is not None:
            nxt = nxt * alive + pad * (1 - alive)
        ids = torch.cat([ids, nxt[:, None]], 1)
        mask = grow_mask(mask)
        if eos is not None:
            alive = alive * (nxt != eos).long()
        if int(alive.max()) == 0 or ids.shape[1] >= L0 + max_new_tokens:
            return ids
