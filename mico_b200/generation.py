"""Caption generation: the decode loop `inference_demo.py:161-174` / `data/model/vast.py:514-553` drive
through `multimodal_encoder.generate(...)`.

The reference inherits HF `GenerationMixin.generate` (transformers 4.31) and customises three hooks (model/bert.py):
  * prepare_inputs_for_generation (:1126-1143): append a [MASK] token and grow the 3-D attention mask, then run the FULL
    sequence through the encoder (no KV cache); the prediction is read at the [MASK] position;
  * update_attention_mask (:1110-1117): new row = copy of the last row + itself  (a causal mask grown one step);
  * _update_model_kwargs_for_generation (:1145-1190): the kwargs mask grows the same way after every step.
This module restates that loop for the three modes the reference uses: beam search (num_beams = config.beam_size,
length_penalty 0.6 -- HF BeamSearchScorer / BeamHypotheses semantics, early_stopping False), greedy (num_beams 1) and top-k
sampling (captioner_mode: do_sample, top_k 10).  The encoder forward and LM head are the CUDA kernels of mico_b200.bert;
the LM head runs on the [MASK] position only.  Search bookkeeping (a handful of candidates per sample) is host-side.
"""
import torch

from ._lib import MicoError


def update_attention_mask(attention_mask):
    """bert.py:1110-1117 on a (b, n, n) mask."""
    b, n, _ = attention_mask.shape
    upd = attention_mask.new_zeros(b, n + 1, n + 1)
    upd[:, :n, :n] = attention_mask
    upd[:, n, :n] = attention_mask[:, n - 1, :n]
    upd[:, n, n] = 1
    return upd


class BeamHypotheses:
    """HF generation/beam_search.py BeamHypotheses (4.31), early_stopping False."""

    def __init__(self, num_beams, length_penalty):
        self.length_penalty = length_penalty
        self.num_beams = num_beams
        self.beams = []
        self.worst_score = 1e9

    def __len__(self):
        return len(self.beams)

    def add(self, hyp, sum_logprobs):
        score = sum_logprobs / (hyp.shape[-1] ** self.length_penalty)
        if len(self) < self.num_beams or score > self.worst_score:
            self.beams.append((score, hyp))
            if len(self) > self.num_beams:
                sorted_next_scores = sorted([(s, idx) for idx, (s, _) in enumerate(self.beams)])
                del self.beams[sorted_next_scores[0][1]]
                self.worst_score = sorted_next_scores[1][0]
            else:
                self.worst_score = min(score, self.worst_score)

    def is_done(self, best_sum_logprobs, cur_len):
        if len(self) < self.num_beams:
            return False
        highest_attainable_score = best_sum_logprobs / cur_len ** self.length_penalty
        return self.worst_score >= highest_attainable_score


def update_position_ids(position_ids):
    """bert.py:1119-1124."""
    b, n = position_ids.shape
    upd = position_ids.new_zeros(b, n + 1)
    upd[:, :n] = position_ids
    upd[:, n] = upd[:, n - 1] + 1
    return upd


def prepare_inputs_for_generation(input_ids, attention_mask, mask_token_id, position_ids=None, encoder_hidden_states=None):
    """bert.py:1126-1143: append a [MASK] token, grow the 3-D mask (and the position ids when given).  Pinned to the
    reference's own function by tests/golden/generation_steps.pt (tests/test_generation.py)."""
    rows = input_ids.shape[0]
    dummy = torch.full((rows, 1), mask_token_id, dtype=torch.long, device=input_ids.device)
    return dict(input_ids=torch.cat([input_ids, dummy], dim=1), attention_mask=update_attention_mask(attention_mask),
                position_ids=update_position_ids(position_ids) if position_ids is not None else None,
                encoder_hidden_states=encoder_hidden_states)


def step_logits(model, input_ids, attention_mask, encoder_hidden_states, mask_token_id):
    """One decode step of prepare_inputs_for_generation + forward: fp32 logits (rows, vocab) at the appended [MASK]."""
    x = prepare_inputs_for_generation(input_ids, attention_mask, mask_token_id)
    return model.mask_position_logits(x["input_ids"], x["attention_mask"], encoder_hidden_states)


@torch.no_grad()
def generate(model, input_ids, attention_mask, encoder_hidden_states=None, max_new_tokens=20, num_beams=1,
             eos_token_id=None, pad_token_id=None, length_penalty=1.0, do_sample=False, top_k=None, mask_token_id=None,
             generator=None, logits_fn=None):
    """Returns (batch, <= 1 + max_new_tokens) token ids: the prompt followed by the generated tokens, finished rows padded
    with pad_token_id (HF `generate` output convention; callers drop column 0, vast.py:548).

    logits_fn(input_ids, attention_mask) -> (rows, vocab) fp32 overrides the model call (tests drive the same search over
    the CPU oracle with it)."""
    if mask_token_id is None:
        tok = getattr(model, "tokenizer", None)
        mask_token_id = getattr(tok, "mask_token_id", None)
        if mask_token_id is None:
            raise MicoError("generate: mask_token_id is required (the reference reads self.tokenizer.mask_token_id, bert.py:1135)")
    if pad_token_id is None:
        pad_token_id = eos_token_id if eos_token_id is not None else 0
    if attention_mask.dim() != 3:
        raise MicoError("generate: the reference passes a 3-D (b, n, n) attention mask (vast.py:524)")
    dev = input_ids.device
    batch, prompt_len = input_ids.shape
    max_length = prompt_len + max_new_tokens
    nb = 1 if do_sample else int(num_beams)

    def logits_of(ids, mask, enc):
        if logits_fn is not None:
            return logits_fn(ids, mask, enc)
        return step_logits(model, ids, mask, enc, mask_token_id)

    if nb == 1:
        # greedy_search / sample (HF 4.31): finished rows keep emitting pad_token_id
        unfinished = torch.ones(batch, dtype=torch.long, device=dev)
        while True:
            logits = logits_of(input_ids, attention_mask, encoder_hidden_states).float()
            if do_sample:
                if top_k:
                    kth = torch.topk(logits, min(int(top_k), logits.shape[-1]))[0][..., -1, None]
                    logits = logits.masked_fill(logits < kth, -float("inf"))      # TopKLogitsWarper
                probs = torch.softmax(logits, dim=-1)
                nxt = torch.multinomial(probs, 1, generator=generator).squeeze(1)
            else:
                nxt = torch.argmax(logits, dim=-1)
            if eos_token_id is not None:
                nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
            input_ids = torch.cat([input_ids, nxt[:, None]], dim=-1)
            attention_mask = update_attention_mask(attention_mask)
            if eos_token_id is not None:
                unfinished = unfinished * (nxt != eos_token_id).long()
            if int(unfinished.max()) == 0 or input_ids.shape[1] >= max_length:
                return input_ids

    # ---- beam search (HF 4.31 beam_search + BeamSearchScorer, num_beam_hyps_to_keep 1)
    input_ids = input_ids.repeat_interleave(nb, dim=0)
    attention_mask = attention_mask.repeat_interleave(nb, dim=0)
    enc = encoder_hidden_states.repeat_interleave(nb, dim=0) if encoder_hidden_states is not None else None
    beam_scores = torch.zeros((batch, nb), dtype=torch.float32, device=dev)
    beam_scores[:, 1:] = -1e9
    beam_scores = beam_scores.view(-1)
    hyps = [BeamHypotheses(nb, length_penalty) for _ in range(batch)]
    done = [False] * batch
    cur_len = prompt_len
    while True:
        logits = logits_of(input_ids, attention_mask, enc).float()
        vocab = logits.shape[-1]
        scores = torch.log_softmax(logits, dim=-1) + beam_scores[:, None]
        top_scores, top_idx = torch.topk(scores.view(batch, nb * vocab), 2 * nb, dim=1, largest=True, sorted=True)
        top_scores_h, top_idx_h = top_scores.cpu(), top_idx.cpu()
        ids_h = input_ids.cpu()
        next_scores = torch.zeros((batch, nb), dtype=torch.float32)
        next_tokens = torch.zeros((batch, nb), dtype=torch.long)
        next_indices = torch.zeros((batch, nb), dtype=torch.long)
        for b in range(batch):
            if done[b]:
                next_tokens[b, :] = pad_token_id       # scores 0, index 0: padded beams of a finished sample
                continue
            k = 0
            for rank in range(2 * nb):
                tok = int(top_idx_h[b, rank]) % vocab
                src = b * nb + int(top_idx_h[b, rank]) // vocab
                sc = float(top_scores_h[b, rank])
                if eos_token_id is not None and tok == eos_token_id:
                    if rank >= nb:
                        continue
                    hyps[b].add(ids_h[src].clone(), sc)
                else:
                    next_scores[b, k], next_tokens[b, k], next_indices[b, k] = sc, tok, src
                    k += 1
                if k == nb:
                    break
            done[b] = done[b] or hyps[b].is_done(float(top_scores_h[b].max()), cur_len + 1)
        beam_scores = next_scores.view(-1).to(dev)
        sel = next_indices.view(-1).to(dev)
        input_ids = torch.cat([input_ids[sel], next_tokens.view(-1, 1).to(dev)], dim=-1)
        attention_mask = update_attention_mask(attention_mask)
        cur_len += 1
        if all(done) or cur_len >= max_length:
            break
    # finalize: open beams become hypotheses; best one per sample
    ids_h, scores_h = input_ids.cpu(), beam_scores.cpu()
    best = []
    for b in range(batch):
        if not done[b]:
            for j in range(nb):
                hyps[b].add(ids_h[b * nb + j], float(scores_h[b * nb + j]))
        best.append(sorted(hyps[b].beams, key=lambda x: x[0])[-1][1])
    sent_max = min(max(len(h) for h in best) + 1, max_length)
    out = torch.full((batch, sent_max), pad_token_id, dtype=torch.long)
    for b, h in enumerate(best):
        out[b, :len(h)] = h
        if len(h) < sent_max and eos_token_id is not None:
            out[b, len(h)] = eos_token_id
    return out.to(dev)
