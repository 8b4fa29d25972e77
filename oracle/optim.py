"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference optimizer step and LR schedule.

Follows data/utils/build_optimizer.py:136-196 (AdamW.step), :13-62 (parameter grouping) and data/utils/sched.py:3-31.
Pinned by tests/golden/adamw.pt, produced by oracle/make_golden.py from the reference's own AdamW class and
get_lr_sched.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this package.
"""
import math


def adamw_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
    """One AdamW update of build_optimizer.py:158-194 on clones; returns (p, m, v)."""
    p, m, v = p.clone(), m.clone(), v.clone()
    beta1, beta2 = betas
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)                         # :170
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)                  # :171
    denom = v.sqrt().add_(eps)                                       # :172
    step_size = lr
    if correct_bias:                                                 # :175-179
        step_size = step_size * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)
    p.addcdiv_(m, denom, value=-step_size)                           # :181
    if weight_decay > 0.0:                                           # :193-194 (on the already-updated parameter)
        p.add_(p, alpha=-lr * weight_decay)
    return p, m, v


def warmup_linear(x, warmup_ratio):                                  # sched.py:16-21
    if x < warmup_ratio:
        return x / warmup_ratio
    return max((x - 1.) / (warmup_ratio - 1.), 0)


def warmup_cosine(x, warmup_ratio):                                  # sched.py:3-6
    if x < warmup_ratio:
        return x / warmup_ratio
    return 0.5 * (1.0 + math.cos(math.pi * x))


def lr_ratio(global_step, num_train_steps, warmup_ratio, scheduler="warmup_linear"):   # sched.py:27-31
    f = dict(warmup_linear=warmup_linear, warmup_cosine=warmup_cosine)[scheduler]
    return f(global_step / num_train_steps, warmup_ratio)


def group_of(name, new_params_name=(), vision_clip=True):
    """Index of the reference parameter group (build_optimizer.py:32-62) a parameter name falls into:
    0 basic, 1 basic no-decay, 2 new, 3 new no-decay, 4 clip visual, 5 clip visual no-decay."""
    no_decay = ['bias', 'LayerNorm.bias', 'LayerNorm.weight']
    nd = any(s in name for s in no_decay)
    if any(s in name for s in new_params_name):
        return 3 if nd else 2
    if vision_clip and 'visual' in name:
        return 5 if nd else 4
    return 1 if nd else 0
