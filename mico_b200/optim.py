"""Optimizer + LR schedule of the reference training loop, with the update fused into ONE CUDA launch.

Mirrors data/utils/build_optimizer.py (parameter grouping :13-62, the AdamW class :105-196) and data/utils/sched.py;
`AdamW.step()` hands every parameter of every group to `mico_adamw_multi` (csrc/optim.cu) instead of looping over
tensors with ~8 torch kernels each.  State layout (`state[p] = {step, exp_avg, exp_avg_sq}`) and `param_groups` keys
(`lr`, `init_lr`, `betas`, `eps`, `weight_decay`, `correct_bias`) are the reference's, so `state_dict()` /
`load_state_dict()` interoperate with checkpoints written by the reference (`ModelSaver`, data/utils/save.py).
"""
import ctypes as C
import math

import numpy as np
import torch
from torch.optim import Adam, Adamax, Optimizer

from ._lib import MicoError, check, lib

_CHUNK = 16384          # fp32 elements per work item (64 KB of each of p, g, m, v)
_MAX_GROUPS = 16

_TENSOR_DT = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("pb", "<u8"), ("n", "<i8"),
                       ("group", "<i4"), ("reserved", "<i4")])
assert _TENSOR_DT.itemsize == 56


class _Hyper(C.Structure):
    _fields_ = [("lr", C.c_float), ("step_size", C.c_float), ("weight_decay", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float)]


# ----------------------------------------------------------------------------- data/utils/sched.py
def warmup_cosine(x, warmup_ratio):
    if x < warmup_ratio:
        return x / warmup_ratio
    return 0.5 * (1.0 + math.cos(math.pi * x))


def warmup_constant(x, warmup_ratio):
    if x < warmup_ratio:
        return x / warmup_ratio
    return 1.0


def warmup_linear(x, warmup_ratio):
    if x < warmup_ratio:
        return x / warmup_ratio
    return max((x - 1.) / (warmup_ratio - 1.), 0)


scheduler_dict = {'warmup_linear': warmup_linear, 'warmup_cosine': warmup_cosine}


def get_lr_sched(global_step, opts):
    """sched.py:27-31: ratio applied to every group's `init_lr` each step (pipeline.py:75-78)."""
    current_ratio = global_step / opts.num_train_steps
    return scheduler_dict[opts.scheduler](current_ratio, opts.warmup_ratio)


def apply_lr_sched(optimizer, global_step, opts):
    """pipeline.py:75-78."""
    lr_ratio = get_lr_sched(global_step, opts)
    for param_group in optimizer.param_groups:
        param_group['lr'] = param_group['init_lr'] * lr_ratio
    return lr_ratio


# ----------------------------------------------------------------------------- AdamW
class AdamW(Optimizer):
    """Adam with the decoupled weight-decay fix (build_optimizer.py:105-196), one fused launch per step.

    Same constructor and defaults as the reference (eps 1e-6, correct_bias True).  `grad_scale` multiplies every gradient
    inside the kernel (e.g. 1/world_size after a summed all-reduce; the reference loop sums, pipeline.py:93-99, so 1)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        if lr < 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super(AdamW, self).__init__(params, defaults)
        self.grad_scale = 1.0
        self._sink_owners = []      # modules exposing bf16_weight_sinks() / mark_weights_fresh()
        self._chunk_cache = {}      # (device, sizes) -> (chunk_tensor, chunk_index) device int32 arrays

    def attach_bf16_sinks(self, module):
        """The step also writes the bf16 GEMM-operand copies that `module` (a mico_b200 tower) caches, instead of the
        tower re-casting every weight on its next forward."""
        self._sink_owners.append(module)

    def _chunks(self, device, sizes):
        key = (device, tuple(sizes))
        hit = self._chunk_cache.get(key)
        if hit is None:
            ct, ci = [], []
            for t, n in enumerate(sizes):
                k = (n + _CHUNK - 1) // _CHUNK
                ct.append(np.full(k, t, np.int32))
                ci.append(np.arange(k, dtype=np.int32))
            ct, ci = np.concatenate(ct), np.concatenate(ci)
            hit = (torch.from_numpy(ct).to(device), torch.from_numpy(ci).to(device), len(ct))
            if len(self._chunk_cache) >= 8:      # the parameter set is stable from step to step: a handful of layouts
                self._chunk_cache.clear()
            self._chunk_cache[key] = hit
        return hit

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        sinks = {}
        for owner in self._sink_owners:
            for p, w, _ in owner.bf16_weight_sinks():
                sinks[id(p)] = w
        per_dev = {}
        hyper, hyper_idx = [], {}
        for gi, group in enumerate(self.param_groups):
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                grad = p.grad
                if grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                if not p.is_cuda:
                    raise MicoError("mico_b200.optim.AdamW runs on CUDA parameters only (no CPU fallback)")
                if p.dtype != torch.float32 or grad.dtype != torch.float32 or not p.is_contiguous():
                    raise MicoError("mico_b200.optim.AdamW expects contiguous fp32 parameters and gradients")
                if not grad.is_contiguous():
                    grad = grad.contiguous()
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                state['step'] += 1
                t = int(state['step'])
                step_size = group['lr']
                if group['correct_bias']:
                    step_size = step_size * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
                hk = (gi, t)
                if hk not in hyper_idx:
                    hyper_idx[hk] = len(hyper)
                    hyper.append((group['lr'], step_size, group['weight_decay'], beta1, beta2, group['eps']))
                w = sinks.get(id(p))
                per_dev.setdefault(p.device, []).append(
                    (p.data_ptr(), grad.data_ptr(), state['exp_avg'].data_ptr(), state['exp_avg_sq'].data_ptr(),
                     w.data_ptr() if w is not None else 0, p.numel(), hyper_idx[hk], 0, grad, p))
        if not per_dev:
            return loss
        updated = set()
        for dev, all_rows in per_dev.items():
            # the kernel takes a table of _MAX_GROUPS hyper-parameter rows per launch.  Parameters that skipped steps
            # (grad None on some batches: the audio / depth heads on image-only batches) keep older step counts, so one
            # step can need more (group, t) rows than that: launch once per slice of _MAX_GROUPS rows -- the reference's
            # per-parameter loop (build_optimizer.py:136-196) has no such limit either.
            for lo in range(0, len(hyper), _MAX_GROUPS):
                rows = [r for r in all_rows if lo <= r[6] < lo + _MAX_GROUPS]
                if not rows:
                    continue
                htab = (_Hyper * _MAX_GROUPS)()
                n_h = min(_MAX_GROUPS, len(hyper) - lo)
                for i in range(n_h):
                    htab[i] = _Hyper(*hyper[lo + i])
                tab = np.empty(len(rows), _TENSOR_DT)
                for i, r in enumerate(rows):
                    tab[i] = r[:6] + (r[6] - lo, 0)
                sizes = [r[5] for r in rows]
                ct, ci, n_chunks = self._chunks(dev, sizes)
                tab_dev = torch.from_numpy(tab.view(np.uint8)).pin_memory().to(dev, non_blocking=True)
                with torch.cuda.device(dev):
                    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                    check(lib.mico_adamw_multi(C.c_void_p(tab_dev.data_ptr()), C.c_void_p(ct.data_ptr()),
                                               C.c_void_p(ci.data_ptr()), int(n_chunks), int(_CHUNK), htab, n_h,
                                               C.c_float(self.grad_scale), C.c_double(float(sum(sizes))), stream),
                          "mico_adamw_multi")
                for r in rows:       # the raw-pointer update is invisible to autograd's version counters
                    torch.autograd.graph.increment_version(r[9])
                    if r[4]:
                        updated.add(id(r[9]))
        for owner in self._sink_owners:
            owner.mark_weights_fresh(updated)
        return loss


# ----------------------------------------------------------------------------- build_optimizer.py:13-103
def build_optimizer(model, args, checkpoint_optim=None):
    """Same six parameter groups as the reference: {basic, new, clip-visual} x {decay, no decay}; names containing
    'bias' / 'LayerNorm.bias' / 'LayerNorm.weight' get no weight decay; `args.run_cfg.new_params_name` selects the
    new-lr groups; 'visual' parameters get `clip_lr` when the vision encoder is a CLIP."""
    vision_clip = 'vision_encoder_type' in args.model_cfg and 'clip' in args.model_cfg.vision_encoder_type
    no_decay = ['bias', 'LayerNorm.bias', 'LayerNorm.weight']
    basic_params, basic_params_no_decay = [], []
    clip_params_visual, clip_params_no_decay_visual = [], []
    new_params, new_params_no_decay, new_params_name = [], [], []
    for k, v in model.named_parameters():
        is_new = any(nd in k for nd in args.run_cfg.new_params_name)
        nd_hit = any(nd in k for nd in no_decay)
        if is_new and not nd_hit:
            new_params.append(v)
            new_params_name.append(k)
        elif is_new and nd_hit:
            new_params_no_decay.append(v)
            new_params_name.append(k)
        elif vision_clip and 'visual' in k and not nd_hit:
            clip_params_visual.append(v)
        elif vision_clip and 'visual' in k and nd_hit:
            clip_params_no_decay_visual.append(v)
        elif not nd_hit:
            basic_params.append(v)
        else:
            basic_params_no_decay.append(v)
    rc = args.run_cfg
    optimizer_grouped_parameters = [
        {'params': basic_params, 'weight_decay': rc.weight_decay, 'lr': rc.learning_rate},
        {'params': basic_params_no_decay, 'weight_decay': 0.0, 'lr': rc.learning_rate},
        {'params': new_params, 'weight_decay': rc.weight_decay, 'lr': rc.new_lr},
        {'params': new_params_no_decay, 'weight_decay': 0.0, 'lr': rc.new_lr},
        {'params': clip_params_visual, 'weight_decay': rc.weight_decay, 'lr': rc.clip_lr},
        {'params': clip_params_no_decay_visual, 'weight_decay': 0.0, 'lr': rc.clip_lr},
    ]
    if rc.optim == 'adam':
        OptimCls = Adam
    elif rc.optim == 'adamax':
        OptimCls = Adamax
    elif rc.optim == 'adamw':
        OptimCls = AdamW
    else:
        raise ValueError('invalid optimizer')
    for i in optimizer_grouped_parameters:
        i['init_lr'] = i['lr']
    optimizer = OptimCls(optimizer_grouped_parameters, lr=rc.learning_rate, betas=tuple(rc.betas))
    optimizer.new_params_name = new_params_name
    optimizer.new_lr = rc.new_lr
    optimizer.basic_lr = rc.learning_rate
    optimizer.clip_lr_visual = rc.clip_lr
    optimizer.clip_lr_visual_len = len(clip_params_visual)
    optimizer.zero_grad()
    if checkpoint_optim:
        optimizer.load_state_dict(checkpoint_optim)
        del checkpoint_optim
    return optimizer
