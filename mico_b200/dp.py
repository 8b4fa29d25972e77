"""Data-parallel gradient synchronisation for the MiCo training step.

Replaces the reference's two forms of gradient sync -- DistributedDataParallel (data/utils/build_model.py:57) or, when
activation checkpointing is on, the manual per-parameter ``dist.all_reduce(p.grad, async_op=True)`` loop of
data/utils/pipeline.py:93-99 (SUM, no division by the world size) -- with

  * ``FlatGrads``: ONE persistent fp32 gradient buffer for the whole model.  The ViT tower's segment is laid out in the
    tower's launch order, so the gradients of a block are contiguous and become final together; every ``p.grad`` is a
    view into the buffer (the tower's backward writes its slices in place, autograd accumulates the rest in place), and
    the fused AdamW (mico_b200.optim) reads the same memory.  No packing / unpacking copies.
  * ``GradSync``: NCCL ``all_reduce(SUM)`` of contiguous bucket slices on a side stream, overlapped with the backward pass:
    the non-tower segment (BERT, heads, embeddings: final when the tower's backward starts, because every autograd node
    created after the tower's forward runs before it) goes first, then one all-reduce per ``bucket_blocks`` finished tower
    blocks.  ``finish()`` makes the compute stream wait for the last bucket.

Semantics are the reference's: gradients are SUMMED over ranks (divide in the optimizer with ``AdamW.grad_scale`` if an
average is wanted).  ``dtype=torch.bfloat16`` halves the NVLink bytes (cast -> all-reduce -> accumulate back) at bf16
summation accuracy; the default fp32 path is exact and, on NVSwitch, 4.75 GB per step costs ~12 ms of collective time
against a backward pass of more than a second.
"""
import torch
import torch.distributed as dist

from . import ops
from .ops import F32, MicoError


def _align4(n):
    return (n + 3) // 4 * 4


class FlatGrads:
    """Persistent flat fp32 gradient buffer: [tower segment | rest segment]."""

    def __init__(self, model, tower=None):
        if tower is None:
            tower = getattr(getattr(model, "vision_encoder", None), "visual", None)
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise MicoError("FlatGrads: the model has no trainable parameters")
        dev = params[0].device
        self.model, self.tower = model, tower
        t_params, t_sizes = [], []
        if tower is not None and hasattr(tower, "_flat_params"):
            t_params = tower._flat_params()
            t_sizes = tower.flat_grad_sizes(t_params)
        t_ids = {id(p) for p in t_params if p is not None}
        seen, self.rest = set(), []
        for p in params:             # shared Parameters (tied decoder / word embeddings) appear once
            if id(p) in t_ids or id(p) in seen:
                continue
            seen.add(id(p))
            self.rest.append(p)
        self.n_tower = sum(t_sizes)
        n_rest = sum(_align4(p.numel()) for p in self.rest)
        self.buf = torch.zeros(self.n_tower + n_rest, device=dev, dtype=F32)
        self.tower_buf = self.buf[:self.n_tower]
        self.rest_buf = self.buf[self.n_tower:]
        self._views = []
        off = 0
        for p, n in zip(t_params, t_sizes):
            if p is not None and p.requires_grad:
                self._views.append((p, self.buf[off:off + p.numel()].view(p.shape)))
            off += n
        self._tower_views = len(self._views)
        off = self.n_tower
        for p in self.rest:
            self._views.append((p, self.buf[off:off + p.numel()].view(p.shape)))
            off += _align4(p.numel())
        if self.n_tower:
            tower.flat_grad = self.tower_buf
        # which non-tower parameters received a gradient this step (the reference's optimizer skips p.grad is None)
        self._touched = set()
        self._hooks = [p.register_post_accumulate_grad_hook(self._mark) for p in self.rest]
        self.zero_grad()

    def _mark(self, p):
        self._touched.add(id(p))

    def zero_grad(self):
        """Start a step: the rest segment is zeroed (autograd accumulates into it), the tower segment is overwritten by the
        tower's single backward pass, and every p.grad is (re-)attached to its slice."""
        self.rest_buf.zero_()
        self._touched.clear()
        if self.n_tower:
            self.tower._flat_grad_written = False
        for p, v in self._views:
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def detach_unused(self):
        """After backward: parameters that took no part in this step's graph get p.grad = None, so that the optimizer skips
        them like the reference's does (build_optimizer.py:150-151) instead of applying weight decay to them."""
        wrote_tower = self.n_tower and self.tower._flat_grad_written
        n = 0
        for i, (p, _) in enumerate(self._views):
            used = wrote_tower if i < self._tower_views else (id(p) in self._touched)
            if not used:
                p.grad = None
                n += 1
        return n

    def close(self):
        for h in self._hooks:
            h.remove()
        if self.n_tower:
            self.tower.flat_grad = None


class GradSync:
    """Overlapped data-parallel gradient SUM over a FlatGrads buffer (pipeline.py:93-99 semantics)."""

    def __init__(self, flat, bucket_blocks=5, group=None, dtype=F32, tail_blocks=2):
        if not (dist.is_available() and dist.is_initialized()):
            raise MicoError("GradSync needs an initialised torch.distributed process group")
        self.flat, self.group, self.dtype = flat, group, dtype
        self.bucket_blocks, self.tail_blocks = max(1, int(bucket_blocks)), int(tail_blocks)
        self.world = dist.get_world_size(group)
        dev = flat.buf.device
        self.cuda = dev.type == "cuda"
        self.comm = torch.cuda.Stream(device=dev) if self.cuda else None
        self._pending, self._blocks_done = [], 0
        self.bytes_reduced = 0
        self.verify = None
        tower = flat.tower
        if flat.n_tower:
            self.n_blocks = len(tower.blocks)
            tower.grad_begin_hook = self._on_tower_begin
            tower.grad_bucket_hook = self._on_bucket

    # ---- collective on one contiguous slice of the flat buffer
    def _reduce(self, t):
        if t.numel() == 0:
            return
        if self.verify is not None:      # self-check: keep this rank's addend of the slice (see check())
            self.verify.append((t, t.clone()))
        self.bytes_reduced += t.numel() * (2 if self.dtype == torch.bfloat16 else 4)
        if not self.cuda:                      # gloo CPU tests: synchronous
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return
        self.comm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.comm):
            if self.dtype == torch.bfloat16:
                tb = ops.cast_bf16(t)
                dist.all_reduce(tb, op=dist.ReduceOp.SUM, group=self.group)
                ops.cast_f32_from_bf16(tb, t)
                tb.record_stream(self.comm)
            else:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def _on_tower_begin(self):
        self._pending, self._blocks_done = [], 0
        self._reduce(self.flat.rest_buf)
        self._rest_done = True

    def _on_bucket(self, bucket):
        """called by the tower with each finished block's gradient slice (last block first), then the tower-level slice"""
        base = self.flat.tower_buf.data_ptr()
        lo = (bucket.data_ptr() - base) // 4
        self._pending.append((lo, bucket.numel()))
        is_last = lo == 0
        self._blocks_done += 0 if is_last else 1
        tail = self.n_blocks - self._blocks_done < self.tail_blocks     # the last buckets stay small: their all-reduce is exposed
        if len(self._pending) < self.bucket_blocks and not is_last and not tail:
            return
        lo = min(o for o, _ in self._pending)
        n = sum(k for _, k in self._pending)
        self._pending = []
        self._reduce(self.flat.tower_buf[lo:lo + n])

    def begin_step(self, verify=False):
        self._rest_done = False
        self.bytes_reduced = 0
        self.verify = [] if verify else None

    def check(self):
        """Self-check of one step run with begin_step(verify=True), after finish(): (a) the reduced slices tile the whole
        flat buffer exactly once, (b) every slice equals an independent all-reduce of the per-rank addends snapshotted
        before the overlapped reduction, (c) every p.grad still aliases the buffer.  Returns (ok, message)."""
        if self.verify is None:
            return False, "begin_step(verify=True) was not called"
        base = self.flat.buf.data_ptr()
        spans = sorted(((t.data_ptr() - base) // 4, t.numel()) for t, _ in self.verify)
        pos = 0
        for lo, n in spans:
            if lo != pos:
                return False, f"gradient buffer not tiled exactly once by the reduced slices (gap / overlap at {pos} vs {lo})"
            pos = lo + n
        if pos != self.flat.buf.numel():
            return False, f"reduced {pos} of {self.flat.buf.numel()} gradient elements"
        worst = 0.0
        for t, local in self.verify:
            dist.all_reduce(local, op=dist.ReduceOp.SUM, group=self.group)
            denom = local.abs().max().clamp_min(1e-20)
            worst = max(worst, float((t - local).abs().max() / denom))
        tol = 2e-2 if self.dtype == torch.bfloat16 else 1e-5
        for p, v in self.flat._views:
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                return False, "a parameter's .grad no longer aliases the flat buffer"
        self.verify = None
        return worst <= tol, f"{len(spans)} slices, max |synced - independent sum| / max|sum| = {worst:.2e} (tol {tol:g})"

    def finish(self):
        """After loss.backward(): reduce whatever the hooks did not cover and join the communication stream."""
        wrote_tower = self.flat.n_tower and self.flat.tower._flat_grad_written
        if not wrote_tower:           # no tower pass in this step's graph: nothing was hooked
            self._reduce(self.flat.rest_buf)
        elif not getattr(self, "_rest_done", False):
            self._reduce(self.flat.rest_buf)
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.comm)

    def close(self):
        if self.flat.n_tower:
            self.flat.tower.grad_begin_hook = None
            self.flat.tower.grad_bucket_hook = None
