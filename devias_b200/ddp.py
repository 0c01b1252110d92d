"""Data-parallel gradient exchange for the clip batch (the only collective on the DEVIAS path: the DDP / DeepSpeed
ZeRO-0 all-reduce of run_slot_finetuning.py:552-563).

One process per GPU.  Gradients live in a few flat fp32 buckets (parameters' .grad are views into them, so autograd
accumulates straight into the communication buffer - no gather/scatter copies); buckets are filled in reverse
execution order and each one is all-reduced asynchronously (NCCL over NVLink/NVSwitch, its own stream) as soon as its
last gradient has been accumulated, overlapping the rest of backward.  `finish()` waits and applies the 1/world mean.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, module: torch.nn.Module, bucket_mb: float = 48.0, process_group=None, first_bucket_mb: float = 8.0):
        from .arena import ParamArena
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # parameters AND gradients live in the module's flat arena (reverse execution order: head / agg first, patch_embed last)
        self.param_arena = pa = ParamArena.of(module)
        params = pa.params
        self.params = params
        self.buckets: List[torch.Tensor] = []
        self.bucket_of = {}
        self._views = {p: pa.grad_view(p) for p in params}
        cap = int(first_bucket_mb * (1 << 20) // 4)
        cur, cur_n = [], 0
        groups = []
        for p in params:
            n = (p.numel() + 7) // 8 * 8
            if cur and cur_n + n > cap:
                groups.append(cur)
                cur, cur_n = [], 0
                cap = int(bucket_mb * (1 << 20) // 4)
            cur.append(p)
            cur_n += n
        if cur:
            groups.append(cur)
        # all buckets are slices of ONE arena, so the un-overlapped path can exchange everything with a single collective
        self.arena = pa.grad
        self._avg = dist.is_initialized() and dist.get_backend(process_group) == 'nccl'   # NCCL averages natively
        for bi, g in enumerate(groups):
            lo, hi = pa.range_of(g)
            for p in g:
                self.bucket_of[p] = bi
            self.buckets.append(self.arena[lo:hi])
        self.sizes = [len(g) for g in groups]
        self._pending = [0] * len(groups)
        self._works: List[Optional[object]] = []
        self._launched = [False] * len(groups)
        self.enabled = True                                      # False on gradient-accumulation micro-steps (begin_backward)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]
        self.zero_grad()

    # ------------------------------------------------------------------------------------------
    def zero_grad(self):
        """replaces optimizer.zero_grad(): one memset per bucket, .grad views stay attached"""
        self.arena.zero_()
        for p in self.params:
            if p.grad is None or p.grad.data_ptr() != self._views[p].data_ptr():
                p.grad = self._views[p]
        self._pending = list(self.sizes)
        self._launched = [False] * len(self.buckets)
        self._works = []

    def reset_step(self):
        """start of a new step when somebody else (the arena optimizer pass) already zero-filled the gradients"""
        self._pending = list(self.sizes)
        self._launched = [False] * len(self.buckets)
        self._works = []

    def begin_backward(self, exchange=True):
        """call before every backward of a step.  exchange=False marks a gradient-accumulation micro-step: its gradients are
        summed into the arena but not counted in, so no bucket is all-reduced before the last micro-step's backward has added
        its share (a bucket exchanged early would be exchanged without the later micro-steps' gradients)."""
        self.enabled = bool(exchange)
        if exchange:
            assert not any(self._launched), 'a bucket was already exchanged in this step (finish() / zero_grad() missing?)'
            self._pending = list(self.sizes)

    def _on_grad(self, p):
        if p.grad.data_ptr() != self._views[p].data_ptr():       # someone replaced .grad (e.g. set_to_none): re-home it
            self._views[p].copy_(p.grad)
            p.grad = self._views[p]
        if not self.enabled:
            return
        b = self.bucket_of[p]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self._launched[b]:
            self._launch(b)

    def _launch(self, b):
        self._launched[b] = True
        if self.world > 1:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self._works.append(dist.all_reduce(self.buckets[b], op=op, group=self.pg, async_op=True))

    def finish(self):
        """call after backward, before the optimizer step"""
        if not self.enabled:
            if self.world > 1:
                raise RuntimeError('GradReducer.finish() on a reducer whose exchange is switched off: the ranks would step on '
                                   'their local gradients (call begin_backward(exchange=True) before the last backward)')
            return
        for b in range(len(self.buckets)):
            if not self._launched[b]:                            # parameters that received no gradient this step
                self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []
        if self.world > 1 and not self._avg:
            self.arena.mul_(1.0 / self.world)

    def allreduce_all(self):
        """un-overlapped exchange of every bucket (used between the two CUDA graphs of a graphed step, where the
        per-parameter hooks do not run): all buckets are enqueued back to back, then awaited and averaged"""
        if self.world <= 1:
            return
        if self._avg:
            dist.all_reduce(self.arena, op=dist.ReduceOp.AVG, group=self.pg)
        else:
            dist.all_reduce(self.arena, op=dist.ReduceOp.SUM, group=self.pg)
            self.arena.mul_(1.0 / self.world)

    def range_of(self, params):
        """[lo, hi) slice of the arena holding exactly the gradients of `params` (which must be a prefix or a suffix of
        the reducer's reverse-execution parameter order)"""
        return self.param_arena.range_of(params)

    def buf16(self):
        """bf16 twin of the gradient arena (created on first use): staging buffer of the compressed exchange"""
        if getattr(self, '_buf16', None) is None:
            self._buf16 = torch.zeros(self.arena.numel(), device=self.arena.device, dtype=torch.bfloat16)
        return self._buf16

    def allreduce_range(self, lo, hi, async_op=False, compress=False):
        """all-reduce (mean) of arena[lo:hi].  compress=True exchanges bf16: the range is cast into the bf16 twin, reduced there and
        LEFT there (half the bytes over NVLink; the optimizer pass reads the reduced values from the bf16 buffer, see
        ArenaAdamW.launch(grad16=...)); `decompress_range` copies them back for consumers that need fp32 gradients."""
        if self.world <= 1 and not compress:
            return None
        if compress:
            from . import ops
            b = self.buf16()[lo:hi]
            ops.cast_bf16(self.arena[lo:hi], b)
            if self.world <= 1:
                return None
            assert self._avg, 'the compressed exchange uses the native average of the nccl backend'
            return dist.all_reduce(b, op=dist.ReduceOp.AVG, group=self.pg, async_op=async_op)
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        w = dist.all_reduce(self.arena[lo:hi], op=op, group=self.pg, async_op=async_op)
        if not self._avg:
            assert not async_op, 'SUM + scale needs the synchronous path'
            self.arena[lo:hi].mul_(1.0 / self.world)
        return w

    def decompress_range(self, lo, hi):
        self.arena[lo:hi].copy_(self.buf16()[lo:hi])

    def remove(self):
        for h in self._hooks:
            h.remove()
