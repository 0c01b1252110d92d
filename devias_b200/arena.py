"""Flat parameter arena: every trainable parameter of a module becomes a view of ONE fp32 buffer, its `.grad` a view of a
second one at the same offset, and (on CUDA) a bf16 shadow at the same offset is what the tensor-core kernels read.

Why (SURVEY.md section 8b "kernels may keep private bf16 copies but must refresh them when the fp32 master changes"; section 8e):
  * the gradient all-reduce of run_slot_finetuning.py:552-563 runs on contiguous ranges of the gradient arena, no bucket copies;
  * weight-gradient kernels reduce-add straight into the gradient arena (`param._grad_sink`), so autograd launches no
    per-parameter accumulation kernels;
  * the optimizer (devias_b200/optim.py) updates parameter, moments and bf16 shadow and clears the gradient in one pass.

Order inside the arena = REVERSE of `module.parameters()` (~ reverse execution order: heads / aggregation block first, patch
embedding last), so the gradients that complete first during backward form a prefix.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Tuple

import torch

ALIGN = 8  # elements: 32 bytes fp32 / 16 bytes bf16 -- every view stays 16-byte aligned for the vector kernels


def _round(n: int) -> int:
    return (n + ALIGN - 1) // ALIGN * ALIGN


class ParamArena:
    def __init__(self, module: torch.nn.Module):
        seen, params = set(), []
        for p in module.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        assert params, 'module has no trainable parameters'
        params = params[::-1]
        dev = params[0].device
        assert all(p.device == dev and p.dtype == torch.float32 for p in params), 'fp32 parameters on one device expected'
        self.params: List[torch.nn.Parameter] = params
        self.offset: Dict[int, int] = {}
        total = 0
        for p in params:
            self.offset[id(p)] = total
            total += _round(p.numel())
        self.numel = total
        self.data = torch.zeros(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self._shadow = None
        self._stamp = None
        with torch.no_grad():
            for p in params:
                v = self._view(self.data, p)
                v.copy_(p.data)
                p.data = v
                old = p.grad
                gv = self._view(self.grad, p)
                if old is not None:
                    gv.copy_(old)
                p.grad = gv
                p._grad_sink = gv             # read by the backward Functions (devias_b200/functional.py) in direct mode
        self._ptrs = [p.data_ptr() for p in params]

    # ------------------------------------------------------------------------------------------
    @classmethod
    def of(cls, module: torch.nn.Module) -> 'ParamArena':
        """the module's arena (created on first use; rebuilt when the parameters were moved, e.g. by .cuda())"""
        a = module.__dict__.get('_param_arena')
        if a is None or not a.valid():
            a = cls(module)
            module.__dict__['_param_arena'] = a
        return a

    def _view(self, flat: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
        o = self.offset[id(p)]
        return flat[o:o + p.numel()].view(p.shape)

    def valid(self) -> bool:
        return all(p.data_ptr() == q for p, q in zip(self.params, self._ptrs))

    def contains(self, p) -> bool:
        return id(p) in self.offset

    def grad_view(self, p) -> torch.Tensor:
        return self._view(self.grad, p)

    def attach_grads(self):
        """point every .grad back at the arena (after someone set them to None or replaced them)"""
        for p in self.params:
            gv = p._grad_sink
            if p.grad is None or p.grad.data_ptr() != gv.data_ptr():
                p.grad = gv

    def zero_grad(self):
        self.grad.zero_()
        self.attach_grads()

    # ---------------------------------------------------------------- bf16 shadow for the tensor-core kernels (CUDA only)
    def view16(self, p) -> torch.Tensor:
        if self._shadow is None:
            self._shadow = torch.empty(self.numel, device=self.data.device, dtype=torch.bfloat16)
            self._stamp = None
        return self._view(self._shadow, p)

    def _version_stamp(self) -> int:
        return sum(p._version for p in self.params)

    def refresh16(self, force: bool = False):
        """re-cast the shadow when any parameter was modified through torch since the last refresh (tensor versions);
        kernels that write parameters through raw pointers keep the shadow fresh themselves and call `mark_fresh`."""
        if self._shadow is None:
            self.view16(self.params[0])
        stamp = self._version_stamp()
        if force or stamp != self._stamp:
            from . import ops
            ops.cast_bf16(self.data, self._shadow)
            self._stamp = stamp

    def mark_fresh(self):
        self._stamp = self._version_stamp()

    def invalidate16(self):
        self._stamp = None

    # ---------------------------------------------------------------- ranges (gradient exchange)
    def range_of(self, params: Iterable[torch.nn.Parameter]) -> Tuple[int, int]:
        """[lo, hi) slice of the arena holding exactly `params` (which must be contiguous in arena order)"""
        ids = {id(p) for p in params}
        lo = hi = None
        inside = 0
        for p in self.params:
            if id(p) in ids:
                o = self.offset[id(p)]
                lo = o if lo is None else lo
                hi = o + _round(p.numel())
                inside += _round(p.numel())
        assert lo is not None and hi - lo == inside, 'parameters are not contiguous in the gradient arena'
        return lo, hi

    def segments(self):
        """[(start granule, parameter)] in arena order (granule = ALIGN elements)"""
        return [(self.offset[id(p)] // ALIGN, p) for p in self.params]
