"""AdamW over the flat parameter arena (devias_b200/arena.py) in one kernel pass (csrc/optim.cu).

Drop-in for the torch.optim.AdamW instance utils/optim_factory.py:94-178 creates: same `param_groups` protocol -- the engine
writes `param_group['lr'] = lr_schedule[it] * param_group['lr_scale']` and `param_group['weight_decay']` every iteration
(engine/engine_for_slot.py:91-97) as python floats; `step()` packs them into a small pinned buffer and ships them to DEVICE
memory, from where the kernel reads them.  A CUDA graph that captured the update therefore follows the schedule: replay =
`sync_hyper()` (host, async copy) + graph launch.  Gradient clipping (`max_norm`, utils/utils.py NativeScaler /
engine_for_slot.py:150-156) is folded into the same pass through the arena's sum of squares.
"""
from __future__ import annotations

import torch

from . import _lib
from .arena import ALIGN, ParamArena

_RING = 16


class ArenaAdamW(torch.optim.Optimizer):
    #: step() leaves the gradient arena zero-filled (no separate zero_grad pass is needed)
    zeroes_grad_in_step = True

    def __init__(self, params, arena: ParamArena, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=0.0,
                 grad_scale=1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.arena = arena
        if not arena.data.is_cuda:
            raise RuntimeError('ArenaAdamW runs on CUDA only (no CPU fallback)')
        dev = arena.data.device
        self.max_norm = float(max_norm or 0.0)
        self.grad_scale = float(grad_scale)
        group_of = {}
        for gi, g in enumerate(self.param_groups):
            assert tuple(g['betas']) == tuple(betas) and g['eps'] == eps, 'betas / eps are global in the arena kernel'
            for p in g['params']:
                if not arena.contains(p):
                    raise ValueError('every optimised parameter must live in the ParamArena')
                group_of[id(p)] = gi
        self._frozen_group = len(self.param_groups)          # arena parameters outside every group: lr = 0, wd = 0
        segs = arena.segments()
        assert len(segs) <= 1024
        start = [s for s, _ in segs] + [arena.numel // ALIGN]
        group = [group_of.get(id(p), self._frozen_group) for _, p in segs]
        assert max(group) < 255
        self.seg_start = torch.tensor(start, dtype=torch.int32, device=dev)
        self.seg_group = torch.tensor(group, dtype=torch.int32, device=dev)
        self.n_seg = len(segs)
        self.exp_avg = torch.zeros_like(arena.data)
        self.exp_avg_sq = torch.zeros_like(arena.data)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.n_hyper = 8 + 2 * (len(self.param_groups) + 1)
        self.hyper = torch.zeros(self.n_hyper, device=dev, dtype=torch.float32)
        self._host = [torch.zeros(self.n_hyper, dtype=torch.float32).pin_memory() for _ in range(_RING)]
        self._host_ev = [None] * _RING
        self._t = 0
        for _, p in segs:                                     # state views, for checkpointing code that walks optimizer.state
            if id(p) in group_of:
                o = arena.offset[id(p)]
                self.state[p] = {'exp_avg': self.exp_avg[o:o + p.numel()].view(p.shape),
                                 'exp_avg_sq': self.exp_avg_sq[o:o + p.numel()].view(p.shape)}

    # ------------------------------------------------------------------------------------------
    def sync_hyper(self, advance=True):
        """ship the current python-side hyper-parameters (param_groups' lr / weight_decay, step count for the bias corrections)
        to the device, ordered on the current stream BEFORE the next update kernel / graph replay"""
        if advance:
            self._t += 1
        t = max(self._t, 1)
        b1, b2 = self.defaults['betas']
        slot = self._t % _RING
        ev = self._host_ev[slot]
        if ev is not None:
            ev.synchronize()                                  # the copy that last used this pinned slot has run (it long has)
        h = self._host[slot]
        h[0], h[1], h[2], h[3], h[4] = 1.0 - b1 ** t, 1.0 - b2 ** t, b1, b2, self.defaults['eps']
        h[5], h[6], h[7] = self.max_norm, self.grad_scale, 0.0
        for gi, g in enumerate(self.param_groups):
            h[8 + 2 * gi] = float(g['lr'])
            h[9 + 2 * gi] = float(g['weight_decay'])
        h[8 + 2 * self._frozen_group] = 0.0
        h[9 + 2 * self._frozen_group] = 0.0
        self.hyper.copy_(h, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._host_ev[slot] = ev

    def launch(self, zero_grad=True, grad16=None):
        """the device work of one update (capturable): [sum of squares of the gradient arena] + the fused update pass.
        grad16: bf16 arena holding the (all-reduced) gradient values to use instead of the fp32 arena's"""
        a = self.arena
        s = torch.cuda.current_stream().cuda_stream
        lib = _lib.lib()
        clip = self.max_norm > 0
        if clip:
            _lib.check(lib.devias_sumsq_f32(a.grad.data_ptr(), a.numel, self.sumsq.data_ptr(), s), 'sumsq_f32')
        shadow = a.view16(a.params[0])  # makes sure the shadow exists
        del shadow
        _lib.check(lib.devias_adamw_arena(a.data.data_ptr(), a.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                          a._shadow.data_ptr(), self.seg_start.data_ptr(), self.seg_group.data_ptr(), self.n_seg,
                                          self.hyper.data_ptr(), self.sumsq.data_ptr() if clip else None, a.numel,
                                          int(zero_grad), None if grad16 is None else grad16.data_ptr(), s), 'adamw_arena')
        a.mark_fresh()                                        # the pass rewrote the bf16 shadow of every parameter

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.arena.attach_grads()
        self.sync_hyper()
        self.launch()
        return loss

    def zero_grad(self, set_to_none: bool = False):
        """gradients live in the arena: they are zero-filled in place (never set to None)"""
        self.arena.zero_grad()

    def grad_norm(self) -> torch.Tensor:
        """total gradient norm seen by the last clipped step (device scalar; what clip_grad_norm_ returns)"""
        return self.sumsq.sqrt() * abs(self.grad_scale)

