"""Train / eval steps mirroring engine/engine_for_slot.py (train_class_batch :50-56, the body of train_one_epoch
:98-171 for the non-DeepSpeed branch, validation_one_epoch :217-253) without its per-step host syncs."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F


def train_class_batch(model, scene_model, samples, target, train_criterion, fg_mask=None, teacher_logits=None):
    """engine/engine_for_slot.py:50-56.  `scene_model` is the frozen scene teacher (any callable returning
    (token, logits)); `teacher_logits` may be given directly when the teacher forward is run elsewhere."""
    student_output = model(samples)
    if teacher_logits is None:
        with torch.no_grad():
            teacher_output = scene_model(samples, return_attn=False)
    else:
        teacher_output = (None, teacher_logits)
    total_loss, output, loss_dict = train_criterion(model, student_output, teacher_output, target, fg_mask=fg_mask)
    return total_loss, output, loss_dict


def train_step(model, scene_model, train_criterion, optimizer, samples, targets, fg_mask, teacher_logits=None,
               update_freq=1, do_update=True, max_norm=0, reducer=None):
    """One iteration of train_one_epoch's loop body (engine/engine_for_slot.py:120-166, loss_scaler branch with bf16:
    no GradScaler is needed).  Returns (loss tensor, output, loss_dict) without synchronising."""
    from .functional import direct_grads
    arena_opt = getattr(optimizer, 'zeroes_grad_in_step', False)      # devias_b200.optim.ArenaAdamW
    # without per-parameter hooks to serve (no reducer) the weight gradients go straight into the optimizer's gradient arena
    with direct_grads(arena_opt and reducer is None):
        return _train_step(model, scene_model, train_criterion, optimizer, samples, targets, fg_mask, teacher_logits, update_freq,
                           do_update, max_norm, reducer, arena_opt)


def _train_step(model, scene_model, train_criterion, optimizer, samples, targets, fg_mask, teacher_logits, update_freq, do_update,
                max_norm, reducer, arena_opt):
    loss, output, loss_dict = train_class_batch(model, scene_model, samples, targets, train_criterion, fg_mask, teacher_logits)
    if reducer is not None:
        # gradient accumulation (engine/engine_for_slot.py:147-166 with update_freq > 1): micro-steps accumulate locally in the
        # arena, only the backward of the LAST micro-step counts gradients in and launches the bucket all-reduces
        reducer.begin_backward(exchange=bool(do_update))
    (loss / update_freq).backward()
    if do_update:
        if reducer is not None:
            reducer.finish()
        if arena_opt:
            if max_norm and max_norm > 0:
                optimizer.max_norm = float(max_norm)             # clipping is folded into the update pass
        elif max_norm and max_norm > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        optimizer.step()
        if arena_opt:
            if reducer is not None:
                reducer.reset_step()     # the update pass already zero-filled the gradient arena
        elif reducer is not None:
            reducer.zero_grad()          # one memset per bucket; .grad stay views of the communication buffers
        else:
            optimizer.zero_grad(set_to_none=True)
    return loss.detach(), output, loss_dict


@torch.no_grad()
def validation_step(model, videos, target):
    """engine/engine_for_slot.py:234-239: logits of the action slot over the unified C+365 row, CE loss, top-1/5."""
    _, (output, scene_output, attn), _ = model(videos)
    loss = F.cross_entropy(output.float(), target)
    top5 = output.topk(5, dim=1).indices
    acc1 = (top5[:, 0] == target).float().mean() * 100.0
    acc5 = (top5 == target.unsqueeze(1)).any(dim=1).float().mean() * 100.0
    return output, scene_output, loss, acc1, acc5


class _TapFn(torch.autograd.Function):
    """identity whose output can be given to autograd.backward(inputs=...) as a cut point: the engine executes this
    (free) node to fill .grad, but not the producer of its input"""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g


class GraphedTrainStep:
    """The whole training step (student forward, [frozen teacher forward,] TrainLoss, backward, gradient exchange, optimizer
    update) captured once into CUDA graphs and replayed: the ~2000 kernel launches of a step cost a handful of graph launches,
    which removes the host launch overhead that otherwise bounds small-batch steps ('launch-bound').

    `batches`: one or more dicts of STATIC device tensors {clip, target, fg, fgf[, teacher]}; one graph set is captured per
    dict (they share a memory pool), so input buffers can be double-buffered against host->device copies.  Without a
    `teacher` entry the frozen `scene_model` runs inside the captured step (engine/engine_for_slot.py:52-53).

    Two optimizer modes:
      * `devias_b200.optim.ArenaAdamW` (arena mode, the fast path): parameters, gradients and moments are flat arenas; weight
        gradients are reduce-added straight into the gradient arena (`functional.direct_grads`), the update is ONE kernel
        pass in its own graph, hyper-parameters (lr / weight-decay schedule, bias corrections, `max_norm`) are read from device
        memory and refreshed by `optimizer.sync_hyper()` before every update replay; `update_freq` micro-steps accumulate
        in the arena (engine/engine_for_slot.py:147-166).
      * any capture-safe torch optimizer (e.g. AdamW(fused=True, capturable=True) or SGD): its hyper-parameters are frozen at
        capture time unless given as device tensors (`set_lr`); `update_freq` must be 1.

    With a `reducer` (data parallel) the backward is cut at the encoder blocks `cuts` (descending): graph piece i runs while
    the NCCL all-reduce of the gradient range completed by piece i-1 is in flight on the communication stream; only the
    last, smallest range (blocks below the last cut + patch embedding) is exchanged un-overlapped before the update graph.
    """

    def __init__(self, model, train_criterion, optimizer, batches, reducer=None, warmup=3, split_block=3, scene_model=None,
                 update_freq=1, cuts=None, grad_exchange='fp32'):
        from . import _lib
        from .optim import ArenaAdamW
        self.model, self.crit, self.opt, self.reducer = model, train_criterion, optimizer, reducer
        self.scene_model = scene_model
        self.batches = list(batches)
        self.update_freq = int(update_freq)
        self.arena_mode = isinstance(optimizer, ArenaAdamW)
        assert grad_exchange in ('fp32', 'bf16')
        #: 'bf16': every range is cast to bf16, all-reduced in bf16 (half the NVLink bytes) and consumed from the bf16 buffer by the
        #: arena optimizer pass (torch optimizers / gradient clipping get it copied back to fp32 first)
        self.compress = grad_exchange == 'bf16' and reducer is not None
        self._grad16 = self.compress and self.arena_mode and not (getattr(optimizer, 'max_norm', 0) > 0)
        assert self.update_freq == 1 or self.arena_mode, 'gradient accumulation across replays needs the gradient arena (ArenaAdamW)'
        self.split = reducer is not None
        self._taps, self._hooks, self._active = {}, [], False
        self._micro = 0
        blocks = list(model.blocks)
        if not self.arena_mode:
            for g in optimizer.param_groups:         # python-float hyper-parameters would be baked into the graph
                if g.get('capturable', False) and not torch.is_tensor(g['lr']):
                    g['lr'] = torch.tensor(float(g['lr']), device=next(model.parameters()).device, dtype=torch.float32)
        if self.split:
            if cuts is None:
                cuts = [split_block]
            # cuts = []: one backward piece and ONE all-reduce of the whole arena after it (nothing overlaps the backward: the
            # collective then has every SM and NVLink to itself instead of taking SMs away from the persistent compute kernels)
            cuts = sorted({max(1, min(int(c), len(blocks) - 1)) for c in cuts}, reverse=True)
            self.cuts = cuts
            # parameter sets per backward piece: piece 0 = everything above the first cut, ..., last = below the last cut
            bounds = cuts + [0]
            piece_params = []
            hi = len(blocks)
            for c in bounds:
                piece_params.append([p for b in blocks[c:hi] for p in b.parameters() if p.requires_grad])
                hi = c
            piece_params[-1] = piece_params[-1] + [p for p in model.patch_embed.parameters() if p.requires_grad]
            if isinstance(getattr(model, 'pos_embed', None), torch.nn.Parameter) and model.pos_embed.requires_grad:
                piece_params[-1].append(model.pos_embed)   # sits between patch_embed and blocks.0 in parameters() order
            low_ids = {id(p) for ps in piece_params for p in ps}
            piece_params[0] = [p for p in reducer.params if id(p) not in low_ids] + piece_params[0]
            self.piece_params = piece_params
            self.ranges = [reducer.range_of(ps) for ps in piece_params]
            for c in cuts:
                self._hooks.append(blocks[c].register_forward_pre_hook(self._make_grab(c)))
        else:
            self.cuts = []
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                      # warm-up off the capture stream (allocator, lazy inits)
            for i in range(warmup):
                self._pieces_eager(self.batches[i % len(self.batches)])
                self._exchange_all()
                if self.arena_mode:
                    self.opt.sync_hyper()
                self._update()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, self.losses = [], []
        pool = None
        for b in self.batches:
            n0 = _lib.launch_count()
            gs = []
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                loss = self._piece0(b)
                if not self.split and not self.arena_mode:
                    self.opt.step()
            pool = g.pool()
            gs.append(g)
            for i in range(1, len(self.cuts) + 1):
                gi = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gi, pool=pool):
                    self._piece(i)
                gs.append(gi)
            self.launches_per_step = _lib.launch_count() - n0
            self.graphs.append(gs)
            self.losses.append(loss)
            if not self.split and not self.arena_mode:
                self.opt.zero_grad(set_to_none=True)
        self.update_graph = None
        if self.split or self.arena_mode:
            n0 = _lib.launch_count()
            self.update_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.update_graph, pool=pool):
                self._update()
            self.launches_per_step += _lib.launch_count() - n0

    # ------------------------------------------------------------------------------------------
    def close(self):
        """remove the cut-point hooks from the model (the captured graphs stay replayable)"""
        for h in self._hooks:
            h.remove()
        self._hooks = []

    def set_lr(self, values):
        """torch-optimizer mode: write the learning rate(s) into the device tensors the captured update reads
        (one value or one per param group, already multiplied by the group's lr_scale as engine_for_slot.py:91-97 does)"""
        groups = self.opt.param_groups
        vals = [values] * len(groups) if not isinstance(values, (list, tuple)) else values
        for g, v in zip(groups, vals):
            if torch.is_tensor(g['lr']):
                g['lr'].fill_(float(v))
            else:
                g['lr'] = float(v)

    def _make_grab(self, c):
        def grab(module, args):
            x = args[0]
            if self._active and torch.is_grad_enabled() and x.requires_grad:
                y = _TapFn.apply(x)
                y.retain_grad()
                self._taps[c] = (x, y)
                return (y,) + tuple(args[1:])
            return None
        return grab

    @contextlib.contextmanager
    def _ctx(self):
        """direct gradient accumulation (arena mode) and the reducer's per-parameter hooks switched off for OUR backward
        pieces only (python hooks do not run on replay; the exchange is issued per range in __call__)"""
        from .functional import direct_grads
        red = self.reducer
        prev = red.enabled if red is not None else None
        if red is not None:
            red.enabled = False
        try:
            with direct_grads(self.arena_mode):
                yield
        finally:
            if red is not None:
                red.enabled = prev

    def _piece0(self, b):
        self._active = True
        try:
            with self._ctx():
                loss, _, _ = train_class_batch(self.model, self.scene_model, b['clip'], b['target'], self.crit, (b['fg'], b['fgf']),
                                               teacher_logits=b.get('teacher'))
                root = loss / self.update_freq if self.update_freq != 1 else loss
                if self.split and self.cuts:
                    torch.autograd.backward([root], inputs=self.piece_params[0] + [self._taps[self.cuts[0]][1]], retain_graph=True)
                else:
                    root.backward()
        finally:
            self._active = False
        return loss.detach()

    def _piece(self, i):
        c = self.cuts[i - 1]
        x_pre, x_post = self._taps.pop(c)
        last = i == len(self.cuts)
        inputs = list(self.piece_params[i]) + ([] if last else [self._taps[self.cuts[i]][1]])
        with self._ctx():
            torch.autograd.backward([x_pre], [x_post.grad], inputs=inputs, retain_graph=not last)
        x_post.grad = None

    def _pieces_eager(self, b):
        self._piece0(b)
        for i in range(1, len(self.cuts) + 1):
            self._piece(i)

    def _exchange_all(self):
        if self.split:
            if self.compress:
                self.reducer.allreduce_range(0, self.reducer.arena.numel(), compress=True)
                if not self._grad16:
                    self.reducer.decompress_range(0, self.reducer.arena.numel())
            else:
                self.reducer.allreduce_all()

    def _update(self):
        if self.arena_mode:
            # hyper-parameters come from device memory (sync_hyper before the replay)
            self.opt.launch(grad16=self.reducer.buf16() if self._grad16 else None)
            return
        self.opt.step()
        if self.split:
            self.reducer.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)

    def __call__(self, index=0):
        """replay one (micro-)step on static batch `index`; returns the (static) loss tensor of that graph.  With
        update_freq > 1 only every update_freq-th call exchanges gradients and updates the parameters."""
        gs = self.graphs[index]
        if self.arena_mode:
            self.opt.arena.refresh16()           # no-op unless parameters were modified through torch since the last update
        self._micro += 1
        do_update = self._micro % self.update_freq == 0
        if not self.split:
            gs[0].replay()
        else:
            works = []
            red = self.reducer
            for i, g in enumerate(gs):
                g.replay()
                if do_update:
                    last = i == len(gs) - 1
                    w = red.allreduce_range(*self.ranges[i], async_op=(red._avg and not last), compress=self.compress)
                    if w is not None and not last:
                        works.append(w)
            for w in works:
                w.wait()
            if do_update and self.compress and not self._grad16:
                red.decompress_range(0, red.arena.numel())
        if do_update and self.update_graph is not None:
            if self.arena_mode:
                self.opt.sync_hyper()
            self.update_graph.replay()
        return self.losses[index]
