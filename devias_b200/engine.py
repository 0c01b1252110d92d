"""Train / eval steps mirroring engine/engine_for_slot.py (train_class_batch :50-56, the body of train_one_epoch
:98-171 for the non-DeepSpeed branch, validation_one_epoch :217-253) without its per-step host syncs."""
from __future__ import annotations

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
    loss, output, loss_dict = train_class_batch(model, scene_model, samples, targets, train_criterion, fg_mask, teacher_logits)
    (loss / update_freq).backward()
    if do_update:
        if reducer is not None:
            reducer.finish()
        if max_norm and max_norm > 0:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        optimizer.step()
        if reducer is not None:
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
    """The whole training step (student forward, TrainLoss, backward, gradient exchange, optimizer update) captured once
    into a CUDA graph and replayed: the ~2000 kernel launches of a step cost one graph launch, which removes the host
    launch overhead that otherwise bounds small-batch steps ('launch-bound').

    `batches`: one or more dicts of STATIC device tensors {clip, target, fg, fgf, teacher}; one graph is captured per
    dict (they share a memory pool), so input buffers can be double-buffered against host->device copies.
    The optimizer must be capture-safe (e.g. torch.optim.AdamW(fused=True, capturable=True)).
    """

    def __init__(self, model, train_criterion, optimizer, batches, reducer=None, warmup=3, split_block=3):
        self.model, self.crit, self.opt, self.reducer = model, train_criterion, optimizer, reducer
        self.batches = list(batches)
        # With a reducer the step becomes three graphs around the NCCL exchange of the flat gradient arena:
        #   graph A1: forward + loss + backward down to the input of encoder block `split_block`
        #   eager   : all-reduce of the gradients produced so far (head, slots, blocks >= split_block) -- asynchronous, it
        #             overlaps graph A2 on the NCCL stream
        #   graph A2: backward of blocks < split_block and the patch embedding
        #   eager   : all-reduce of the remaining gradients
        #   graph B : optimizer update + arena memset
        # (per-parameter hooks are python and do not run on replay; NCCL inside a captured graph dead-locked here.)
        self.split = reducer is not None
        self._tap = None
        if self.split:
            reducer.enabled = False
            blocks = list(model.blocks)
            split_block = max(1, min(split_block, len(blocks) - 1))
            lower = [p for p in model.patch_embed.parameters()] + [p for b in blocks[:split_block] for p in b.parameters()]
            lower = [p for p in lower if p.requires_grad]
            low_ids = {id(p) for p in lower}
            self.lower = lower
            self.upper = [p for p in reducer.params if id(p) not in low_ids]
            self.lo_range = reducer.range_of(lower)
            self.up_range = reducer.range_of(self.upper)
            self._hook = blocks[split_block].register_forward_pre_hook(self._grab)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                      # warm-up off the capture stream (allocator, lazy inits)
            for i in range(warmup):
                self._fwd_bwd1(self.batches[i % len(self.batches)])
                self._bwd2()
                self._exchange_all()
                self._update()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graphs, self.graphs2, self.losses = [], [], []
        pool = None
        from . import _lib
        for b in self.batches:
            g = torch.cuda.CUDAGraph()
            n0 = _lib.launch_count()
            with torch.cuda.graph(g, pool=pool):
                loss = self._fwd_bwd1(b)
                if not self.split:
                    self.opt.step()
            pool = g.pool()
            g2 = None
            if self.split:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=pool):
                    self._bwd2()
            self.launches_per_step = _lib.launch_count() - n0
            self.graphs.append(g)
            self.graphs2.append(g2)
            self.losses.append(loss)
            if not self.split:
                self.opt.zero_grad(set_to_none=True)
        self.update_graph = None
        if self.split:
            self.update_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.update_graph, pool=pool):
                self._update()

    def _grab(self, module, args):
        x = args[0]
        if torch.is_grad_enabled() and x.requires_grad:
            y = _TapFn.apply(x)
            y.retain_grad()
            self._tap = (x, y)
            return (y,) + tuple(args[1:])
        return None

    def _fwd_bwd1(self, b):
        loss, _, _ = train_class_batch(self.model, None, b['clip'], b['target'], self.crit, (b['fg'], b['fgf']),
                                       teacher_logits=b['teacher'])
        if self.split:
            torch.autograd.backward([loss], inputs=self.upper + [self._tap[1]], retain_graph=True)
        else:
            loss.backward()
        return loss.detach()

    def _bwd2(self):
        if self.split:
            (x_pre, x_post), self._tap = self._tap, None
            torch.autograd.backward([x_pre], [x_post.grad], inputs=self.lower)
            x_post.grad = None

    def _exchange_all(self):
        if self.split:
            self.reducer.allreduce_all()

    def _update(self):
        self.opt.step()
        if self.split:
            self.reducer.zero_grad()
        else:
            self.opt.zero_grad(set_to_none=True)

    def __call__(self, index=0):
        """replay the step on static batch `index`; returns the (static) loss tensor of that graph"""
        self.graphs[index].replay()
        if self.split:
            w = self.reducer.allreduce_range(*self.up_range, async_op=self.reducer._avg)
            self.graphs2[index].replay()
            if w is not None:
                w.wait()
            self.reducer.allreduce_range(*self.lo_range)
            self.update_graph.replay()
        return self.losses[index]
