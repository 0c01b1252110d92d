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
