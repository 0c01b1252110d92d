"""Device-resident DEVIAS training objective -- drop-in for utils/loss/train_loss.py:7-187 ('matching' branch).

The reference solves one tiny Hungarian problem per clip on the CPU (scipy, utils/loss/train_loss.py:112-122: B
device->host syncs) and accumulates O(B*S) scalar kernels.  With two label columns (action, scene) the assignment is
`argmin_{i != j} cost[i, action] + cost[j, scene]`, solved here for the whole batch at once on the GPU, and every term
is evaluated batched.  Values are identical to the reference's (tests/test_loss.py, golden from the reference)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class _FusedTrainLossFn(torch.autograd.Function):
    """csrc/train_loss.cu: the whole objective in one launch, its gradients in one more (include/devias_b200.h)"""

    @staticmethod
    def forward(ctx, head, attn, maskp, slots, target, teacher, fg, fgf, cfg):
        from . import _lib
        C, scene_ce, w_scene, w_mp, w_md = cfg
        B = target.shape[0]
        S = head.shape[0] // B
        H = attn.shape[0] // B
        head, attn, maskp, slots = head.contiguous(), attn.contiguous(), maskp.contiguous(), slots.contiguous()
        teacher, fg, fgf, target = teacher.contiguous(), fg.contiguous(), fgf.contiguous(), target.contiguous()
        var = (teacher.min() - 1.0).reshape(1)                                      # train_loss.py:103 (batch-wide scalar)
        out6 = torch.empty(6, device=head.device, dtype=torch.float32)
        idx = torch.empty(B, 2, device=head.device, dtype=torch.int64)
        dims = (B, S, head.shape[1], C, H, attn.shape[-1], maskp.shape[-1], slots.shape[-1], int(scene_ce))
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().devias_train_loss_fwd(head.data_ptr(), attn.data_ptr(), maskp.data_ptr(), slots.data_ptr(), target.data_ptr(),
                                                    teacher.data_ptr(), var.data_ptr(), fg.data_ptr(), fgf.data_ptr(), *dims,
                                                    float(w_scene), float(w_mp), float(w_md), out6.data_ptr(), idx.data_ptr(), s),
                   'train_loss_fwd')
        ctx.save_for_backward(head, attn, maskp, slots, target, teacher, var, fg, fgf)
        ctx.dims, ctx.weights = dims, (float(w_scene), float(w_mp), float(w_md))
        ctx.mark_non_differentiable(idx)
        return out6[5], out6[:5], idx

    @staticmethod
    def backward(ctx, gtotal, gparts, gidx):
        from . import _lib
        head, attn, maskp, slots, target, teacher, var, fg, fgf = ctx.saved_tensors
        g = gtotal.reshape(1).float().contiguous()
        dhead, dattn, dmaskp, dslots = (torch.empty_like(t) for t in (head, attn, maskp, slots))
        s = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().devias_train_loss_bwd(head.data_ptr(), attn.data_ptr(), maskp.data_ptr(), slots.data_ptr(), target.data_ptr(),
                                                    teacher.data_ptr(), var.data_ptr(), fg.data_ptr(), fgf.data_ptr(), *ctx.dims,
                                                    *ctx.weights, g.data_ptr(), dhead.data_ptr(), dattn.data_ptr(), dmaskp.data_ptr(),
                                                    dslots.data_ptr(), s), 'train_loss_bwd')
        return dhead, dattn, dmaskp, dslots, None, None, None, None, None


class TrainLoss(nn.Module):
    def __init__(self, criterion, scene_criterion, num_action_classes: int, slot_matching_method='matching',
                 scene_loss_weight=2000, mask_prediction_loss_weight=1, mask_distill_loss_weight=3, sync_items=False):
        super().__init__()
        self.criterion = criterion
        self.scene_criterion = scene_criterion
        self.num_action_classes = num_action_classes
        self.num_scene_classes = 365
        self.slot_matching_method = slot_matching_method
        self.mask_prediction_loss_weight = mask_prediction_loss_weight
        self.mask_distill_loss_weight = mask_distill_loss_weight
        self.scene_loss_weight = scene_loss_weight
        #: True reproduces the reference's five `.item()` host syncs per step (python floats in the dict)
        self.sync_items = sync_items
        #: CUDA inputs take the fused kernels (csrc/train_loss.cu); False evaluates the torch expressions below instead (tests)
        self.fused = True
        if slot_matching_method != 'matching':
            raise NotImplementedError("only the live 'matching' branch is provided (hard_select crashes in the reference, "
                                      "SURVEY.md R8)")

    @staticmethod
    def match(slots_head_softmax, target, scene_target):
        """batched utils/loss/train_loss.py:112-122: returns (action_slot [B], scene_slot [B]) int64"""
        B, S, _ = slots_head_softmax.shape
        ar = torch.arange(B, device=target.device)
        ca = -slots_head_softmax[ar, :, target]            # [B, S]
        cs = -slots_head_softmax[ar, :, scene_target]
        pair = ca.unsqueeze(2) + cs.unsqueeze(1)            # [B, i, j]
        pair = pair.masked_fill(torch.eye(S, dtype=torch.bool, device=pair.device), float('inf'))
        flat = pair.reshape(B, -1).argmin(dim=1)
        return flat // S, flat % S

    def forward(self, model, student_output, teacher_outputs, target, fg_mask=None):
        _, (action_output, _, attn), (slots_head, slots, mask_predictions) = student_output
        bs = target.shape[0]
        S = slots_head.shape[0] // bs
        if self.fused and slots_head.is_cuda and 2 <= S <= 8 and self.scene_criterion in ('KL', 'CE'):
            # one kernel launch for the objective, one for its gradients (csrc/train_loss.cu); the torch expressions below are the
            # same arithmetic and serve CPU tensors (unit tests against the reference's values)
            fg, fg_frames = fg_mask
            total, parts5, idx = _FusedTrainLossFn.apply(
                slots_head.float(), attn.float(), mask_predictions.float().reshape(bs * S, -1), slots.float().reshape(bs * S, -1),
                target, teacher_outputs[1].float(), fg.float(), fg_frames.float(),
                (self.num_action_classes, self.scene_criterion == 'CE', self.scene_loss_weight, self.mask_prediction_loss_weight,
                 self.mask_distill_loss_weight))
            act_rows = slots_head.view(bs, S, -1)[torch.arange(bs, device=target.device), idx[:, 0]]
            names = ('action_loss', 'scene_loss', 'cosine_loss', 'mask_prediction_loss', 'mask_distill_loss')
            parts5 = parts5.detach()
            parts = {k: (parts5[i].item() if self.sync_items else parts5[i]) for i, k in enumerate(names)}
            return total, act_rows, parts
        H = attn.size(0) // bs
        C = self.num_action_classes
        slots_head = slots_head.float()
        attn = attn.float().reshape(bs, H, S, -1).mean(dim=1)                        # :97
        mask_predictions = mask_predictions.float().reshape(bs, S, -1)
        _, teacher_scene_logit = teacher_outputs
        teacher_scene_logit = teacher_scene_logit.float()
        scene_target = torch.argmax(teacher_scene_logit, dim=1) + C                  # :101,:107
        var = teacher_scene_logit.min() - 1.0                                        # :103
        teacher_full = torch.cat([var.expand(bs, C), teacher_scene_logit], dim=1)    # :104-106
        head3 = slots_head.view(bs, S, -1)
        with torch.no_grad():
            ai, si = self.match(head3.softmax(-1), target, scene_target)
        ar = torch.arange(bs, device=target.device)
        fg, fg_frames = fg_mask
        act_rows = head3[ar, ai]                                                     # [B, C+365]
        action_loss = F.cross_entropy(act_rows, target, reduction='sum') / bs        # :150,:168
        if self.scene_criterion == 'CE':
            scene_loss = F.cross_entropy(head3[ar, si], scene_target, reduction='sum') / bs
        else:  # 'KL' (:160-165): F.kl_div(..., 'batchmean') on a 1-D row divides by the number of classes
            kl = F.kl_div(F.log_softmax(head3[ar, si], dim=-1), F.log_softmax(teacher_full, dim=-1),
                          reduction='none', log_target=True).sum(-1) / head3.shape[-1]
            scene_loss = kl.sum() * self.scene_loss_weight / bs
        md = (attn[ar, ai] - fg_frames.float()).square().mean(-1).sum() * self.mask_distill_loss_weight / bs      # :145
        mp = F.binary_cross_entropy_with_logits(mask_predictions[ar, ai], fg.float(), reduction='none').mean(-1).sum() \
            * self.mask_prediction_loss_weight / bs                                                               # :146-149
        sl = F.normalize(slots.float().reshape(bs, S, -1), p=2, dim=2)                                           # :173-178
        cs = torch.bmm(sl, sl.transpose(1, 2)) * (1 - torch.eye(S, device=sl.device))
        cosine_loss = (cs.sum(dim=(1, 2)) / (S * (S - 1))).mean()
        total = action_loss + scene_loss + cosine_loss + mp + md
        parts = {'action_loss': action_loss, 'scene_loss': scene_loss, 'cosine_loss': cosine_loss,
                 'mask_prediction_loss': mp, 'mask_distill_loss': md}
        parts = {k: (v.item() if self.sync_items else v.detach()) for k, v in parts.items()}
        return total, act_rows, parts
