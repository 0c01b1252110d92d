"""FAME foreground/background augmentation of the DEVIAS training loop on the device, without kornia
(reference: utils/transform/fame.py:13-153, called once per step from engine/engine_for_slot.py before the student forward;
SURVEY.md section 8f row N4).  Same constructor and `forward(videos, label, center_frame=None)` contract:

    returns (videos', labels', (fg_mask [B, 196], fg_mask_per_frame [B, 8*196])[, center_frame'])

The two kornia calls of the reference are restated from kornia's published definitions.  kornia is not in this image, so parity
is pinned on hand-computed known answers of those definitions (tests/golden/fame_known_answers.json: kornia's docstring Gaussian
kernels, the 'reflect' border, the HSV sextants) and on an independent numpy restatement in oracle/ -- not on kornia's outputs:
  * kornia.filters.GaussianBlur2d((k, k), (k/3, k/3)): separable normalised Gaussian, 'reflect' border;
  * kornia.color.rgb_to_hsv: h in [0, 2 pi), s = delta / (max + 1e-8), v = max.
Random draws follow the reference call for call (randperm on the video's device, rand on the CPU), so a seeded CPU run
reproduces the reference's sample selection.  Everything is torch on the clip's device: no host round trip."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def gaussian_kernel1d(ksize: int, sigma: float, device=None, dtype=torch.float32):
    x = torch.arange(ksize, device=device, dtype=dtype) - ksize // 2
    if ksize % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_blur2d(x: torch.Tensor, ksize: int, sigma: float):
    """x [N, 1, H, W]; separable Gaussian with reflect padding"""
    k = gaussian_kernel1d(ksize, sigma, x.device, x.dtype)
    p = ksize // 2
    x = F.pad(x, (p, p, p, p), mode='reflect')
    x = F.conv2d(x, k.view(1, 1, 1, ksize))
    return F.conv2d(x, k.view(1, 1, ksize, 1))


def rgb_to_hsv(image: torch.Tensor, eps: float = 1e-8):
    """image [N, 3, H, W] in [0, 1] -> hsv with h in [0, 2 pi)"""
    max_rgb, argmax_rgb = image.max(-3)
    min_rgb = image.min(-3)[0]
    deltac = max_rgb - min_rgb
    v = max_rgb
    s = deltac / (max_rgb + eps)
    deltac = torch.where(deltac == 0, torch.ones_like(deltac), deltac)
    rc, gc, bc = torch.unbind(max_rgb.unsqueeze(-3) - image, dim=-3)
    h = torch.stack((bc - gc, (rc - bc) + 2.0 * deltac, (gc - rc) + 4.0 * deltac), dim=-3) / deltac.unsqueeze(-3)
    h = torch.gather(h, dim=-3, index=argmax_rgb.unsqueeze(-3)).squeeze(-3)
    h = 2.0 * math.pi * ((h / 6.0) % 1.0)
    return torch.stack((h, s, v), dim=-3)


class FAME(nn.Module):
    def __init__(self, crop_size=112, beta=0.5, device="cpu", eps=1e-8, prob_aug=0.5):
        super().__init__()
        self.frame_mean = [0.485, 0.456, 0.406]
        self.frame_std = [0.229, 0.224, 0.225]
        self.crop_size = crop_size
        self.gauss_size = int(0.1 * crop_size) // 2 * 2 + 1
        self.gauss_sigma = self.gauss_size / 3
        self.device = device
        self.eps = eps
        self.beta = beta          # portion of foreground
        self.prob_aug = prob_aug

    def gauss(self, x):
        return gaussian_blur2d(x, self.gauss_size, self.gauss_sigma)

    def norm_batch(self, matrix):
        """min-max normalisation per sample (utils/transform/fame.py:29-35)"""
        B, H, W = matrix.shape
        matrix = matrix.flatten(start_dim=1)
        matrix = matrix - matrix.min(dim=-1, keepdim=True)[0]
        matrix = matrix / (matrix.max(dim=-1, keepdim=True)[0] + self.eps)
        return matrix.reshape(B, H, W)

    @staticmethod
    def batched_bincount(x, max_value):
        target = torch.zeros(x.shape[0], max_value, dtype=x.dtype, device=x.device)
        target.scatter_add_(1, x, torch.ones_like(x))
        return target

    def getSeg(self, mask, video_clips):
        """colour-histogram refinement of the motion mask (utils/transform/fame.py:43-85)"""
        B, C, T, H, W = video_clips.shape
        img_hsv = rgb_to_hsv(video_clips.mean(dim=2).reshape(-1, C, H, W))
        flat = mask.reshape(B, -1)
        fg_index = torch.topk(flat, k=int(0.5 * H * W), dim=-1)[1]
        bg_index = torch.topk(flat, k=int(0.1 * H * W), dim=-1, largest=False)[1]
        dimH, dimS, dimV = 10, 10, 10
        img_h, img_s, img_v = img_hsv[:, 0], img_hsv[:, 1], img_hsv[:, 2]
        hx = (img_s * torch.cos(img_h * 2 * math.pi) + 1) / 2      # (the reference multiplies the radian hue by 2 pi again)
        hy = (img_s * torch.sin(img_h * 2 * math.pi) + 1) / 2
        h = torch.round(hx * (dimH - 1) + 1)
        s = torch.round(hy * (dimS - 1) + 1)
        v = torch.round(img_v * (dimV - 1) + 1)
        color_map = (h + (s - 1) * dimH + (v - 1) * dimH * dimS).reshape(B, -1).long()
        nbin = dimH * dimS * dimV
        dict_fg = self.batched_bincount(color_map.gather(index=fg_index, dim=-1), nbin).float()
        dict_bg = self.batched_bincount(color_map.gather(index=bg_index, dim=-1), nbin).float() + 1
        dict_fg = dict_fg / (dict_fg.sum(dim=-1, keepdim=True) + self.eps)
        dict_bg = dict_bg / (dict_bg.sum(dim=-1, keepdim=True) + self.eps)
        pr_fg = dict_fg.gather(dim=1, index=color_map)
        pr_bg = dict_bg.gather(dim=1, index=color_map)
        refine = pr_fg / (pr_bg + pr_fg)
        m = self.norm_batch(self.gauss(refine.reshape(-1, 1, H, W)).reshape(-1, H, W))
        num_fg = int(self.beta * H * W)
        sampled = torch.topk(m.reshape(B, -1), k=num_fg, dim=-1)[1]
        out = torch.zeros(B, H * W, dtype=m.dtype, device=m.device)
        out.scatter_(1, sampled, 1.0)
        return out.reshape(B, H, W)

    def _mask_from_diff(self, im_diff, video_clips):
        H, W = im_diff.shape[-2:]
        m = self.gauss(im_diff.reshape(-1, 1, H, W))
        m = self.norm_batch(m.reshape(-1, H, W))
        return self.getSeg(m, video_clips)

    def getmask(self, video_clips):
        """clip-level mask from the mean absolute frame difference (utils/transform/fame.py:87-96)"""
        im_diff = (video_clips[:, :, 0:-1] - video_clips[:, :, 1:]).abs().sum(dim=1).mean(dim=1)
        return self._mask_from_diff(im_diff, video_clips)

    def getmask_per_frame(self, video_clips):
        """one mask per tube (frame pair) (utils/transform/fame.py:98-110)"""
        T = video_clips.shape[2]
        return [self._mask_from_diff((video_clips[:, :, i] - video_clips[:, :, i + 1]).abs().sum(dim=1), video_clips)
                for i in range(0, T, 2)]

    @torch.no_grad()
    def forward(self, videos, label, center_frame=None):
        batch_size = videos.shape[0]
        std = torch.tensor(self.frame_std, device=videos.device, dtype=videos.dtype).reshape(1, 3, 1, 1, 1)
        mean = torch.tensor(self.frame_mean, device=videos.device, dtype=videos.dtype).reshape(1, 3, 1, 1, 1)
        tmp = (videos.contiguous() * std + mean).float()                       # de-normalised
        mask = self.getmask(tmp)
        masks_per_frame = torch.stack(self.getmask_per_frame(tmp)).permute(1, 0, 2, 3)
        mask = mask.to(videos.dtype).unsqueeze(1).unsqueeze(1)
        masks_per_frame = masks_per_frame.to(videos.dtype)
        index = torch.randperm(batch_size, device=videos.device)
        video_fuse = videos[index] * (1 - mask) + videos * mask
        all_center_frame = center_frame
        if self.prob_aug < 1:
            rand_batch = torch.rand(batch_size)
            aug_ind = torch.where(rand_batch < self.prob_aug)[0].to(videos.device)
            ori_ind = torch.where(rand_batch >= self.prob_aug)[0].to(videos.device)
            all_videos = torch.cat([video_fuse[aug_ind], videos[ori_ind]], dim=0).contiguous()
            li = lambda t, i: t[i.to(t.device)]
            all_label = torch.cat([li(label, aug_ind), li(label, ori_ind)], dim=0).contiguous()
            if center_frame is not None:
                all_center_frame = torch.cat([li(center_frame, aug_ind), li(center_frame, ori_ind)], dim=0).contiguous()
            mask = torch.cat([mask[aug_ind], mask[ori_ind]], dim=0).contiguous()
            masks_per_frame = torch.cat([masks_per_frame[aug_ind], masks_per_frame[ori_ind]], dim=0).contiguous()
        else:
            all_videos, all_label = video_fuse, label
        pooled = F.avg_pool2d(mask.squeeze(1).squeeze(1), kernel_size=16, stride=16).view(batch_size, -1)
        pooled_pf = F.avg_pool2d(masks_per_frame, kernel_size=16, stride=16).reshape(batch_size, -1)
        masks = (pooled.to(label.device, non_blocking=True), pooled_pf.to(label.device, non_blocking=True))
        if center_frame is not None:
            return all_videos, all_label, masks, all_center_frame
        return all_videos, all_label, masks
