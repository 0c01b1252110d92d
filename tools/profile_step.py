"""torch.profiler breakdown of one training step of bench.py's workload (GPU box only): per-kernel device time,
launch counts and the wall-clock/GPU-busy ratio.  Usage: python tools/profile_step.py [--workload ucf] [--batch 8]"""
import argparse, contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

import bench
from devias_b200 import engine
from devias_b200.loss import TrainLoss
from devias_b200.modeling_slot import slot_vit_base_patch16_224

ap = argparse.ArgumentParser()
ap.add_argument('--workload', default='ucf'); ap.add_argument('--batch', type=int, default=0); ap.add_argument('--steps', type=int, default=3)
a = ap.parse_args()
cfg = dict(bench.WORKLOADS[a.workload]); B = a.batch or cfg['batch']; C = cfg['num_classes']
with contextlib.redirect_stdout(io.StringIO()):
    m = slot_vit_base_patch16_224(num_classes=C, drop_path_rate=cfg['drop_path_rate'], fc_drop_rate=cfg['fc_drop_rate'], init_scale=0.001,
                                  num_latents=cfg['num_latents'], slot_matching_method='matching', agg_weights_tie=cfg['agg_weights_tie'],
                                  agg_depth=cfg['agg_depth']).cuda().train()
crit = TrainLoss(None, 'KL', C); opt = torch.optim.AdamW(m.parameters(), lr=1e-4, fused=True)
clip = torch.randn(B, 3, 16, 224, 224, device='cuda'); tgt = torch.randint(0, C, (B,), device='cuda')
fg = (torch.rand(B, 196, device='cuda'), torch.rand(B, 1568, device='cuda')); teacher = torch.randn(B, 365, device='cuda')
step = lambda: engine.train_step(m, None, crit, opt, clip, tgt, fg, teacher_logits=teacher)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    e0.record()
    for _ in range(a.steps): step()
    e1.record(); torch.cuda.synchronize()
wall = e0.elapsed_time(e1) / a.steps
rows = {}
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        r = rows.setdefault(ev.name[:90], [0.0, 0]); r[0] += ev.device_time / 1e3 if hasattr(ev, 'device_time') else ev.cuda_time / 1e3; r[1] += 1
tot = sum(r[0] for r in rows.values()) / a.steps
print(f'wall ms/step {wall:.3f}   sum of kernel ms/step {tot:.3f}   clips/s {B / wall * 1e3:.1f}')
for k, (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:45]:
    print(f'{t / a.steps:9.3f} ms  {n // a.steps:5d}x  {k}')
