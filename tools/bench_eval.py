"""Inference sweep (BASELINE.json config 4): action+scene logits on synthetic 16x224^2 clips, batch 1..256, one B200.
Eval forward captured in a CUDA graph per batch size.  GPU box: python tools/bench_eval.py"""
import contextlib, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from devias_b200.modeling_slot import slot_vit_base_patch16_224
with contextlib.redirect_stdout(io.StringIO()):
    m = slot_vit_base_patch16_224(num_classes=101, num_latents=2, agg_depth=4, agg_weights_tie=True, slot_matching_method='matching',
                                  init_scale=1.0).cuda().eval()
sizes = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 128, 256]
for B in sizes:
    x = torch.randn(B, 3, 16, 224, 224, device='cuda')
    with torch.no_grad():
        for _ in range(3): out = m(x)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m(x)
        for _ in range(2): g.replay()
        torch.cuda.synchronize()
        n = max(3, min(50, 2000 // B))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): g.replay()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        e0.record()
        for _ in range(n): out = m(x)
        e1.record(); torch.cuda.synchronize()
        ems = e0.elapsed_time(e1) / n
    print(f'B={B:4d}  graph {ms:8.3f} ms  {B / ms * 1e3:8.1f} clips/s  ({B * 360.69 / ms:6.0f} TFLOP/s)   eager {ems:8.3f} ms')
    del x, g, out
    torch.cuda.empty_cache()
