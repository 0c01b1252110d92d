#!/usr/bin/env python
"""bench.py -- DEVIAS hot-path benchmark (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ucf|k400]

metric  : train clips/sec, 16x224^2 clips, student fwd + loss + bwd + AdamW step (BASELINE.json config[1]:
          UCF-101-shaped: 101 action / 365 scene classes, 2 slots, weight-tied agg depth 4, batch 8 per GPU, bf16).
value   : whole-job clips/s with the step's inputs already resident in HBM.
e2e     : same step driven through the public API with HOST (pinned) buffers: H2D of clip/labels/masks and D2H of the
          loss inside the timed region (copies double-buffered on a side stream, as a DataLoader with pin_memory does).
roofline: the dominant kernel family (the tcgen05 GEMM) timed live with CUDA events on its launch stream during the
          timed region: achieved TFLOP/s over algorithmic FLOPs vs the measured cuBLAS bf16 peak.
cpu_baseline / --impl reference: the CPU oracle port of the reference's PyTorch path (oracle/devias_oracle.py) on the
          host cores, one clip per step (bounded sample).
"""
import argparse
import contextlib
import io
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'train clips/sec (16x224^2, fwd+bwd)'
WORKLOADS = {
    # docs/TRAIN.md:73-128 (UCF-101 recipe) and :12-63 (K400 recipe)
    'ucf': dict(num_classes=101, num_latents=2, agg_depth=4, agg_weights_tie=True, batch=8, drop_path_rate=0.2, fc_drop_rate=0.5),
    'k400': dict(num_classes=400, num_latents=2, agg_depth=8, agg_weights_tie=True, batch=32, drop_path_rate=0.1, fc_drop_rate=0.0),
}
FWD_GFLOP_PER_CLIP = 360.69          # SURVEY.md section 8d (encoder, algorithmic)
TRAIN_GFLOP_PER_CLIP = 1082.07 - 3.7  # fwd + bwd, patch-embed dgrad not needed


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='ucf', choices=list(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='clips per GPU (default: the workload recipe)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--split-block', type=int, default=3, help='data-parallel step: encoder block at which the backward graph is cut')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of the captured CUDA graph')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1346.6), d.get('hbm_gbs', 6549.4), 'measured'
    return 1400.0, 6650.0, 'fallback'


class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.path = f'/tmp/devias_clocks_{os.getpid()}.csv'
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.f = open(self.path, 'w')
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.gpu), '-lms', '100'], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            c = [t.strip() for t in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_reference_step_fn(cfg):
    """the reference's CPU path (oracle port): fwd + TrainLoss + bwd on ONE clip, all host threads"""
    import numpy as np
    import torch
    from oracle import devias_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    C = cfg['num_classes']
    sd = O.synth_state_dict(num_classes=C, num_latents=cfg['num_latents'], agg_depth=cfg['agg_depth'],
                            agg_weights_tie=cfg['agg_weights_tie'], seed=0)
    uniq = {}
    for v in sd.values():
        uniq.setdefault(id(v), v.requires_grad_(True))
    clip = O.synth_clips(1, seed=0)
    rs = np.random.RandomState(0)
    target = torch.from_numpy(rs.randint(0, C, size=(1,)).astype(np.int64))
    teacher = torch.from_numpy(rs.standard_normal(size=(1, 365)).astype(np.float32))
    fg = (torch.from_numpy(rs.uniform(size=(1, 196)).astype(np.float32)), torch.from_numpy(rs.uniform(size=(1, 1568)).astype(np.float32)))

    def step():
        for v in uniq.values():
            v.grad = None
        out = O.student_forward(sd, clip, C)
        total, _, _ = O.train_loss(out, teacher, target, fg, C)
        total.backward()
        return float(total)
    return step, torch.get_num_threads()


def run_reference(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    step, cores = cpu_reference_step_fn(cfg)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = args.steps / dt
    sample = '1 clip per step (fwd + TrainLoss + bwd, fp32), oracle port of the reference PyTorch path'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'clips/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload}: DEVIAS ViT-B/16 (1568 tube tokens) + {cfg["num_latents"]}-slot aggregation, '
                               f'tied={cfg["agg_weights_tie"]} depth {cfg["agg_depth"]}, {cfg["num_classes"]}+365 classes, train step '
                               f'(fwd + TrainLoss + bwd), reference CPU path (oracle port), bounded sample of 1 clip per step',
                   'clips_per_step': 1, 'parallelism': 'cpu'},
        'cpu_baseline': {'value': v, 'unit': 'clips/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'clips/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def main():
    args = parse()
    cfg = dict(WORKLOADS[args.workload])
    if args.batch:
        cfg['batch'] = args.batch
    if args.impl == 'reference':
        return run_reference(args, cfg)

    import numpy as np
    import torch
    import torch.distributed as dist
    from devias_b200 import _lib, engine
    from devias_b200.ddp import GradReducer
    from devias_b200.loss import TrainLoss
    from devias_b200.modeling_slot import slot_vit_base_patch16_224

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, C = cfg['batch'], cfg['num_classes']

    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = slot_vit_base_patch16_224(num_classes=C, all_frames=16, tubelet_size=2, drop_path_rate=cfg['drop_path_rate'],
                                          fc_drop_rate=cfg['fc_drop_rate'], init_scale=0.001, num_latents=cfg['num_latents'],
                                          head_type='linear', slot_matching_method='matching',
                                          agg_weights_tie=cfg['agg_weights_tie'], agg_depth=cfg['agg_depth'],
                                          num_scene_classes=365)
    model = model.to(dev).train()
    crit = TrainLoss(torch.nn.CrossEntropyLoss(), 'KL', C)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.05, fused=True, capturable=True)
    reducer = GradReducer(model) if world > 1 else None

    # synthetic step inputs: N(0,1) clips, random labels, FAME-like masks, teacher logits
    nbuf = 2
    rs = np.random.RandomState(100 + rank)
    host = []
    for i in range(nbuf):
        host.append(dict(
            clip=torch.from_numpy(rs.standard_normal(size=(B, 3, 16, 224, 224)).astype(np.float32)).pin_memory(),
            target=torch.from_numpy(rs.randint(0, C, size=(B,)).astype(np.int64)).pin_memory(),
            fg=torch.from_numpy((rs.uniform(size=(B, 196)) > 0.5).astype(np.float32)).pin_memory(),
            fgf=torch.from_numpy((rs.uniform(size=(B, 1568)) > 0.5).astype(np.float32)).pin_memory(),
            teacher=torch.from_numpy(rs.standard_normal(size=(B, 365)).astype(np.float32)).pin_memory()))
    devb = [{k: v.to(dev) for k, v in h.items()} for h in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def step(b):
        loss, _, _ = engine.train_step(model, None, crit, opt, b['clip'], b['target'], (b['fg'], b['fgf']),
                                       teacher_logits=b['teacher'], reducer=reducer)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- warm-up (+ graph capture of the whole step)
    use_graph = not args.no_graph
    for i in range(max(args.warmup, 3)):
        step(devb[i % nbuf])
    barrier()
    graphed = None
    if use_graph:
        graphed = engine.GraphedTrainStep(model, crit, opt, devb, reducer=reducer, warmup=1, split_block=args.split_block)
        for i in range(2):
            graphed(i % nbuf)
        barrier()

    def run(i):
        return graphed(i % nbuf) if graphed is not None else step(devb[i % nbuf])

    # ---------------------------------------------------------------- value: inputs resident in HBM
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    n0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        loss = run(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = (graphed.launches_per_step * args.steps) if graphed is not None else (_lib.launch_count() - n0)
    # roofline leg: the same steps launched eagerly with every GEMM bracketed by CUDA events on its stream
    # (events cannot be timed inside a replayed graph); same kernels, same shapes, same data
    _lib.profile_begin()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for i in range(args.steps):
        step(devb[i % nbuf])
    ev3.record()
    barrier()
    eager_ms = ev2.elapsed_time(ev3)
    gemm_ms, gemm_flops, gemm_n = _lib.profile_end(0)
    attn_ms, attn_flops, attn_n = _lib.profile_end(1)
    slot_ms, slot_bytes, slot_n = _lib.profile_end(2)
    clk = clocks.stop() if rank == 0 else None
    loss_val = float(loss)
    assert loss_val == loss_val, 'loss is NaN'

    # ---------------------------------------------------------------- e2e: host buffers, H2D + D2H in the timed region
    e2e_ms = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        main_stream = torch.cuda.current_stream()
        ready = [torch.cuda.Event() for _ in range(nbuf)]
        done = [torch.cuda.Event() for _ in range(nbuf)]
        loss_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()

        def h2d(i):
            s = i % nbuf
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[s])
                for k, v in host[s].items():
                    devb[s][k].copy_(v, non_blocking=True)
                ready[s].record(copy_stream)

        for s in range(nbuf):
            done[s].record(main_stream)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h2d(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                h2d(i + 1)
            main_stream.wait_event(ready[i % nbuf])
            l = run(i)
            done[i % nbuf].record(main_stream)
            loss_host[i:i + 1].copy_(l.reshape(1), non_blocking=True)
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1)

    # ---------------------------------------------------------------- slot-attention micro-measure (BASELINE config 5)
    slot_micro = None
    if rank == 0:
        try:
            from devias_b200 import ops as _ops
            Bm, Sm = 256, cfg['num_latents']     # 1.23 GB of fp32 tokens: far beyond the 126 MB L2, every pass streams from HBM
            tok = torch.randn(Bm, 1568, 768, device=dev) * 1.5
            g_ = torch.randn(Bm, 4 * Sm, 768, device=dev) * 0.05
            G_ = g_.sum(-1).contiguous(); c0_ = torch.randn(Bm, 4 * Sm, device=dev)

            def _time(fn, n=10):
                for _ in range(3):
                    fn()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); s0.record()
                for _ in range(n):
                    fn()
                s1.record(); torch.cuda.synchronize()
                return s0.elapsed_time(s1) / n
            t_ms = _time(lambda: _ops.slot_stream_fwd(tok, g_, G_, c0_))
            nbytes = Bm * 1568 * 768 * 4
            slot_micro = {'batch': Bm, 'slots': Sm, 'tokens_dtype': 'f32', 'us_per_pass': t_ms * 1e3,
                          'gbs': nbytes / (t_ms * 1e-3) / 1e9}
            if Sm in (2, 4):                      # streaming backward: reads the tokens, writes their gradient
                U_, m_, A_, at_, mu_, r_ = _ops.slot_stream_fwd(tok, g_, G_, c0_)
                dU_, dm_, dA_ = torch.randn_like(U_), torch.randn_like(m_), torch.randn_like(A_)
                tb_ms = _time(lambda: _ops.slot_stream_bwd(tok, mu_, r_, g_, G_, at_, dU_, dm_, dA_))
                slot_micro.update({'bwd_us_per_pass': tb_ms * 1e3, 'bwd_gbs': 2 * nbytes / (tb_ms * 1e-3) / 1e9})
                del U_, m_, A_, at_, mu_, r_, dU_, dm_, dA_
            del tok, g_, G_, c0_
        except Exception as e:  # the micro-measure must never take the headline down
            slot_micro = {'error': repr(e)}

    # ---------------------------------------------------------------- max over ranks
    t = torch.tensor([ms, e2e_ms or 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        peak_tf, peak_gbs, peak_src = peaks()
        value = world * B * args.steps / (ms * 1e-3)
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, 'profiles', 'gemm_traffic.json')
        if os.path.isfile(tp):
            try:
                traffic = json.load(open(tp)).get('dram_bytes_per_launch')
            except Exception:
                traffic = None
        out = {
            'metric': METRIC, 'value': value, 'unit': 'clips/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': f'{args.workload}: DEVIAS ViT-B/16 (1568 tube tokens) + {cfg["num_latents"]}-slot aggregation, '
                                   f'tied={cfg["agg_weights_tie"]} depth {cfg["agg_depth"]}, {C}+365 classes, train step '
                                   f'(fwd + TrainLoss + bwd + fused AdamW), drop_path {cfg["drop_path_rate"]}',
                       'clips_per_gpu': B, 'global_batch': B * world, 'parallelism': f'dp{world}',
                       'l2': 'per-step working set (activations + weights, several GB) far exceeds the 126 MB L2; no flush needed',
                       'loss': loss_val, 'launch_mode': ('eager' if graphed is None else 'cuda-graph replay of the whole step' if world == 1 else
                                       'three cuda graphs per step (fwd+bwd down to block 3 | bwd of blocks 0-2 | AdamW); the NCCL all-reduce of the upper '
                                       'gradients (flat fp32 arena) runs between them, overlapping the lower backward')},
            'clocks': clk,
            'gpu_launches': int(launches),
            'roofline': {'kernel': 'gemm_bf16_kernel (tcgen05/TMEM/TMA)', 'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf,
                         'unit': 'TFLOP/s', 'frac': achieved / peak_tf, 'traffic': traffic, 'peak_source': f'{peak_src} (sustained cuBLAS bf16)',
                         'launches': int(gemm_n), 'share_of_step': gemm_ms / eager_ms, 'timed_over': 'eager re-run of the timed steps',
                         'eager_ms_per_step': eager_ms / args.steps,
                         'attention': {'tflops': attn_flops / (attn_ms * 1e-3) / 1e12 if attn_ms > 0 else None, 'share_of_step': attn_ms / eager_ms, 'launches': int(attn_n)},
                         'slot_attention': {'in_step_gbs': slot_bytes / (slot_ms * 1e-3) / 1e9 if slot_ms > 0 else None, 'peak_gbs': peak_gbs,
                                            'share_of_step': slot_ms / eager_ms, 'launches': int(slot_n), 'microbench': slot_micro,
                                            'microbench_frac_of_hbm': (slot_micro['gbs'] / peak_gbs) if slot_micro and 'gbs' in slot_micro else None},
                         'step_tensor_frac': value / world * TRAIN_GFLOP_PER_CLIP / 1e3 / peak_tf},
        }
        if e2e_ms:
            out['e2e'] = {'value': world * B * args.steps / (e2e_ms * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': h2d_bytes,
                          'd2h_bytes_per_step': 4}
        if world == 1 and not args.no_cpu_baseline:
            stepf, cores = cpu_reference_step_fn(cfg)
            stepf()
            best = 1e30
            for _ in range(2):
                t0 = time.perf_counter(); stepf(); best = min(best, time.perf_counter() - t0)
            out['cpu_baseline'] = {'value': 1.0 / best, 'unit': 'clips/s', 'cores': cores, 'kind': 'port',
                                   'sample': '1 clip fwd + TrainLoss + bwd (fp32), best of 2 after 1 warm-up'}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
