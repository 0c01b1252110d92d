#!/usr/bin/env python
"""bench.py -- DEVIAS hot-path benchmark (contract: see the task statement / DESIGN.md "Measurement").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload k400|ucf]

metric  : train clips/sec, 16x224^2 clips, student fwd + TrainLoss + bwd + AdamW step.  Default workload = the configuration
          BASELINE.json quotes the metric on "at 1/2/4/8 B200" (configs[2]): K400-shaped, 400 action / 365 scene classes, 2
          slots, weight-tied aggregation depth 8, batch 32 per GPU, bf16.  The UCF-101 recipe (configs[1]: 101 classes,
          aggregation depth 4, batch 8 per GPU) is measured in the same run and reported under `secondary`.
value   : whole-job clips/s with the step's inputs already resident in HBM (CUDA-graph replay of the captured step).
e2e     : same step driven through the public API with HOST (pinned) buffers: H2D of clip/labels/masks and D2H of the
          loss inside the timed region (copies double-buffered on a side stream, as a DataLoader with pin_memory does).
roofline: the dominant kernel family (the tcgen05 GEMM) timed INSIDE the replayed step: a second capture of the same step
          carries external event-record nodes around every GEMM / attention / slot-stream launch; algorithmic FLOPs over
          the summed event times vs the measured sustained cuBLAS bf16 peak.
parity  : before timing, the benchmark model (loaded with the seeded fixture weights) must reproduce the REFERENCE's logits
          of tests/golden (<= 1e-2 relative, top-1 agreement) on the first clips of the fixture batch.
gpu_reference: the reference's PyTorch path (oracle port: eager ATen/cuBLAS, materialised attention) on the SAME GPU, fp32 and
          bf16 autocast, bounded sample -- the practical bar next to the CPU baseline.
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
    'ucf': dict(golden='model_d12_ucf_b8', num_classes=101, num_latents=2, agg_depth=4, agg_weights_tie=True, batch=8, drop_path_rate=0.2, fc_drop_rate=0.5),
    'k400': dict(golden='model_d12_k400_b32', num_classes=400, num_latents=2, agg_depth=8, agg_weights_tie=True, batch=32, drop_path_rate=0.1, fc_drop_rate=0.0),
}
FWD_GFLOP_PER_CLIP = 360.69          # SURVEY.md section 8d (encoder, algorithmic)
TRAIN_GFLOP_PER_CLIP = 1082.07 - 3.7  # fwd + bwd, patch-embed dgrad not needed


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='k400', choices=list(WORKLOADS))
    ap.add_argument('--batch', type=int, default=0, help='clips per GPU (default: the workload recipe)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--cuts', default='1', help="data-parallel step: encoder blocks at which the backward graph is cut ('none': one piece)")
    ap.add_argument('--grad-exchange', default='bf16', choices=['fp32', 'bf16'], help='data-parallel gradient all-reduce precision')
    ap.add_argument('--no-exchange', action='store_true', help='diagnostic: N independent replicas, no gradient all-reduce (isolates what the collective costs)')
    ap.add_argument('--no-secondary', action='store_true', help='skip the UCF-101 B=8 secondary line')
    ap.add_argument('--no-extras', action='store_true', help='skip slot grid / eval sweep / gpu reference (N=1 extras)')
    ap.add_argument('--torch-adamw', action='store_true', help='torch fused AdamW instead of the arena optimizer pass')
    ap.add_argument('--no-graph', action='store_true', help='eager launches instead of the captured CUDA graph')
    ap.add_argument('--token-dtype', default='f32', choices=['f32', 'bf16'], help='dtype of the encoder tokens handed to the slot block (bf16: the tcgen05 slot kernels)')
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


def workload_text(name, cfg, C):
    return (f'{name}: DEVIAS ViT-B/16 (1568 tube tokens) + {cfg["num_latents"]}-slot aggregation, tied={cfg["agg_weights_tie"]} '
            f'depth {cfg["agg_depth"]}, {C}+365 classes, train step (fwd + TrainLoss + bwd + AdamW), drop_path {cfg["drop_path_rate"]}')


def parity_gate(model, cfg, dev, n=2):
    """the benchmark model with the fixture weights vs the REFERENCE's outputs (tests/golden, oracle/make_golden.py)"""
    import numpy as np
    import torch
    from oracle import devias_oracle as O
    g = np.load(os.path.join(ROOT, 'tests', 'golden', cfg['golden'] + '.npz'))
    C = cfg['num_classes']
    was_training = model.training
    model.eval()
    with torch.no_grad():
        _, (al, sl, _), _ = model(O.synth_clips(n, seed=13).to(dev))
    model.train(was_training)
    ref_a, ref_s = torch.from_numpy(g['action_logit'][:n]).to(dev), torch.from_numpy(g['scene_logit'][:n]).to(dev)
    err = max(float((al - ref_a).abs().max() / ref_a.abs().max()), float((sl - ref_s).abs().max() / ref_s.abs().max()))
    top1 = bool((al[:, :C].argmax(-1) == ref_a[:, :C].argmax(-1)).all() and (sl[:, C:].argmax(-1) == ref_s[:, C:].argmax(-1)).all()
                and (al.argmax(-1) == ref_a.argmax(-1)).all())
    assert err <= 1e-2, f'parity gate: logits differ from the reference golden by {err:.3e} (> 1e-2)'
    assert top1, 'parity gate: top-1 disagreement with the reference golden'
    return {'fixture': cfg['golden'], 'clips': n, 'logits_rel_max_err': err, 'tolerance': 1e-2, 'top1_agree': top1}


def measure_train(name, cfg, args, dev, world, rank, with_roofline=True, with_e2e=True):
    """captures and times the training step of one workload; returns a dict of raw measurements (this rank)"""
    import numpy as np
    import torch
    import torch.distributed as dist
    from devias_b200 import _lib, engine
    from devias_b200.arena import ParamArena
    from devias_b200.ddp import GradReducer
    from devias_b200.loss import TrainLoss
    from devias_b200.modeling_slot import slot_vit_base_patch16_224
    from devias_b200.optim import ArenaAdamW
    from oracle import devias_oracle as O     # fixture WEIGHTS + golden outputs for the parity gate (the checker, not the product)

    B, C = cfg['batch'], cfg['num_classes']
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = slot_vit_base_patch16_224(num_classes=C, all_frames=16, tubelet_size=2, drop_path_rate=cfg['drop_path_rate'],
                                          fc_drop_rate=cfg['fc_drop_rate'], init_scale=0.001, num_latents=cfg['num_latents'],
                                          head_type='linear', slot_matching_method='matching',
                                          agg_weights_tie=cfg['agg_weights_tie'], agg_depth=cfg['agg_depth'],
                                          num_scene_classes=365)
    model.load_state_dict(O.synth_state_dict(num_classes=C, num_latents=cfg['num_latents'], agg_depth=cfg['agg_depth'],
                                             agg_weights_tie=cfg['agg_weights_tie'], depth=12, seed=3))
    model = model.to(dev).train()
    if args.token_dtype == 'bf16':
        model.token_dtype = torch.bfloat16
    gate = parity_gate(model, cfg, dev)
    crit = TrainLoss(torch.nn.CrossEntropyLoss(), 'KL', C)
    decay = [p for n_, p in model.named_parameters() if p.dim() > 1 and n_ not in model.no_weight_decay()]
    rest = [p for n_, p in model.named_parameters() if not (p.dim() > 1 and n_ not in model.no_weight_decay())]
    groups = [dict(params=decay, weight_decay=0.05), dict(params=rest, weight_decay=0.0)]
    reducer = GradReducer(model) if (world > 1 and not args.no_exchange) else None
    if args.torch_adamw:
        opt = torch.optim.AdamW(groups, lr=1e-4, fused=True, capturable=True)
    else:
        opt = ArenaAdamW(groups, ParamArena.of(model), lr=1e-4, betas=(0.9, 0.999), eps=1e-8)

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

    cuts = [int(c) for c in args.cuts.split(',') if c.strip().isdigit()]
    # ---------------------------------------------------------------- warm-up (+ graph capture of the whole step)
    use_graph = not args.no_graph
    # (DEVIAS_BENCH_LAUNCH_LIST=1: the ncu launch-list run of tools/evidence.sh -- eager, --warmup steps exactly, no roofline leg;
    #  every extra step costs minutes under the profiler.  Its printed numbers are not bench values.)
    launch_list = bool(os.environ.get('DEVIAS_BENCH_LAUNCH_LIST')) and not use_graph
    for i in range((args.warmup if launch_list else max(args.warmup, 3)) if not use_graph else 1):
        step(devb[i % nbuf])
    barrier()
    graphed = None
    if use_graph:
        graphed = engine.GraphedTrainStep(model, crit, opt, devb, reducer=reducer, warmup=1, cuts=cuts, grad_exchange=args.grad_exchange)
        for i in range(max(args.warmup, 3)):
            graphed(i % nbuf)
        barrier()

    def run(i):
        return graphed(i % nbuf) if graphed is not None else step(devb[i % nbuf])

    # ---------------------------------------------------------------- value: inputs resident in HBM
    clocks = ClockSampler(dev.index)
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
    clk = clocks.stop() if rank == 0 else None
    launches = (graphed.launches_per_step * args.steps) if graphed is not None else (_lib.launch_count() - n0)
    loss_val = float(loss)
    assert loss_val == loss_val, 'loss is NaN'

    # ---------------------------------------------------------------- e2e: host buffers, H2D + D2H in the timed region
    e2e_ms = None
    if with_e2e and not args.no_e2e:
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

    # ---------------------------------------------------------------- roofline leg: kernels timed INSIDE the replayed step
    # a second capture of the same step whose GEMM / attention / slot-stream launches are bracketed by external event-record
    # nodes; every replay re-records them (runtime.cu), so the times are those of kernels running back to back in the graph
    prof = None
    if with_roofline and use_graph:
        _lib.profile_begin(capture_only=True)
        pstep = engine.GraphedTrainStep(model, crit, opt, [devb[0]], reducer=reducer, warmup=0, cuts=cuts, grad_exchange=args.grad_exchange)
        _lib.profile_pause()
        pstep(0)
        barrier()
        acc = {k: [0.0, 0.0, 0] for k in (0, 1, 2)}
        reps = min(args.steps, 5)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(reps):
            pe0.record()
            pstep(0)
            pe1.record()
            barrier()
            tot += pe0.elapsed_time(pe1)
            for k in acc:
                t, w, n = _lib.profile_end(k)
                acc[k][0] += t; acc[k][1] += w; acc[k][2] += n
        pstep.close()
        prof = {'reps': reps, 'ms_per_step_instrumented': tot / reps, 'kinds': acc}
    elif with_roofline and not launch_list:
        _lib.profile_begin()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for i in range(args.steps):
            step(devb[i % nbuf])
        pe1.record()
        barrier()
        prof = {'reps': args.steps, 'ms_per_step_instrumented': pe0.elapsed_time(pe1) / args.steps,
                'kinds': {k: list(_lib.profile_end(k)) for k in (0, 1, 2)}}
    if graphed is not None:
        graphed.close()
    return dict(name=name, cfg=cfg, model=model, B=B, C=C, ms=ms, e2e_ms=e2e_ms, launches=int(launches), loss=loss_val, clk=clk,
                prof=prof, gate=gate, h2d_bytes=h2d_bytes, graphed=graphed is not None, cuts=cuts, devb=devb,
                optimizer='torch fused AdamW' if args.torch_adamw else 'ArenaAdamW (one pass: update + bf16 shadow + grad zero-fill)')


def slot_grid(dev, peak_gbs):
    """BASELINE.json configs[4]: streaming slot attention, 1568 tokens x 768, S in {2, 4, 8}, B in {1 .. 512}: one forward pass
    (= one aggregation layer) and its backward, achieved GB/s over the algorithmic bytes"""
    import torch
    from devias_b200 import ops

    def timeit(fn, n, graph_calls=0):
        """ms per call; warm-up and timed region are both >= ~40 ms long so that the SM clock has settled on this (memory-bound)
        load after the power-capped GEMM phases that ran before.  graph_calls > 0: that many calls are captured into one CUDA
        graph and the graph is replayed -- below ~64 clips a call is shorter than the ~30 us the python / ctypes launch path
        takes, and the aggregation block's callers replay it from a graph anyway"""
        if graph_calls:
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(graph_calls):
                    fn()
            call, per_call = g.replay, graph_calls
        else:
            call, per_call = fn, 1
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s0.record()
        for _ in range(3):
            call()
        s1.record(); torch.cuda.synchronize()
        per = max(s0.elapsed_time(s1) / 3, 1e-3)
        reps = int(min(max(n / per_call, 40.0 / per), 2000))
        for _ in range(reps):
            call()
        torch.cuda.synchronize(); s0.record()
        for _ in range(reps):
            call()
        s1.record(); torch.cuda.synchronize()
        return s0.elapsed_time(s1) / reps / per_call

    out = []
    for dt in (torch.float32, torch.bfloat16):
        for S in (2, 4, 8):
            for Bm in (1, 8, 64, 256, 512):
                try:
                    if dt is not torch.float32 and not getattr(ops, 'SLOT_BF16_TOKENS', False):
                        out.append({'S': S, 'batch': Bm, 'tokens': 'bf16', 'unsupported': 'fp32 token stream only'})
                        continue
                    tok = (torch.randn(Bm, 1568, 768, device=dev) * 1.5).to(dt)
                    nbytes = tok.numel() * tok.element_size()
                    # small batches would be served from the 126 MB L2 on repeated passes: rotate over copies totalling > 2x L2
                    copies = 1 if nbytes >= (300 << 20) else (300 << 20) // nbytes + 1
                    toks = [tok] + [tok.clone() for _ in range(copies - 1)]
                    g_ = torch.randn(Bm, 4 * S, 768, device=dev) * 0.05
                    G_ = g_.sum(-1).contiguous(); c0_ = torch.randn(Bm, 4 * S, device=dev)
                    cnt = [0]

                    def fwd():
                        cnt[0] += 1
                        return ops.slot_stream_fwd(toks[cnt[0] % copies], g_, G_, c0_)
                    n = 10 if Bm >= 64 else 2 * copies if copies > 10 else 20
                    gc = 0 if Bm >= 64 else 2 * copies            # small batches: graph replay (every rotating copy twice per replay)
                    tf = timeit(fwd, n, gc)
                    U_, m_, A_, at_, mu_, r_ = ops.slot_stream_fwd(tok, g_, G_, c0_)
                    dU_, dm_, dA_ = torch.randn_like(U_), torch.randn_like(m_), torch.randn_like(A_)

                    def bwd():
                        cnt[0] += 1
                        return ops.slot_stream_bwd(toks[cnt[0] % copies], mu_, r_, g_, G_, at_, dU_, dm_, dA_)
                    tb = timeit(bwd, n, gc)
                    fb = nbytes + Bm * 4 * S * 1568 * 4            # tokens once + the returned slot-axis softmax
                    bb = nbytes + Bm * 1568 * 768 * 4              # tokens once + the token gradient written once
                    out.append({'S': S, 'batch': Bm, 'tokens': 'f32' if dt is torch.float32 else 'bf16',
                                'fwd_us': tf * 1e3, 'fwd_gbs': fb / tf / 1e6, 'fwd_frac': fb / tf / 1e6 / peak_gbs,
                                'bwd_us': tb * 1e3, 'bwd_gbs': bb / tb / 1e6, 'bwd_frac': bb / tb / 1e6 / peak_gbs,
                                'rotating_copies': copies, 'launch': 'cuda-graph replay' if gc else 'eager'})
                    del tok, toks, g_, G_, c0_, U_, m_, A_, at_, mu_, r_, dU_, dm_, dA_
                except Exception as e:  # the grid must never take the headline down
                    out.append({'S': S, 'batch': Bm, 'error': repr(e)[:200]})
    return out


def eval_sweep(model, dev, sizes=(1, 2, 4, 8, 16, 32, 64, 128, 256)):
    """BASELINE.json configs[3]: action + scene logits on synthetic clips, batch 1..256, graph-captured eval forward"""
    import torch
    model.eval()
    out = []
    with torch.no_grad():
        for Bv in sizes:
            try:
                x = torch.randn(Bv, 3, 16, 224, 224, device=dev)
                for _ in range(2):
                    model(x)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    model(x)
                for _ in range(2):
                    g.replay()
                n = max(3, min(30, 600 // Bv))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record()
                for _ in range(n):
                    g.replay()
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                out.append({'batch': Bv, 'latency_ms': ms, 'clips_per_s': Bv / ms * 1e3, 'fwd_tflops': Bv * FWD_GFLOP_PER_CLIP / ms})
                del g, x
            except Exception as e:
                out.append({'batch': Bv, 'error': repr(e)[:200]})
    model.train()
    return out


def gpu_reference(cfg, dev, clips=8):
    """The reference's PyTorch path on the SAME GPU (oracle port of model/modeling_slot.py + agg_block + TrainLoss: eager
    ATen / cuBLAS kernels, materialised 1568 x 1568 attention), fwd + loss + bwd, fp32 and bf16 autocast; bounded sample."""
    import numpy as np
    import torch
    from oracle import devias_oracle as O
    C = cfg['num_classes']
    out = {'clips_per_step': clips, 'what': 'oracle port of the reference PyTorch modules on cuda (eager ATen/cuBLAS), fwd + TrainLoss + bwd'}
    try:
        sd = O.synth_state_dict(num_classes=C, num_latents=cfg['num_latents'], agg_depth=cfg['agg_depth'],
                                agg_weights_tie=cfg['agg_weights_tie'], seed=0)
        moved = {}
        for k, v in sd.items():
            if id(v) not in moved:
                moved[id(v)] = v.to(dev).requires_grad_(True)
        sd = {k: moved[id(v)] for k, v in sd.items()}
        clip = O.synth_clips(clips, seed=0).to(dev)
        rs = np.random.RandomState(0)
        target = torch.from_numpy(rs.randint(0, C, size=(clips,)).astype(np.int64)).to(dev)
        teacher = torch.from_numpy(rs.standard_normal(size=(clips, 365)).astype(np.float32)).to(dev)
        fg = (torch.from_numpy(rs.uniform(size=(clips, 196)).astype(np.float32)).to(dev),
              torch.from_numpy(rs.uniform(size=(clips, 1568)).astype(np.float32)).to(dev))

        def step(autocast):
            for v in moved.values():
                v.grad = None
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                o = O.student_forward(sd, clip, C)
            f = lambda t: t.float() if torch.is_tensor(t) else t
            o = tuple(tuple(f(t) for t in grp) for grp in o)
            total, _, _ = O.train_loss(o, teacher, target, fg, C)
            total.backward()

        for name, ac in (('bf16_autocast', True), ('fp32', False)):
            step(ac)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            n = 3
            for _ in range(n):
                step(ac)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            out[name] = {'ms_per_step': ms, 'clips_per_s': clips / ms * 1e3}
        del sd, moved, clip
        torch.cuda.empty_cache()
    except Exception as e:
        out['error'] = repr(e)[:300]
    return out


def main():
    args = parse()
    cfg = dict(WORKLOADS[args.workload])
    if args.batch:
        cfg['batch'] = args.batch
    if args.impl == 'reference':
        return run_reference(args, cfg)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def reduce_max(*vals):
        t = torch.tensor([v or 0.0 for v in vals], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    peak_tf, peak_gbs, peak_src = peaks()

    def summarise(r, steps):
        ms, e2e_ms = reduce_max(r['ms'], r['e2e_ms'])
        B = r['B']
        value = world * B * steps / (ms * 1e-3)
        d = {'value': value, 'ms_per_step': ms / steps, 'clips_per_gpu': B, 'loss': r['loss'], 'gpu_launches': r['launches'],
             'parity_gate': r['gate'], 'step_tensor_frac': value / world * TRAIN_GFLOP_PER_CLIP / 1e3 / peak_tf}
        if e2e_ms:
            d['e2e'] = {'value': world * B * steps / (e2e_ms * 1e-3), 'unit': 'clips/s', 'h2d_bytes_per_step': r['h2d_bytes'],
                        'd2h_bytes_per_step': 4}
        p = r['prof']
        if p:
            k = p['kinds']
            gm, gw, gn = k[0]
            am, aw, an = k[1]
            sm, sw, sn = k[2]
            inst = p['ms_per_step_instrumented'] * p['reps']
            ach = gw / (gm * 1e-3) / 1e12 if gm > 0 else 0.0
            d['roofline'] = {
                'kernel': 'gemm_bf16_kernel (tcgen05/TMEM/TMA)', 'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': ach / peak_tf, 'peak_source': f'{peak_src} (sustained cuBLAS bf16)', 'launches_per_step': int(gn / p['reps']),
                'avg_launch_us': gm / max(gn, 1) * 1e3, 'share_of_step': gm / inst,
                'timed_over': (f'{p["reps"]} replays of the captured training step: external CUDA-event record nodes around every launch of '
                               f'the family inside the graph (instrumented replay {p["ms_per_step_instrumented"]:.2f} ms/step vs '
                               f'{ms / steps:.2f} ms/step uninstrumented)') if r['graphed'] else 'eager steps, events around every launch',
                'attention': {'kernel': 'flash_fwd2_kernel + flash_bwd_kernel', 'tflops': aw / (am * 1e-3) / 1e12 if am > 0 else None,
                              'frac': (aw / (am * 1e-3) / 1e12 / peak_tf) if am > 0 else None, 'share_of_step': am / inst,
                              'launches_per_step': int(an / p['reps'])},
                'slot_attention': {'in_step_gbs': sw / (sm * 1e-3) / 1e9 if sm > 0 else None, 'peak_gbs': peak_gbs,
                                   'share_of_step': sm / inst, 'launches_per_step': int(sn / p['reps'])},
            }
        return d

    primary = measure_train(args.workload, cfg, args, dev, world, rank)
    out_p = summarise(primary, args.steps)
    model = primary.pop('model')
    secondary = None
    if args.workload == 'k400' and not args.no_secondary and not args.batch:
        extras = None
        if rank == 0 and world == 1 and not args.no_extras:
            extras = {'eval_sweep': eval_sweep(model, dev), 'slot_grid': slot_grid(dev, peak_gbs),
                      'gpu_reference': gpu_reference(cfg, dev)}
        del model
        primary.pop('devb')
        torch.cuda.empty_cache()
        sec = measure_train('ucf', dict(WORKLOADS['ucf']), args, dev, world, rank)
        secondary = summarise(sec, args.steps)
        secondary['workload'] = workload_text('ucf', sec['cfg'], sec['C'])
        sec.pop('model'); sec.pop('devb')
    else:
        extras = None
        if rank == 0 and world == 1 and not args.no_extras:
            extras = {'eval_sweep': eval_sweep(model, dev), 'slot_grid': slot_grid(dev, peak_gbs),
                      'gpu_reference': gpu_reference(cfg, dev)}
        del model

    if rank == 0:
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, 'profiles', 'gemm_traffic.json')
        if os.path.isfile(tp):
            try:
                tj = json.load(open(tp))
                traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
            except Exception:
                traffic = None
        C, B = primary['C'], primary['B']
        mode = ('eager' if not primary['graphed'] else 'cuda-graph replay: fwd+loss+bwd graph + update graph' if world == 1 else
                f'cuda graphs per step: backward cut at encoder blocks {primary["cuts"]}; the NCCL all-reduce of each finished gradient '
                f'range (flat arena, exchanged as {args.grad_exchange}) overlaps the next backward piece; update graph last')
        out = {
            'metric': METRIC, 'value': out_p['value'], 'unit': 'clips/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': out_p['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': workload_text(args.workload, cfg, C), 'clips_per_gpu': B, 'global_batch': B * world,
                       'parallelism': f'dp{world}',
                       'l2': 'per-step working set (activations + weights, tens of GB) far exceeds the 126 MB L2; no flush needed',
                       'loss': out_p['loss'], 'launch_mode': mode, 'optimizer': primary['optimizer']},
            'clocks': primary['clk'],
            'gpu_launches': out_p['gpu_launches'],
            'parity_gate': out_p['parity_gate'],
        }
        if 'roofline' in out_p:
            out['roofline'] = out_p['roofline']
            out['roofline']['traffic'] = traffic
            out['roofline']['traffic_source'] = traffic_src
            out['roofline']['step_tensor_frac'] = out_p['step_tensor_frac']
        if 'e2e' in out_p:
            out['e2e'] = out_p['e2e']
        if secondary is not None:
            out['secondary'] = secondary
        if extras:
            out.update(extras)
            grid = [g for g in extras['slot_grid'] if g.get('S') == cfg['num_latents'] and g.get('batch') == 256 and 'fwd_gbs' in g]
            if grid and 'roofline' in out:
                f32 = [g for g in grid if g.get('tokens') == 'f32'] or grid
                b16 = [g for g in grid if g.get('tokens') == 'bf16']
                out['roofline']['slot_attention']['microbench'] = f32[0]          # fp32 token stream (what the model feeds, 1e-5 contract)
                out['roofline']['slot_attention']['microbench_frac_of_hbm'] = f32[0]['fwd_frac']
                if b16:                                                            # bf16 token stream: the tcgen05 kernels
                    out['roofline']['slot_attention']['microbench_bf16'] = b16[0]
                    out['roofline']['slot_attention']['microbench_bf16_frac_of_hbm'] = b16[0]['fwd_frac']
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
