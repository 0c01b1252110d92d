import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
import test_engine_gpu as T
from devias_b200 import engine, functional
from devias_b200.arena import ParamArena
from devias_b200.loss import TrainLoss
from devias_b200.optim import ArenaAdamW
C = 11
bs = [T._batch(C), T._batch(C)]
crit = TrainLoss(None, 'KL', C)
lrs = [2e-3, 1e-3, 3e-3]

def run_eager(kind):
    m = T._model(C)
    init = T._params(m)
    if kind == 'torch':
        opt = torch.optim.AdamW(T._adamw_groups(m, lrs[0]), betas=(0.9, 0.999), eps=1e-8)
    else:
        opt = ArenaAdamW(T._adamw_groups(m, lrs[0]), ParamArena.of(m), betas=(0.9, 0.999), eps=1e-8)
    for i, lr in enumerate(lrs):
        for g in opt.param_groups: g['lr'] = lr
        b = bs[i % 2]
        engine.train_step(m, None, crit, opt, b['clip'], b['target'], (b['fg'], b['fgf']), teacher_logits=b['teacher'])
    torch.cuda.synchronize()
    p = T._params(m)
    return torch.cat([(p[k] - init[k]).flatten() for k in p]).double()

def run_graph():
    m = T._model(C)
    opt = ArenaAdamW(T._adamw_groups(m, lrs[0]), ParamArena.of(m), betas=(0.9, 0.999), eps=1e-8)
    snap = {k: v.detach().clone() for k, v in m.state_dict().items()}
    init = T._params(m)
    step = engine.GraphedTrainStep(m, crit, opt, bs, warmup=1)
    m.load_state_dict(snap)
    opt.exp_avg.zero_(); opt.exp_avg_sq.zero_(); opt._t = 0; opt.arena.grad.zero_()
    for i, lr in enumerate(lrs):
        for g in opt.param_groups: g['lr'] = lr
        step(i % 2)
    torch.cuda.synchronize()
    p = T._params(m)
    return torch.cat([(p[k] - init[k]).flatten() for k in p]).double()

def rel(a, b): return float((a - b).norm() / b.norm())
t1, t2 = run_eager('torch'), run_eager('torch')
a1 = run_eager('arena')
g1 = run_graph()
print('torch vs torch', rel(t2, t1), 'arena-eager vs torch', rel(a1, t1), 'graph vs torch', rel(g1, t1), 'graph vs arena-eager', rel(g1, a1))
print('norms', float(t1.norm()), float(a1.norm()), float(g1.norm()), 'abs mean', float(t1.abs().mean()), float(g1.abs().mean()))
