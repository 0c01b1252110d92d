"""CPU, world_size 2 over gloo: the bucketed gradient reducer averages gradients exactly like a single process on the
concatenated batch, leaves all ranks with identical parameters after a step, and handles gradient accumulation."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from devias_b200.ddp import GradReducer


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.GELU(), torch.nn.LayerNorm(64), torch.nn.Linear(64, 19))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        m = _model()
        red = GradReducer(m, bucket_mb=0.004, first_bucket_mb=0.001)   # tiny buckets -> several of them
        assert len(red.buckets) > 2
        opt = torch.optim.SGD(m.parameters(), lr=0.1)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(8, 37, generator=g); y = torch.randn(8, 19, generator=g)
        xs, ys = x.chunk(world)[rank], y.chunk(world)[rank]
        # step 1: plain
        red.zero_grad()
        ((m(xs) - ys) ** 2).mean().backward()
        red.finish()
        grads1 = [p.grad.clone().numpy() for p in m.parameters()]
        opt.step()
        # step 2: two accumulation micro-steps, exchange on the second
        red.zero_grad()
        red.enabled = False
        ((m(xs[:2]) - ys[:2]) ** 2).mean().backward()
        red.enabled = True
        for b in range(len(red.buckets)):
            red._pending[b] = red.sizes[b]
        ((m(xs[2:]) - ys[2:]) ** 2).mean().backward()
        red.finish()
        opt.step()
        q.put((rank, grads1, [p.detach().clone().numpy() for p in m.parameters()]))
    finally:
        dist.destroy_process_group()


def test_reducer_world2_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference on the whole batch
    m = _model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 37, generator=g); y = torch.randn(8, 19, generator=g)
    ((m(x) - y) ** 2).mean().backward()
    for a, b0, b1 in zip(m.parameters(), res[0][1], res[1][1]):
        assert torch.allclose(a.grad, torch.from_numpy(b0), atol=1e-6) and (b0 == b1).all()
    for p0, p1 in zip(res[0][2], res[1][2]):
        assert (p0 == p1).all()


def test_reducer_single_process_is_passthrough():
    m = _model()
    red = GradReducer(m)
    x = torch.randn(4, 37)
    red.zero_grad()
    m(x).sum().backward()
    red.finish()
    m2 = _model()
    m2(x).sum().backward()
    for a, b in zip(m.parameters(), m2.parameters()):
        assert torch.allclose(a.grad, b.grad)
        assert a.grad.data_ptr() == red._views[a].data_ptr()


def _worker_ranges(rank, world, port, q):
    """the exchange pattern of engine.GraphedTrainStep: hooks off, the arena reduced as an upper and a lower range"""
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        m = _model()
        red = GradReducer(m)
        red.enabled = False
        lower = list(m[0].parameters())                     # executed first in forward = last in backward
        low_ids = {id(p) for p in lower}
        upper = [p for p in red.params if id(p) not in low_ids]
        lo_rng, up_rng = red.range_of(lower), red.range_of(upper)
        total = sum((p.numel() + 7) // 8 * 8 for p in red.params)
        assert sorted([lo_rng, up_rng]) == [(0, sorted([lo_rng, up_rng])[0][1]), (sorted([lo_rng, up_rng])[0][1], total)]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(8, 37, generator=g); y = torch.randn(8, 19, generator=g)
        xs, ys = x.chunk(world)[rank], y.chunk(world)[rank]
        red.zero_grad()
        ((m(xs) - ys) ** 2).mean().backward()
        red.allreduce_range(*up_rng)
        red.allreduce_range(*lo_rng)
        try:
            red.range_of([list(m.parameters())[0], list(m.parameters())[-1]])   # not contiguous in the arena
            ok = False
        except AssertionError:
            ok = True
        q.put((rank, [p.grad.clone().numpy() for p in m.parameters()], ok))
    finally:
        dist.destroy_process_group()


def test_range_exchange_world2_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_ranges, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    m = _model()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(8, 37, generator=g); y = torch.randn(8, 19, generator=g)
    ((m(x) - y) ** 2).mean().backward()
    for a, b0, b1 in zip(m.parameters(), res[0][1], res[1][1]):
        assert torch.allclose(a.grad, torch.from_numpy(b0), atol=1e-6) and (b0 == b1).all()
    assert res[0][2] and res[1][2]


class _ToyStudent(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.net = _model()

    def forward(self, x):
        return self.net(x)


def _toy_criterion(model, student_output, teacher_output, target, fg_mask=None):
    loss = ((student_output - target) ** 2).mean()
    return loss, student_output, {}


def _worker_accumulate(rank, world, port, q):
    """gradient accumulation driven THROUGH engine.train_step with a reducer (update_freq = 2)"""
    from devias_b200 import engine
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        m = _ToyStudent()
        red = GradReducer(m, bucket_mb=0.004, first_bucket_mb=0.001)
        opt = torch.optim.SGD(m.parameters(), lr=0.1)
        g = torch.Generator().manual_seed(1)
        x = torch.randn(16, 37, generator=g); y = torch.randn(16, 19, generator=g)
        xs, ys = x.chunk(world)[rank], y.chunk(world)[rank]
        for it in range(2):                                   # two optimizer steps, each of two micro-steps
            for u in range(2):
                sl = slice(4 * u, 4 * u + 4)
                engine.train_step(m, None, _toy_criterion, opt, xs[sl], ys[sl], None, teacher_logits=ys[sl], update_freq=2,
                                  do_update=(u == 1), reducer=red)
        q.put((rank, [p.detach().clone().numpy() for p in m.parameters()]))
    finally:
        dist.destroy_process_group()


def test_accumulation_through_train_step_world2_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_accumulate, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for p0, p1 in zip(res[0][1], res[1][1]):
        assert (p0 == p1).all(), 'ranks diverged under gradient accumulation'
    # single process: every optimizer step sees the mean gradient over the same 16 samples
    m = _ToyStudent()
    opt = torch.optim.SGD(m.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(16, 37, generator=g); y = torch.randn(16, 19, generator=g)
    for it in range(2):
        opt.zero_grad()
        ((m(x) - y) ** 2).mean().backward()
        opt.step()
    for a, b0 in zip(m.parameters(), res[0][1]):
        assert torch.allclose(a.detach(), torch.from_numpy(b0), atol=1e-6)


def test_finish_raises_when_exchange_is_off():
    """a reducer whose exchange is switched off must not silently let the ranks step on local gradients"""
    m = _model()
    red = GradReducer(m)
    red.world = 2                     # pretend: no process group is needed to reach the check
    red.enabled = False
    try:
        red.finish()
        raised = False
    except RuntimeError:
        raised = True
    assert raised
