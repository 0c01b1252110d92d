"""All-reduce timing of gradient-arena-sized buffers on this box's NCCL (run under torchrun): what the data-parallel step pays.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_nccl.py"""
import os
import torch
import torch.distributed as dist
rank, local = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
for dtype in (torch.float32, torch.bfloat16):
    for mb in (8, 33, 100, 197, 394):
        n = mb * (1 << 20) // (4 if dtype is torch.float32 else 2)
        x = torch.ones(n, device='cuda', dtype=dtype)
        for _ in range(5):
            dist.all_reduce(x, op=dist.ReduceOp.AVG)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(x, op=dist.ReduceOp.AVG)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if rank == 0:
            print(f'{str(dtype):16s} {mb:4d} MiB  {ms * 1e3:8.1f} us  algbw {mb * 1.048576 / ms:7.1f} GB/s', flush=True)
dist.destroy_process_group()
