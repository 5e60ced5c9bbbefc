"""NCCL all-reduce of the Mixer-B/16 gradient set (61.8 M bf16 elements) in the forms dp.reduce_static can take: one flat
buffer vs 18 grouped buffers, AVG vs SUM; CUDA events around each, max over ranks.  Run under torchrun."""
import os
import sys

import torch
import torch.distributed as dist

dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda")
n_block, n_rest = 5_036_000, [589_824, 768, 768, 768, 768_000, 1000]
flat = torch.randn(12 * n_block + sum(n_rest), device=dev).bfloat16()
bufs = [torch.randn(n_block, device=dev).bfloat16() for _ in range(12)] + [torch.randn(n, device=dev).bfloat16() for n in n_rest]
filler = torch.randn(8192, 8192, device=dev).bfloat16()


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    ts = []
    for _ in range(iters):
        (filler @ filler)                     # the exchange follows a busy GPU, like the step it follows
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def grouped(op):
    with dist._coalescing_manager(device=dev, async_ops=False):
        for b in bufs:
            dist.all_reduce(b, op=op)


res = {"world": world,
       "flat_avg_ms": timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG)),
       "flat_sum_ms": timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.SUM)),
       "grouped18_avg_ms": timed(lambda: grouped(dist.ReduceOp.AVG)),
       "grouped18_sum_ms": timed(lambda: grouped(dist.ReduceOp.SUM)),
       "flat_4chunks_avg_ms": timed(lambda: [dist.all_reduce(c, op=dist.ReduceOp.AVG) for c in flat.chunk(4)])}
if rank == 0:
    print("allreduce_probe", res, flush=True)
dist.barrier()
dist.destroy_process_group()
