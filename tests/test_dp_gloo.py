"""Data-parallel host logic (jittor-mlp_b200/dp.py) on CPU: world_size 2, gloo.  The fused blocks cannot run
without a GPU, so a plain torch module stands in for the stem/head path ("rest" bucket) and hand-made flat buffers
stand in for the per-block gradient buckets the fused backward registers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, mode, done):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      VMLP_DP_MODE=mode)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import jittor_mlp_b200 as J
    from jittor_mlp_b200 import dp
    torch.manual_seed(100 + rank)                      # different init per rank: the wrapper must broadcast rank 0's
    model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    unused = torch.nn.Parameter(torch.ones(3))         # never receives a gradient (SURVEY.md F6 situation)
    model.register_parameter("unused", unused)
    ddp = dp.DataParallel(model)
    w0 = model[0].weight.detach().clone()
    xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(7))      # global batch 4 -> 2 per rank
    x = xs[rank * 2:(rank + 1) * 2]
    with ddp:
        loss = model(x).square().mean()
        loss.backward()
        # emulate one fused-block bucket: a flat buffer whose views are parameter grads
        bucket = torch.full((5,), float(rank + 1))
        ddp.reduce_bucket_async(bucket)
        ddp.finish()
    q.put((rank, w0, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None},
           bucket.clone(), unused.grad is None))
    done.wait(120)      # tensors travel through the queue by file descriptor: stay alive until the parent has them
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["overlap", "end"])
def test_two_rank_gradient_average_matches_single_process_big_batch(mode):
    """Both bucket schedules of dp.DataParallel: reduce each block's bucket as soon as its backward is enqueued
    ("overlap"), or all buckets back to back after the backward ("end")."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q, done = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, mode, done)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    done.set()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0a, ga, ba, ua), (_, w0b, gb, bb, ub) = res
    assert torch.equal(w0a, w0b)                       # replicas identical after the initial broadcast
    for k in ga:
        assert torch.allclose(ga[k], gb[k])            # every rank holds the same averaged gradient
    assert torch.allclose(ba, torch.full((5,), 1.5)) and torch.allclose(bb, ba)   # bucket averaged in place
    assert ua and ub                                   # unused parameter: grad stays None on every rank
    # reference: one process, the concatenated batch, mean-of-shard-means == mean over the batch for equal shards
    torch.manual_seed(100)
    ref = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.GELU(), torch.nn.Linear(16, 4))
    xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(7))
    ref(xs).square().mean().backward()
    for k, p in ref.named_parameters():
        assert torch.allclose(ga[k], p.grad, atol=1e-6), k


class _FlatLinearFn(torch.autograd.Function):
    """CPU stand-in for a fused block: its backward produces all parameter gradients as views of ONE flat buffer and
    registers that buffer as a DP bucket, exactly like ops._finish_grads does for the CUDA blocks."""

    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w, b)
        return x @ w.t() + b

    @staticmethod
    def backward(ctx, dy):
        from jittor_mlp_b200 import dp
        x, w, b = ctx.saved_tensors
        flat = torch.cat([(dy.t() @ x).flatten(), dy.sum(0)])
        arena = dp.active().take(flat.numel(), flat.device) if dp.active() is not None else None
        if arena is not None:                          # graph-mode capture with a planned gradient arena
            flat = arena.copy_(flat)
        if dp.active() is not None:
            dp.active().reduce_bucket_async(flat, (w, b))
            for work, _ in dp.active()._pending:      # a fast network: the collective lands before AccumulateGrad runs
                work.wait()
        return dy @ w, flat[:w.numel()].view_as(w), flat[w.numel():]


def _accum_worker(rank, world, port, q, done):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      VMLP_DP_MODE="overlap")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jittor_mlp_b200 import dp

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(5)
            self.w = torch.nn.Parameter(torch.randn(4, 8))
            self.b = torch.nn.Parameter(torch.randn(4))
            self.head = torch.nn.Linear(4, 2)

        def forward(self, x):
            return self.head(_FlatLinearFn.apply(x, self.w, self.b))

    model = M()
    ddp = dp.DataParallel(model)
    xs = torch.randn(2, 4, 8, generator=torch.Generator().manual_seed(7))    # 2 micro-batches x global batch 4
    for mb in range(2):                                                       # no zero_grad in between: accumulation
        ddp.step_fwd_bwd(xs[mb, rank * 2:(rank + 1) * 2], lambda o: o.square().mean())
    q.put((rank, {k: p.grad.clone() for k, p in model.named_parameters()}))
    done.wait(120)      # tensors travel through the queue by file descriptor: stay alive until the parent has them
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_accumulation_over_two_micro_batches():
    """with an existing p.grad the block bucket must not be reduced in place while AccumulateGrad adds its
    views into p.grad (and must not be averaged twice): the accumulated gradient equals the single-process sum of the
    two global-batch gradients on every rank."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q, done = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_accum_worker, args=(r, world, port, q, done)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    done.set()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(5)
    w = torch.randn(4, 8, requires_grad=True)
    b = torch.randn(4, requires_grad=True)
    torch.manual_seed(5)
    _ = torch.randn(4, 8), torch.randn(4)
    head = torch.nn.Linear(4, 2)
    xs = torch.randn(2, 4, 8, generator=torch.Generator().manual_seed(7))
    for mb in range(2):
        head(xs[mb] @ w.t() + b).square().mean().backward()
    ref = {"w": w.grad, "b": b.grad, "head.weight": head.weight.grad, "head.bias": head.bias.grad}
    for rank, grads in res:
        for k, g in ref.items():
            assert torch.allclose(grads[k], g, atol=1e-5), (rank, k, (grads[k] - g).abs().max())


def _static_worker(rank, world, port, q, done):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jittor_mlp_b200 import dp
    torch.manual_seed(5)
    w = torch.nn.Parameter(torch.randn(4, 8))
    b = torch.nn.Parameter(torch.randn(4))
    head = torch.nn.Linear(4, 2)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w, self.b, self.head = w, b, head

        def forward(self, x):
            return self.head(_FlatLinearFn.apply(x, self.w, self.b))

    model = M()
    ddp = dp.DataParallel(model)
    xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(7))
    # what GraphedStep(ddp=...) does around the capture, minus the CUDA graph: record the buffers, exchange afterwards
    ddp.begin_static_capture()
    with ddp:
        model(xs[rank * 2:(rank + 1) * 2]).square().mean().backward()
    ddp.end_static_capture()
    n_static = len(ddp._static)
    ddp.reduce_static()
    q.put((rank, n_static, {k: p.grad.clone() for k, p in model.named_parameters()}))
    done.wait(120)      # tensors travel through the queue by file descriptor: stay alive until the parent has them
    dist.barrier()
    dist.destroy_process_group()


def test_static_buffers_of_a_captured_step_are_exchanged_once():
    """CUDA-graph mode of the DP wrapper (graph.GraphedStep(ddp=...)): block buckets are recorded during the capture, the
    remaining p.grad tensors are added, and reduce_static() averages each buffer exactly once."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q, done = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_static_worker, args=(r, world, port, q, done)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    done.set()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(5)
    w = torch.randn(4, 8, requires_grad=True)
    b = torch.randn(4, requires_grad=True)
    head = torch.nn.Linear(4, 2)
    xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(7))
    head(xs @ w.t() + b).square().mean().backward()
    ref = {"w": w.grad, "b": b.grad, "head.weight": head.weight.grad, "head.bias": head.bias.grad}
    for rank, n_static, grads in res:
        assert n_static == 3                      # one block bucket (w, b) + head.weight + head.bias
        for k, g in ref.items():
            assert torch.allclose(grads[k], g, atol=1e-5), (rank, k)


def _arena_worker(rank, world, port, q, with_block, done):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jittor_mlp_b200 import dp
    torch.manual_seed(5)
    w = torch.nn.Parameter(torch.randn(4, 8))
    b = torch.nn.Parameter(torch.randn(4))
    head = torch.nn.Linear(4, 2)

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w, self.b, self.head = w, b, head

        def forward(self, x):
            return self.head(_FlatLinearFn.apply(x, self.w, self.b) if with_block else x @ self.w.t() + self.b)

    model = M()
    ddp = dp.DataParallel(model)
    xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(7))

    def step():
        with ddp:
            model(xs[rank * 2:(rank + 1) * 2]).square().mean().backward()
    # GraphedStep(ddp=...) minus the CUDA graph: warm-up under recording, plan the arena, "capture", exchange
    ddp.begin_static_capture()
    step()
    ddp.plan_arena()
    model.zero_grad(set_to_none=True)
    ddp.begin_static_capture()
    step()
    ddp.end_static_capture()
    in_arena = ddp._arena is not None and (not with_block or model.w.grad.data_ptr() == ddp._arena.data_ptr())
    n_rest = len(ddp._rest_views or [])
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda t, *a, **k: (calls.append(t.numel()), real(t, *a, **k))[1]
    ddp.reduce_static()
    dist.all_reduce = real
    q.put((rank, in_arena, n_rest, calls, {k: p.grad.clone() for k, p in model.named_parameters()}))
    done.wait(120)      # tensors travel through the queue by file descriptor: stay alive until the parent has them
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("with_block", [True, False])
def test_gradient_arena_makes_the_exchange_one_all_reduce(with_block):
    """Graph mode with the arena (dp.plan_arena / take): the block bucket lives in ONE flat buffer, the gradients outside the
    blocks travel in its tail, and reduce_static() issues exactly one all-reduce -- also for a model without fused blocks
    (all gradients travel in the arena)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q, done = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_arena_worker, args=(r, world, port, q, with_block, done)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    done.set()
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(5)
    w = torch.randn(4, 8, requires_grad=True)
    b = torch.randn(4, requires_grad=True)
    head = torch.nn.Linear(4, 2)
    xs = torch.randn(4, 8, generator=torch.Generator().manual_seed(7))
    head(xs @ w.t() + b).square().mean().backward()
    ref = {"w": w.grad, "b": b.grad, "head.weight": head.weight.grad, "head.bias": head.bias.grad}
    for rank, in_arena, n_rest, calls, grads in res:
        if with_block:
            assert in_arena and n_rest == 2       # bucket (w, b) in the arena; head.weight, head.bias in its tail
            assert calls == [40 + 8 + 8]          # one all-reduce: 36 -> 40 (bucket, padded to 8) + 8 + 2 -> 8
        else:
            assert in_arena and n_rest == 4       # w, b, head.weight, head.bias all in the arena
            assert calls == [32 + 8 + 8 + 8]
        for k, g in ref.items():
            assert torch.allclose(grads[k], g, atol=1e-5), (rank, k)
