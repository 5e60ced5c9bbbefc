"""Data-parallel wrapper: one process per GPU, batch sharded across ranks, ONE gradient all-reduce per step.

The reference has no distributed code; no block mixes samples (ConvMixer's BatchNorm aside),
so the path shards along the batch axis with a single exchange: the gradient average.  The
fused block backward hands back all parameter gradients of a block as views of one flat bf16 buffer; with an
active DataParallel context that buffer is all-reduced in place on NCCL's stream as soon as the block's
backward has been enqueued, so the exchange of block i overlaps the backward of blocks i-1 .. 0.
Parameters outside the fused blocks (stem / head) go in one more flat bucket at the end.
"""
import torch
import torch.distributed as dist

_active = None


def active():
    return _active


class DataParallel(torch.nn.Module):
    def __init__(self, module, process_group=None):
        super().__init__()
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        # NCCL averages in the collective; gloo (CPU tests of this host logic) has no AVG: sum, then scale
        self._nccl = dist.is_initialized() and dist.get_backend(process_group) == "nccl"
        import os
        # "overlap": each block's bucket is reduced on NCCL's stream while the backward of earlier blocks runs;
        # "end": all buckets are reduced back-to-back after backward (the GEMMs are persistent 1-CTA/SM kernels with a
        # static tile schedule, so an SM borrowed by a NCCL CTA stretches the whole GEMM; VMLP_DP_MODE picks)
        self.mode = os.environ.get("VMLP_DP_MODE", "overlap")
        self._deferred = []
        self._pending = []      # (work handle, flat bucket tensor)
        self._bucketed = set()  # data_ptrs already covered by an in-flight bucket
        self._collect = None    # list being filled while a CUDA-graph capture records the static gradient buffers
        self._static = None     # gradient tensors of a captured step (fixed addresses): exchanged after every replay
        self._arena = None      # ONE flat bf16 buffer holding every block bucket of a captured step (+ room for the rest)
        self._arena_off = 0
        self._arena_plan = None
        self._other = []
        if self.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)   # identical replicas

    def forward(self, *a, **kw):
        return self.module(*a, **kw)

    # called by the fused block backward (ops.py) with the block's flat bf16 gradient buffer and its owning parameters
    def reduce_bucket_async(self, flat, owners=()):
        if self.world == 1:
            return
        # The in-place reduction of `flat` is only sound when autograd will STEAL its views as the new p.grad (p.grad is
        # None).  With gradient accumulation / zero_grad(set_to_none=False), AccumulateGrad runs `p.grad += view` on the
        # compute stream while the collective rewrites the buffer, and finish() would average p.grad a second time.
        # Such a bucket is not exchanged here at all: its parameters fall through to finish()'s "rest" path, which
        # averages the accumulated p.grad once, after backward.
        if any(p.grad is not None for p in owners):
            return
        if self._collect is not None:   # CUDA-graph capture: no collective inside the graph, remember the buffer
            self._collect.append(flat)
            return
        if self.mode == "end":          # exchange after the whole backward: no SM contention with the persistent GEMMs
            self._deferred.append(flat)
            return
        work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM, group=self.group,
                               async_op=True)
        self._pending.append((work, flat))

    def __enter__(self):
        global _active
        _active = self
        self._pending.clear()
        self._deferred.clear()
        return self

    def __exit__(self, *exc):
        global _active
        _active = None
        return False

    def finish(self):
        """Reduce what the block buckets did not cover, then join the communication stream."""
        if self.world == 1:
            return
        for flat in self._deferred:
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM, group=self.group,
                                   async_op=True)
            self._pending.append((work, flat))
        self._deferred.clear()
        covered = []
        for _, flat in self._pending:
            covered.append((flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size()))
        rest = []
        for p in self.module.parameters():
            if p.grad is None:
                continue   # never-used parameters have no gradient on any rank
            a = p.grad.data_ptr()
            if not any(lo <= a < hi for lo, hi in covered):
                rest.append(p)
        if rest:
            flat = torch._utils._flatten_dense_tensors([p.grad for p in rest])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM, group=self.group)
            if not self._nccl:
                flat.div_(self.world)
            for p, g in zip(rest, torch._utils._unflatten_dense_tensors(flat, [p.grad for p in rest])):
                p.grad.copy_(g)
        for work, flat in self._pending:
            work.wait()          # makes the current stream wait for the NCCL stream
            if not self._nccl:
                flat.div_(self.world)
        self._pending.clear()

    # ---- CUDA-graph mode: the captured step is compute only; ONE NCCL all-reduce follows every replay ------------------
    def begin_static_capture(self):
        """Call before capturing loss.backward() in a CUDA graph: block buckets are recorded instead of exchanged."""
        self._collect = []
        self._arena_off = 0

    def abort_static_capture(self):
        """A capture that failed must not leave the wrapper in recording mode (eager steps would then record their
        buckets instead of exchanging them)."""
        self._collect = None
        self._arena = None
        self._arena_plan = None
        self._static = None

    def plan_arena(self):
        """Call after a warm-up step run under begin_static_capture(): sizes ONE flat bf16 arena for all block buckets of a
        step plus every gradient outside them.  The capture that follows takes its buckets from the arena in the same
        order (`take`), so the whole gradient exchange of a step is a single all-reduce of one buffer -- 18 grouped
        all-reduces of ~10 MB cost 1.1 ms per step on 8 GPUs (per-operation latency), one of 123 MB about half of that."""
        buckets = self._collect or []
        covered = [(f.data_ptr(), f.data_ptr() + f.numel() * f.element_size()) for f in buckets]
        rest = [p for p in self.module.parameters()
                if p.grad is not None and not any(lo <= p.grad.data_ptr() < hi for lo, hi in covered)]
        if not buckets and not rest:
            return
        ref = buckets[0] if buckets else rest[0].grad
        rest = [p for p in rest if p.grad.dtype == ref.dtype]       # other dtypes stay separate buffers
        pad8 = lambda n: (n + 7) & ~7
        n_b = sum(pad8(f.numel()) for f in buckets)
        n_r = sum(pad8(p.numel()) for p in rest)
        self._arena = torch.zeros(n_b + n_r, dtype=ref.dtype, device=ref.device)
        self._arena_plan = (n_b, [f.numel() for f in buckets])

    def take(self, n, device):
        """A bucket of n elements from the arena (None when no arena is active: the caller allocates)."""
        if self._arena is None or self._collect is None:
            return None
        n_b, sizes = self._arena_plan
        k = len(self._collect)
        if k >= len(sizes) or sizes[k] != n or self._arena.device != device:
            raise RuntimeError("data-parallel gradient arena: the captured step produces different buckets than the warm-up step")
        out = self._arena[self._arena_off:self._arena_off + n]
        self._arena_off += (n + 7) & ~7
        return out

    def end_static_capture(self):
        """Call after the capture: fixes the list of gradient buffers (block buckets + every p.grad outside them)."""
        buckets, self._collect = self._collect, None
        covered = [(f.data_ptr(), f.data_ptr() + f.numel() * f.element_size()) for f in buckets]
        rest = [p.grad for p in self.module.parameters()
                if p.grad is not None and not any(lo <= p.grad.data_ptr() < hi for lo, hi in covered)]
        self._static = buckets + rest
        self._rest_views = None
        in_arena = self._arena is not None and (
            (buckets and buckets[0].data_ptr() == self._arena.data_ptr()) or (not buckets and self._arena_plan[0] == 0))
        if in_arena:
            mine = [g for g in rest if g.dtype == self._arena.dtype]
            self._other = [g for g in rest if g.dtype != self._arena.dtype]
            off, views = self._arena_plan[0], []
            if off + sum((g.numel() + 7) & ~7 for g in mine) > self._arena.numel():
                raise RuntimeError("data-parallel gradient arena: planned before the warm-up gradients existed")
            for g in mine:                               # gradients outside the blocks travel in the arena's tail
                views.append(self._arena[off:off + g.numel()].view(g.shape))
                off += (g.numel() + 7) & ~7
            self._rest, self._rest_views = mine, views
        else:
            self._arena = None

    def reduce_static(self):
        """Average the captured step's gradients over ranks after the backward has finished -- nothing of the exchange
        shares SMs with the persistent GEMM kernels, and the replayed graph contains no collective (no teardown-order
        hazards).  With the arena: two multi-tensor copies for the few gradients outside the blocks and ONE all-reduce;
        without it (buckets not from the arena): one grouped NCCL launch over all buffers."""
        if self.world == 1 or not self._static:
            return
        if self._arena is not None and self._rest_views is not None:
            if self._rest:
                torch._foreach_copy_(self._rest_views, self._rest)
            if self._nccl:
                dist.all_reduce(self._arena, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(self._arena, op=dist.ReduceOp.SUM, group=self.group)
                self._arena.div_(self.world)
            if self._rest:
                torch._foreach_copy_(self._rest, self._rest_views)
            for t in self._other:                            # gradients of another dtype than the arena (none in the models here)
                dist.all_reduce(t, op=dist.ReduceOp.AVG if self._nccl else dist.ReduceOp.SUM, group=self.group)
                if not self._nccl:
                    t.div_(self.world)
            return
        if self._nccl:
            with dist._coalescing_manager(group=self.group, device=self._static[0].device, async_ops=False):
                for t in self._static:
                    dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
        else:                                   # gloo (CPU tests of this logic)
            for t in self._static:
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
                t.div_(self.world)

    def step_fwd_bwd(self, x, loss_fn):
        """One data-parallel fwd+bwd on this rank's shard; gradients are averaged over ranks on return."""
        with self:
            loss = loss_fn(self.module(x))
            loss.backward()
            self.finish()
        return loss
