"""CUDA-graph capture of one forward+backward step.

The shift-family blocks are sequenced from Python (hundreds of C-ABI launches per step, ~30 us of host time each), so at
moderate batch sizes the GPU waits for the host; the whole step is static-shaped and free of host synchronisation,
which makes it capturable once and replayable with a single launch.  Every kernel of the library runs on torch's
current stream, so it lands in the capture like any ATen op; TMA descriptors are encoded on the host at capture time and
are baked into the graph's kernel parameters (activations come from the graph's private memory pool, parameters keep
their addresses).
"""
import torch


class GraphedStep:
    """loss = loss_fn(model(x)); loss.backward() as one CUDA graph.  Gradients land in ``p.grad`` (static tensors that
    every replay OVERWRITES -- the captured step starts from grad=None, so there is no accumulation across replays;
    ``run()`` re-attaches them if ``zero_grad(set_to_none=True)`` detached them in between).

    ``step_fn(x) -> loss`` replaces the default body.  ``ddp`` (a ``dp.DataParallel``) makes it the data-parallel step:
    the graph holds this rank's compute only (no collective is captured) and ``run()`` follows every replay with ONE
    grouped NCCL all-reduce over the step's static gradient buffers (``DataParallel.reduce_static``).  NCCL's watchdog
    thread touches CUDA while we capture, hence the thread-local capture mode.
    """

    def __init__(self, model, example_x, loss_fn, warmup=3, step_fn=None, ddp=None):
        self.model, self.loss_fn, self.ddp = model, loss_fn, ddp
        self.static_x = example_x.clone()
        if step_fn is None:
            def step_fn(x):
                loss = loss_fn(model(x))
                loss.backward()
                return loss
        if ddp is not None:
            # data parallel: the graph holds the compute of this rank's shard only; the block gradient buckets announce
            # themselves to `ddp` during the capture and are exchanged by ONE all-reduce of the gradient arena after each replay
            def step_fn(x, _inner=step_fn):
                with ddp:
                    return _inner(x)
            ddp.begin_static_capture()                     # also covers the warm-up steps: nothing is exchanged in them
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                  # warm-up off the default stream (allocator, lazy inits)
                for _ in range(warmup):
                    model.zero_grad(set_to_none=True)
                    if ddp is not None:
                        ddp.begin_static_capture()         # the LAST warm-up step's buckets size the gradient arena
                    step_fn(self.static_x)
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            if ddp is not None:
                ddp.plan_arena()                           # one flat buffer for all buckets + the other gradients, sized
            model.zero_grad(set_to_none=True)              # from the last warm-up step (while its p.grad still exist)
            if ddp is not None:
                ddp.begin_static_capture()                 # start the bucket list afresh: these are the graph's buffers
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.static_loss = step_fn(self.static_x)
        except BaseException:
            if ddp is not None:
                ddp.abort_static_capture()                 # never leave the wrapper recording after a failed capture
            raise
        # the replay writes gradients into exactly these tensors (graph-private memory): keep them, so that a
        # zero_grad(set_to_none=True) between replays cannot orphan them
        self._grads = [(p, p.grad) for p in model.parameters() if p.grad is not None]
        if ddp is not None:
            ddp.end_static_capture()

    def run(self, x=None, non_blocking=True):
        if x is not None:
            self.static_x.copy_(x, non_blocking=non_blocking)
        self.graph.replay()
        if self.ddp is not None:
            self.ddp.reduce_static()
        for p, g in self._grads:          # re-attach after model.zero_grad() / optimizer.zero_grad() (set_to_none=True)
            if p.grad is not g:
                p.grad = g
        return self.static_loss
