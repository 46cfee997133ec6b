"""CUDA-graph capture of one whole PDE step through the reference-facing call.

The library's scratch memory is allocated INSIDE the capture (deepphysinet_b200._native.workspace), i.e. from the graph's private
pool: it lives exactly as long as this object and no later eager call can free or reuse it.

`place_one_batch` + `backward()` is ~500 small PyTorch launches (encoder, hyper-network, their backward) around the
fused operator; on the host they cost more than the GPU work they enqueue.  `GraphedPlaceOneBatch` captures the whole
step once - host->device copies of the pinned inputs, encoder, fused operator, backward - and replays it with a single
launch.  The library itself is capture-safe (no allocation, no synchronisation, caller's stream).

    step = GraphedPlaceOneBatch(model, (x, y, t, f, field, input_data, forecast_h), criterion, loss_factor, device)
    x.copy_(new_x) ...            # refresh the PINNED host buffers in place
    loss = step()                  # one graph launch; gradients are in model.physics_net parameters' .grad

`PrefetchedPlaceOneBatch` adds double buffering: the host -> device copies of step i+1 run on a copy stream underneath the
compute of step i (two device input sets, two graphs sharing one memory pool):

    step = PrefetchedPlaceOneBatch(model, host_inputs, criterion, loss_factor, device)
    step.prefetch()                # H2D of the batch now in the pinned buffers (asynchronous)
    for ...:
        loss = step()              # computes on the batch prefetched last
        refill the pinned buffers with the next batch (after step.copied() / any sync), then step.prefetch()
"""
import torch


class GraphedPlaceOneBatch:
    def __init__(self, model, host_inputs, criterion, loss_factor, device, rank=0, warmup=3):
        self.model, self.inputs, self.device = model, tuple(host_inputs), torch.device(device)
        for a in self.inputs:
            if a.is_cuda or not a.is_pinned():
                raise ValueError("GraphedPlaceOneBatch replays host->device copies: inputs must be pinned host tensors")
        self.criterion, self.loss_factor, self.rank = criterion, loss_factor, rank
        params = [p for p in model.physics_net.parameters() if p.requires_grad]
        # warm-up and capture run on side streams on purpose; the AccumulateGrad nodes of the warm-up are dropped with zero_grad below
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                      # warm-up on a side stream (allocator, workspace, func attributes)
            for _ in range(warmup):
                model.physics_net.zero_grad(set_to_none=True)
                self._eager().backward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from . import _native
        _native.release_workspaces()                       # the warm-up's scratch was keyed to the side stream; the capture allocates its own
        model.physics_net.zero_grad(set_to_none=True)      # grads are (re)created inside the capture: static addresses
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
            self.loss.backward()
        self.params, self.grads = params, [p.grad for p in params]

    def _eager(self):
        x, y, t, f, field, data, fh = self.inputs
        return self.model.place_one_batch(x, y, t, f, field, data, fh, self.criterion, self.loss_factor, 0, self.rank,
                                          self.device)

    def __call__(self):
        self.graph.replay()
        for p, g in zip(self.params, self.grads):          # the graph owns the gradient buffers: (re)attach them
            p.grad = g
        return self.loss


class PrefetchedPlaceOneBatch:
    """Double-buffered variant: two device copies of the inputs, two CUDA graphs (one per copy, sharing a memory pool and
    therefore the library workspace), host -> device copies on a dedicated stream.  `prefetch()` enqueues the copies of the
    batch currently in the pinned host buffers into the idle set; `__call__()` waits for that set's copy event on the compute
    stream and replays its graph.  Per step exactly one H2D of every input still happens - it just no longer sits in front
    of the encoder on the critical path."""

    def __init__(self, model, host_inputs, criterion, loss_factor, device, rank=0, warmup=3):
        self.model, self.host, self.device = model, tuple(host_inputs), torch.device(device)
        for a in self.host:
            if a.is_cuda or not a.is_pinned():
                raise ValueError("PrefetchedPlaceOneBatch copies from pinned host tensors")
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.sets = [tuple(torch.empty_like(a, device=self.device) for a in self.host) for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        self.criterion, self.loss_factor, self.rank = criterion, loss_factor, rank
        params = [p for p in model.physics_net.parameters() if p.requires_grad]
        quiet = getattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch", None)
        if quiet is not None:
            quiet(False)
        for s in self.sets:                                  # something valid to warm up / capture on
            for d, h in zip(s, self.host):
                d.copy_(h, non_blocking=True)
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                model.physics_net.zero_grad(set_to_none=True)
                self._eager(self.sets[0]).backward()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from . import _native
        _native.release_workspaces()
        model.physics_net.zero_grad(set_to_none=True)
        pool = torch.cuda.graph_pool_handle()
        self.graphs, self.losses, self.grads = [], [], []
        for s in self.sets:
            model.physics_net.zero_grad(set_to_none=True)    # each graph owns its gradient buffers
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                loss = self._eager(s)
                loss.backward()
            self.graphs.append(g)
            self.losses.append(loss)
            self.grads.append([p.grad for p in params])
        self.params = params
        self.next_fill, self.next_run, self.pending = 0, 0, 0

    def _eager(self, dev_inputs):
        x, y, t, f, field, data, fh = dev_inputs
        return self.model.place_one_batch(x, y, t, f, field, data, fh, self.criterion, self.loss_factor, 0, self.rank, self.device)

    def prefetch(self):
        """Asynchronous H2D of the pinned host buffers into the idle device set."""
        if self.pending >= 2:
            raise RuntimeError("both device input sets hold batches that have not been consumed yet")
        i = self.next_fill
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i])    # the graph that last read this set has finished
            for d, h in zip(self.sets[i], self.host):
                d.copy_(h, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.next_fill ^= 1
        self.pending += 1

    def copied(self):
        """Blocks the host until the last prefetch has left the pinned buffers (they may be refilled afterwards)."""
        self.copy_stream.synchronize()

    def __call__(self):
        if self.pending == 0:
            self.prefetch()
        i = self.next_run
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[i])
        self.graphs[i].replay()
        self.consumed[i].record(cur)
        for p, g in zip(self.params, self.grads[i]):
            p.grad = g
        self.next_run ^= 1
        self.pending -= 1
        return self.losses[i]
