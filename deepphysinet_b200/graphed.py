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
