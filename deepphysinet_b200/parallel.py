"""Data-parallel plumbing for the hot path: one process per GPU, samples (or one sample's
query points) sharded across ranks, ONE gradient all-reduce per step over NCCL / NVLink.

The reference specifies this with torch DDP + DistributedSampler (interface_physics.py:899-907,936) but never
initialises a process group; here it is explicit and bucket-free: all gradients are packed into one flat fp32
buffer (22.4 MB for the whole PhysicsNet) and reduced with a single collective, which NVSwitch/NVLS serves at
full bandwidth regardless of message count.  Works with the gloo backend on CPU for the host-logic tests.
"""
import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE (torchrun) and initialises the default process group if needed.
    Returns (rank, local_rank, world_size); a plain `python` launch gives (0, 0, 1) and no process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of n units (samples, or points of one sample) for `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class FlatGradAllReduce:
    """Packs the .grad of `params` into one flat buffer and all-reduces it with ONE collective.

    op="mean" (default): writes back SUM / world - DDP semantics for LEVEL-1 sharding (every rank owns whole samples and its
        loss is already the mean over ITS samples; the global loss is the mean over ranks).
    op="sum": writes back the plain SUM - LEVEL-2 sharding (one sample's query points split across ranks, the operator called
        with n_norm = total points of the sample): every rank's gradient is then a partial sum that is already normalised by
        n_norm, and averaging would scale the step by 1/world.  The partial loss terms add up the same way:
        `reduce_partial_losses` below.
    Parameters whose grad is None contribute zeros."""

    def __init__(self, params: Iterable[torch.nn.Parameter], op: str = "mean"):
        if op not in ("mean", "sum"):
            raise ValueError("op must be 'mean' (sample sharding) or 'sum' (point sharding), got %r" % (op,))
        self.op = op
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        self.flat = None

    def __call__(self, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return None
        p0 = self.params[0]
        if self.flat is None or self.flat.device != p0.device:
            self.flat = torch.zeros(self.numel, dtype=torch.float32, device=p0.device)
        views, off = [], 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
            views.append(v)
        dead = [v for v, p in zip(views, self.params) if p.grad is None]
        live = [(v, p) for v, p in zip(views, self.params) if p.grad is not None]
        if dead:
            torch._foreach_zero_(dead)
        if live:
            torch._foreach_copy_([v for v, _ in live], [p.grad for _, p in live])
        # NCCL divides inside the collective (ReduceOp.AVG); gloo (CPU tests) has no AVG: sum, then one scale pass
        fused_avg = self.op == "mean" and dist.get_backend() == "nccl"
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.AVG if fused_avg else dist.ReduceOp.SUM, async_op=async_op)

        def finish():
            if work is not None:
                work.wait()
            if self.op == "mean" and not fused_avg:
                self.flat.mul_(1.0 / dist.get_world_size())
            dst, src = [], []
            for v, p in zip(views, self.params):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    dst.append(p.grad)
                    src.append(v)
            if dst:
                torch._foreach_copy_(dst, src)
        if async_op:
            return finish
        finish()
        return None


def reduce_partial_losses(total: torch.Tensor, terms: torch.Tensor = None):
    """LEVEL-2 (point) sharding: `total` / `terms` returned by the operator on this rank's points with n_norm = all points of
    the sample are PARTIAL sums; the sample's loss is their sum over ranks.  Returns detached, reduced copies (for logging and
    the loss value - the gradient path is FlatGradAllReduce(op="sum"))."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return total.detach(), terms
    buf = torch.cat([total.detach().reshape(1).double(), terms.detach().reshape(-1).double()]) if terms is not None \
        else total.detach().reshape(1).double()
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf[0].to(total.dtype), (buf[1:].reshape(terms.shape) if terms is not None else None)


def allreduce_max(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
