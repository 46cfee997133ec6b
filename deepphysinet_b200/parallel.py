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
    """Packs the .grad of `params` into one flat buffer, all-reduces it (sum) and writes back the mean over ranks
    (DDP semantics: loss = mean over samples).  Parameters whose grad is None contribute zeros."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
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
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)

        def finish():
            if work is not None:
                work.wait()
            self.flat.mul_(1.0 / dist.get_world_size())
            for v, p in zip(views, self.params):
                if p.grad is None:
                    p.grad = v.clone()
                else:
                    p.grad.copy_(v)
        if async_op:
            return finish
        finish()
        return None


def allreduce_max(value: float, device) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
