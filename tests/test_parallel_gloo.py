"""CPU, world_size 2, gloo: the host-side data-parallel logic (sharding + the flat gradient all-reduce)."""
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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from deepphysinet_b200 import parallel as P
    r, lr, w = P.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    lin[1].bias.grad = None                                     # unused parameter on this rank
    P.FlatGradAllReduce(lin.parameters())()
    exp = sum(range(1, world + 1)) / world
    ok = all(torch.allclose(p.grad, torch.full_like(p, exp)) for n, p in lin.named_parameters() if n != "1.bias")
    ok = ok and torch.allclose(lin[1].bias.grad, torch.zeros(3))
    # second step: every parameter has a gradient now (no "dead" list)
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(2 * rank))
    P.FlatGradAllReduce(lin.parameters())()
    exp2 = sum(2 * r for r in range(world)) / world
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, exp2)) for p in lin.parameters())
    # point sharding (level 2): partial gradients and partial loss terms ADD UP, no division by the world size
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    P.FlatGradAllReduce(lin.parameters(), op="sum")()
    ok = ok and all(torch.allclose(p.grad, torch.full_like(p, float(sum(range(1, world + 1))))) for p in lin.parameters())
    tot, terms = P.reduce_partial_losses(torch.tensor(1.5 * (rank + 1)), torch.full((1, 6), float(rank + 1), dtype=torch.float64))
    ok = ok and abs(float(tot) - 1.5 * sum(range(1, world + 1))) < 1e-6 and bool((terms == sum(range(1, world + 1))).all())
    # max-over-ranks timing reduction
    ok = ok and P.allreduce_max(float(rank), "cpu") == world - 1
    # sample sharding covers [0, n) exactly once
    lo, hi = P.shard_range(11, rank, world)
    cover = torch.zeros(11)
    cover[lo:hi] = 1
    dist.all_reduce(cover)
    ok = ok and bool((cover == 1).all())
    q.put((rank, ok))
    dist.destroy_process_group()


def test_flat_grad_allreduce_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_range_is_balanced():
    from deepphysinet_b200.parallel import shard_range
    for n in (1, 8, 13, 65536):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
