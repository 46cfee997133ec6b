"""GPU (-m gpu): CUDA-graph capture of the reference-facing call (deepphysinet_b200.graphed) - the replayed step reproduces the
eager loss and gradients, and follows in-place updates of the pinned host inputs."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_graphed_place_one_batch_matches_eager():
    from deepphysinet_b200 import InterfacePhysics
    from deepphysinet_b200.config import DEFAULT_LOSS_FACTOR, DEFAULT_OBS_NORM
    from deepphysinet_b200.graphed import GraphedPlaceOneBatch
    from oracle import dpn_oracle as O
    obs = {k: dict(v, norm_type="mean_norm", use_norm=True) for k, v in DEFAULT_OBS_NORM.items()}
    torch.manual_seed(0)
    m = InterfacePhysics(H.META_CFG, H.NET_CFG, obs, None, dict(img_size=(145, 257), dx=27000, dy=27000)).cuda()
    g = torch.Generator().manual_seed(5)
    B, N = 2, 700
    pts = [O.synthetic_points(N, g) for _ in range(B)]
    host = [torch.stack([p[i].reshape(-1) for p in pts]).float().pin_memory() for i in range(4)]          # x, y, t, f  [B,N]
    cd = torch.stack([p[4] for p in pts]).float().pin_memory()
    field = torch.randn(B, 159, 2405, generator=g).pin_memory()
    fh = torch.full((B, 1, 1), 24.0 / 360.0).pin_memory()
    crit = torch.nn.MSELoss()
    args = (host[0], host[1], host[2], host[3], field, cd, fh)

    def eager():
        m.physics_net.zero_grad(set_to_none=True)
        loss = m.place_one_batch(*args, crit, DEFAULT_LOSS_FACTOR, 0, 0, "cuda:0")
        loss.backward()
        return loss.item(), {k: p.grad.clone() for k, p in m.physics_net.named_parameters() if p.grad is not None}

    l0, g0 = eager()
    step = GraphedPlaceOneBatch(m, args, crit, DEFAULT_LOSS_FACTOR, "cuda:0")
    l1 = step().item()
    assert abs(l1 - l0) <= 1e-5 * abs(l0)
    for k, p in m.physics_net.named_parameters():
        if k in g0 and not k.endswith("key_projection.bias"):                # analytically zero gradient: pure round-off noise
            assert H.rel(p.grad, g0[k]) < 1e-4, k
    cd.mul_(0.5)                                                             # refresh a pinned input in place: the replay must see it
    l2 = step().item()
    l3, _ = eager()
    assert abs(l2 - l1) > 1e-3 * abs(l1) and abs(l2 - l3) <= 1e-5 * abs(l3)


def test_prefetched_steps_follow_the_host_buffers():
    """Double-buffered variant: step i computes on the batch prefetched before it, while the next batch is already being copied."""
    from deepphysinet_b200 import InterfacePhysics
    from deepphysinet_b200.config import DEFAULT_LOSS_FACTOR, DEFAULT_OBS_NORM
    from deepphysinet_b200.graphed import PrefetchedPlaceOneBatch
    from oracle import dpn_oracle as O
    obs = {k: dict(v, norm_type="mean_norm", use_norm=True) for k, v in DEFAULT_OBS_NORM.items()}
    torch.manual_seed(0)
    m = InterfacePhysics(H.META_CFG, H.NET_CFG, obs, None, dict(img_size=(145, 257), dx=27000, dy=27000)).cuda()
    g = torch.Generator().manual_seed(6)
    B, N = 2, 300
    pts = [O.synthetic_points(N, g) for _ in range(B)]
    host = [torch.stack([p[i].reshape(-1) for p in pts]).float().pin_memory() for i in range(4)]
    cd = torch.stack([p[4] for p in pts]).float().pin_memory()
    field = torch.randn(B, 159, 2405, generator=g).pin_memory()
    fh = torch.full((B, 1, 1), 24.0 / 360.0).pin_memory()
    crit = torch.nn.MSELoss()
    args = (host[0], host[1], host[2], host[3], field, cd, fh)

    def eager():
        m.physics_net.zero_grad(set_to_none=True)
        loss = m.place_one_batch(*args, crit, DEFAULT_LOSS_FACTOR, 0, 0, "cuda:0")
        loss.backward()
        return loss.item(), {k: p.grad.clone() for k, p in m.physics_net.named_parameters() if p.grad is not None}

    step = PrefetchedPlaceOneBatch(m, args, crit, DEFAULT_LOSS_FACTOR, "cuda:0")
    expect = []
    for scale in (1.0, 0.5, 0.25, 2.0):                                      # four different batches through the two buffer sets
        cd.mul_(scale)
        expect.append(eager())
        step.prefetch()
        step.copied()                                                        # the pinned buffers may be refilled now
        if len(expect) == 1:
            continue                                                         # keep one batch in flight ahead of the compute
        l = step().item()
        want, gw = expect[len(expect) - 2]
        assert abs(l - want) <= 1e-5 * abs(want), (l, want)
    l = step().item()                                                        # drain the last prefetched batch
    want, gw = expect[-1]
    assert abs(l - want) <= 1e-5 * abs(want)
    for k, p in m.physics_net.named_parameters():
        if k in gw and not k.endswith("key_projection.bias"):
            assert H.rel(p.grad, gw[k]) < 1e-4, k
