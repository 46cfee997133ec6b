"""CPU: the batched hyper-network (PhysicsNet.decoder_weights: one GEMM for the 12 Linears of variable_net.py:57-65) against the
per-net structure of the reference, values and gradients."""
import torch

from tests import helpers as H


def test_batched_hypernet_equals_per_net_generation():
    from deepphysinet_b200.physics_net import PhysicsNet
    torch.manual_seed(3)
    net = PhysicsNet(H.META_CFG, H.NET_CFG).double()
    field = torch.randn(2, 159, 2405, dtype=torch.float64)
    fh = torch.tensor([[[24.0 / 360.0]], [[48.0 / 360.0]]], dtype=torch.float64)
    a = net.decoder_weights(field, fh)
    b = net.decoder_weights_per_net(field, fh)
    for name, x, y in zip(a._fields, a, b):
        assert x.shape == y.shape, name
        assert torch.allclose(x, y, rtol=1e-12, atol=1e-14), name
    # gradients into every parameter (hyper-network, encoder, static decoder) through a random linear functional
    torch.manual_seed(4)
    probes = [torch.randn_like(x) for x in a]
    grads = []
    for fn in (net.decoder_weights, net.decoder_weights_per_net):
        net.zero_grad(set_to_none=True)
        sum((w * p).sum() for w, p in zip(fn(field, fh), probes)).backward()
        grads.append({n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys()
    gmax = max(g.abs().max() for g in grads[1].values())
    for n in grads[0]:
        # key_projection.bias has an analytically zero gradient (softmax shift invariance): pure round-off, judged on the global scale
        bound = 1e-10 * grads[1][n].abs().max() + 1e-13 * gmax
        assert (grads[0][n] - grads[1][n]).abs().max() <= bound, n
