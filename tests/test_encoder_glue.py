"""CPU: the two re-formulations inside the (PyTorch) encoder that exist only for speed on the GPU must be the same function as the
reference layers they replace (model/embed.py:36-49 circular Conv1d; nn.Linear backward): checked in float64, forward and gradients."""
import torch
import torch.nn.functional as F

from deepphysinet_b200 import encoder as E


def test_token_conv_as_gemm_equals_conv1d():
    torch.manual_seed(0)
    m = E._TokenConv(37, 16).double()
    x = torch.randn(3, 11, 37, dtype=torch.float64, requires_grad=True)
    want = m.tokenConv(x.transpose(1, 2)).transpose(1, 2)
    gw, gb, gx = torch.autograd.grad(want.square().sum(), [m.tokenConv.weight, m.tokenConv.bias, x])
    got = m(x)
    hw, hb, hx = torch.autograd.grad(got.square().sum(), [m.tokenConv.weight, m.tokenConv.bias, x])
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
    for a, b in ((hw, gw), (hb, gb), (hx, gx)):
        assert torch.allclose(a, b, rtol=1e-11, atol=1e-11)


def test_batched_grad_linear_equals_linear():
    torch.manual_seed(1)
    x = torch.randn(4, 9, 12, dtype=torch.float64, requires_grad=True)
    w = torch.randn(7, 12, dtype=torch.float64, requires_grad=True)
    b = torch.randn(7, dtype=torch.float64, requires_grad=True)
    want = F.linear(x, w, b)
    g = torch.autograd.grad(want.sin().sum(), [x, w, b])
    got = E._BatchedGradLinear.apply(x, w, b)
    h = torch.autograd.grad(got.sin().sum(), [x, w, b])
    assert torch.equal(got, want)
    for a, c in zip(h, g):
        assert torch.allclose(a, c, rtol=1e-12, atol=1e-12)
    # no gradient wanted for the input: the Function must not compute it
    x2 = x.detach()
    (gw2,) = torch.autograd.grad(E._BatchedGradLinear.apply(x2, w, b).sin().sum(), [w])
    assert torch.allclose(gw2, g[1], rtol=1e-12, atol=1e-12)
