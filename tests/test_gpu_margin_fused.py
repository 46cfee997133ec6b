"""GPU (-m gpu): the supervised margin loss fused into the PDE call (dpn_pde_margin_fwd_bwd) against the two
separate calls the reference structure implies (values-only decoder + WeightSmoothL1Loss, interface_physics.py:464-474, and
place_one_batch on the same points, :489-496) - loss values, normalised outputs and every weight gradient."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [("fp32", 2e-5), ("f16x3", 1e-4), ("bf16", 0.2)])
def test_fused_margin_equals_separate_calls(mode, tol):
    from deepphysinet_b200 import functional as Fn, testing as T
    B, N = 2, 700
    W, pts = T.random_decoder_weights(B=B, N=N, seed=31, device="cuda")
    g = torch.Generator().manual_seed(5)
    # targets around the skip connection: both branches of smooth_l1 (|d| < beta and beyond) are exercised
    target = (pts["coord_data"].cpu() + 0.2 * torch.randn(B, N, 6, generator=g)).cuda()
    beta, factor = 0.1, 1.0e6
    a = [w.detach().clone().requires_grad_(True) for w in W]
    total, terms, mloss, o = Fn.pde_margin_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], target,
                                                    Fn.DecoderWeights(*a), beta=beta, factor=factor, mode=mode)
    total.backward()
    b = [w.detach().clone().requires_grad_(True) for w in W]
    Wb = Fn.DecoderWeights(*b)
    tot_pde, terms2 = Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Wb, mode=mode)
    o2 = Fn.decoder_values(None, pts["coord_data"], Wb, xyz=(pts["x"], pts["y"], pts["t"]), mode=mode)
    ml = torch.nn.functional.smooth_l1_loss(o2, target, beta=beta, reduction="none").mean(dim=(1, 2)) * factor
    (tot_pde + ml.mean()).backward()
    assert torch.allclose(terms, terms2, rtol=1e-9)
    assert ((o - o2.detach()).abs().max() / o2.abs().max()).item() < (1e-6 if mode != "bf16" else 1e-2)
    assert torch.allclose(mloss, ml.detach().double(), rtol=max(tol, 1e-5))
    frac_quad = ((o2.detach() - target).abs() < beta).float().mean().item()
    assert 0.1 < frac_quad < 0.9, frac_quad
    for name, ga, gb in zip(Fn.DecoderWeights._fields, a, b):
        rel = ((ga.grad - gb.grad).norm() / gb.grad.norm().clamp_min(1e-30)).item()
        assert rel < tol, (name, rel)


def test_training_losses_use_the_fused_margin_path():
    """InterfacePhysics.training_losses: the margin points go through ONE library call for data loss + PDE loss; the result equals
    the separate-call composition of interface_physics.py:464-501."""
    from tests.test_gpu_trainer import _model_and_batch
    from deepphysinet_b200 import functional as Fn
    from deepphysinet_b200.config import DEFAULT_LOSS_FACTOR
    model, batch = _model_and_batch()
    lf = dict(DEFAULT_LOSS_FACTOR, margin_factor=1.0e6)
    model.physics_net.zero_grad(set_to_none=True)
    total, parts = model.training_losses(batch, lf, with_pde=True)
    total.backward()
    ga = {n: p.grad.clone() for n, p in model.physics_net.named_parameters() if p.grad is not None}
    model.physics_net.zero_grad(set_to_none=True)
    total2, parts2 = model.training_losses(batch, lf, with_pde=True, fuse_margin=False)
    total2.backward()
    assert abs(total.item() - total2.item()) <= 1e-5 * abs(total2.item())
    assert abs(parts["margin_loss"].item() - parts2["margin_loss"].item()) <= 1e-5 * abs(parts2["margin_loss"].item())
    gmax = max(p.grad.abs().max() for p in model.physics_net.parameters() if p.grad is not None)
    for n, p in model.physics_net.named_parameters():
        if p.grad is None:
            continue
        err = (ga[n] - p.grad).norm()
        assert err <= 2e-4 * p.grad.norm() + 1e-7 * gmax, (n, err.item(), p.grad.norm().item())
