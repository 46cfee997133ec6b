"""GPU (-m gpu): the training-step harness - phase switch at pde_start_step, gradient clipping, Adam step,
per-epoch cosine schedule and the reference checkpoint format - against a hand-rolled PyTorch step on the same losses."""
import copy

import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _model():
    from deepphysinet_b200 import InterfacePhysics
    from deepphysinet_b200.config import DEFAULT_OBS_NORM
    obs = {k: dict(v, norm_type="mean_norm", use_norm=True) for k, v in DEFAULT_OBS_NORM.items()}
    torch.manual_seed(0)
    return InterfacePhysics(H.META_CFG, H.NET_CFG, obs, None, dict(img_size=(145, 257), dx=27000, dy=27000)).cuda()


def _batch(seed=3, n_inter=160, n_margin=96):
    from oracle import dpn_oracle as O
    g = torch.Generator().manual_seed(seed)
    mk = lambda n: [a.reshape(1, -1).float().cuda() for a in O.synthetic_points(n, g)[:4]]
    ix, iy, it_, if_ = mk(n_inter)
    mx, my, mt, mf = mk(n_margin)
    return dict(field_data=torch.randn(1, 159, 2405, generator=g).cuda(), forecast_h=torch.full((1, 1, 1), 24.0 / 360.0).cuda(),
                inter_x=ix, inter_y=iy, inter_t=it_, inter_f=if_, inter_data=(0.5 * torch.randn(1, n_inter, 6, generator=g)).cuda(),
                margin_x=mx, margin_y=my, margin_t=mt, margin_f=mf,
                margin_input_data=(0.5 * torch.randn(1, n_margin, 6, generator=g)).cuda(),
                margin_data=(0.5 * torch.randn(1, n_margin, 6, generator=g)).cuda())


def _model_and_batch():
    return _model(), _batch()


def test_train_step_phases_and_update(tmp_path):
    from deepphysinet_b200.trainer import TrainStep
    m = _model()
    ref = copy.deepcopy(m)
    batch = _batch()
    step = TrainStep(m, pde_start_step=1)
    # step 1: data loss only (global_step 0 < pde_start_step), step 2: data + interior PDE + margin PDE
    p1 = step(batch)
    assert "inter_pde_loss" not in p1 and torch.isfinite(p1["train_loss"])
    p2 = step(batch)
    assert "inter_pde_loss" in p2 and "margin_pde_loss" in p2
    assert abs((p2["margin_loss"] + p2["inter_pde_loss"] + p2["margin_pde_loss"]).item() - p2["train_loss"].item()) <= 1e-6 * abs(p2["train_loss"].item())
    # the same two steps by hand on a copy: Adam(1e-4, wd 1e-4), clip at 2.5e7 (interface_physics.py:505-515)
    opt = torch.optim.Adam(ref.physics_net.parameters(), lr=1e-4, weight_decay=1e-4)
    for with_pde in (False, True):
        tot, _ = ref.training_losses(batch, step.loss_factor, with_pde=with_pde)
        opt.zero_grad(set_to_none=True)
        tot.backward()
        gn = torch.nn.utils.clip_grad_norm_(ref.physics_net.parameters(), max_norm=2.5e7)
        opt.step()
    assert gn.item() > 2.5e7                       # the clip is active at random initialisation
    assert abs(gn.item() - p2["grad_norm"].item()) <= 1e-3 * gn.item()
    # Adam turns every gradient into a step of ~lr whatever its size, so elements whose gradient is round-off noise (the four
    # key_projection.bias tensors have an analytically zero gradient) may move by up to 2 lr in either run;
    # everything else must agree to a small fraction of one step
    for (k, a), (_, b) in zip(m.physics_net.named_parameters(), ref.physics_net.named_parameters()):
        if k.endswith("key_projection.bias"):
            assert (a - b).abs().max().item() <= 4.1e-4, k
            continue
        off = ((a - b).abs() > 1e-5).float().mean().item()
        assert off < 1e-3, (k, off, (a - b).abs().max().item())
    # per-epoch schedule + checkpoint in the reference's format (:53-62), then resume (:64-88)
    lr1 = step.end_epoch(0, str(tmp_path))
    assert lr1 < 1e-4
    state = torch.load(tmp_path / "physics_latest.pth", map_location="cpu", weights_only=False)
    assert set(state) >= {"model", "epoch", "gobal_step", "dx", "dy", "dt", "pred_t_span", "obs_norm_cfg"} and state["gobal_step"] == 2
    m2 = _model()
    step2, epoch = TrainStep.resume(m2, str(tmp_path), pde_start_step=1)
    assert epoch == 1 and step2.global_step == 2
    # the reference rebuilds the scheduler with last_epoch = epoch - 1 (:397): the epoch counter continues (what the learning
    # rate of the first resumed epoch is is torch's business - recent versions restart it from initial_lr)
    assert step2.scheduler.last_epoch == step.scheduler.last_epoch == 1
    for (k, a), (_, b) in zip(m.physics_net.named_parameters(), m2.physics_net.named_parameters()):
        assert torch.equal(a, b), k
