"""GPU (-m gpu): the values-only entry points dpn_decoder_fwd / dpn_decoder_bwd (PhysicsNet.forward surface,
dense-grid inference, supervised margin loss) from raw coordinates, in both arithmetic modes."""
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _oracle_values(W, pts, consts):
    from deepphysinet_b200 import functional as Fn
    from oracle import dpn_oracle as O
    names = Fn.DecoderWeights._fields
    B = W.W1.shape[0]
    leaves = [w.detach().double().cpu().requires_grad_(True) for w in W]
    outs = []
    for b in range(B):
        Wb = {n: (l[b] if n in ("W1", "b1", "W2", "b2", "e") else l) for n, l in zip(names, leaves)}
        col = lambda k: pts[k][b].double().cpu().reshape(-1, 1)
        pe = O.encoding_coord(col("x"), col("y"), col("t"), consts.dx, consts.dy, consts.lat_size, consts.lon_size,
                              consts.pred_t_span)
        outs.append(torch.cat(O.decode_generated(pe, pts["coord_data"][b].double().cpu(), Wb), 1))
    return torch.stack(outs), leaves


@pytest.mark.parametrize("mode,tol_v,tol_g", [("fp32", 1e-4, 1e-4), ("f16x3", 1e-5, 1e-4), ("f16x3a", 1e-5, 1e-4), ("bf16x3", 1e-5, 2e-3), ("bf16", 2e-2, 5e-2)])
def test_decoder_values_and_backward_from_xyz(mode, tol_v, tol_g):
    from deepphysinet_b200 import functional as Fn, testing as T
    from deepphysinet_b200.config import PhysicsConsts
    consts = PhysicsConsts()
    W, pts = T.random_decoder_weights(B=2, N=333, seed=17, device="cuda")
    ref, leaves64 = _oracle_values(W, pts, consts)
    wts = torch.linspace(-1.0, 1.0, ref.numel(), dtype=torch.float64).reshape(ref.shape)
    (ref * wts).sum().backward()

    leaves = [w.detach().clone().requires_grad_(True) for w in W]
    o = Fn.decoder_values(None, pts["coord_data"], Fn.DecoderWeights(*leaves), xyz=(pts["x"], pts["y"], pts["t"]),
                          consts=consts, mode=mode)
    assert o.shape == ref.shape
    assert H.rel(o.detach().cpu(), ref.detach()) < tol_v
    (o * wts.float().cuda()).sum().backward()
    for n, l, l64 in zip(Fn.DecoderWeights._fields, leaves, leaves64):
        assert H.rel(l.grad.cpu(), l64.grad) < tol_g, (n, H.rel(l.grad.cpu(), l64.grad))


def test_dense_grid_inference_order_and_inverse_norm():
    """interface_physics.py:538-563: all grid nodes of one time slice, x-major, values only, with_clip = False."""
    from deepphysinet_b200 import functional as Fn, testing as T
    from deepphysinet_b200.config import PhysicsConsts
    consts = PhysicsConsts(lat_size=9, lon_size=11, with_clip=False)
    W, _ = T.random_decoder_weights(B=1, N=8, seed=23, device="cuda", consts=consts)
    xs, ys = torch.meshgrid(torch.arange(11.0), torch.arange(9.0), indexing="ij")       # x-major node list (:541-545)
    x = (xs.reshape(1, -1) * consts.dx).cuda()
    y = (ys.reshape(1, -1) * consts.dy).cuda()
    t = torch.full_like(x, 3 * 3600.0)
    cd = 0.3 * torch.randn(1, x.shape[1], 6, generator=torch.Generator().manual_seed(5)).cuda()
    pts = dict(x=x, y=y, t=t, coord_data=cd)
    ref, _ = _oracle_values(W, pts, consts)
    o = Fn.decoder_values(None, cd, W, xyz=(x, y, t), consts=consts, mode="fp32")
    assert H.rel(o.cpu(), ref.detach()) < 1e-4


def _interface(mode="fp32", img=(145, 257)):
    from deepphysinet_b200 import InterfacePhysics
    from deepphysinet_b200.config import DEFAULT_OBS_NORM
    obs = {k: dict(v, norm_type="mean_norm", use_norm=True) for k, v in DEFAULT_OBS_NORM.items()}
    torch.manual_seed(0)
    m = InterfacePhysics(H.META_CFG, H.NET_CFG, obs, None, dict(img_size=img, dx=27000, dy=27000)).double().cuda()
    m.mode = mode
    return m


def test_training_losses_compose_like_the_reference_loop():
    """interface_physics.py:464-501: margin SmoothL1 x 1e6 + interior PDE + margin PDE, one encoder pass."""
    from deepphysinet_b200.config import DEFAULT_LOSS_FACTOR
    from oracle import dpn_oracle as O
    m = _interface()
    g = torch.Generator().manual_seed(3)
    dt = torch.float64
    mk = lambda n: [a.reshape(1, -1).cuda() for a in O.synthetic_points(n, g, dtype=dt)[:4]]
    ix, iy, it_, if_ = mk(96)
    mx, my, mt, mf = mk(64)
    batch = dict(field_data=torch.randn(1, 159, 2405, generator=g, dtype=dt).cuda(),
                 forecast_h=torch.full((1, 1, 1), 24.0 / 360.0, dtype=dt).cuda(),
                 inter_x=ix, inter_y=iy, inter_t=it_, inter_f=if_, inter_data=(0.5 * torch.randn(1, 96, 6, generator=g, dtype=dt)).cuda(),
                 margin_x=mx, margin_y=my, margin_t=mt, margin_f=mf,
                 margin_input_data=(0.5 * torch.randn(1, 64, 6, generator=g, dtype=dt)).cuda(),
                 margin_data=(0.5 * torch.randn(1, 64, 6, generator=g, dtype=dt)).cuda())
    lf = dict(DEFAULT_LOSS_FACTOR, margin_factor=1.0e6)
    total, parts = m.training_losses(batch, lf)
    total.backward()
    # oracle: same composition on the CPU in fp64
    net = m.physics_net
    cpu = {k: v.detach().cpu() for k, v in batch.items()}
    params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.named_parameters()}
    import torch.func as tf
    meta = tf.functional_call(net.meta_net.cpu(), {k[len("meta_net."):]: v for k, v in params.items() if k.startswith("meta_net.")},
                              (cpu["field_data"], cpu["forecast_h"]))
    net.cuda()
    sp = O.split_params(params)
    col = lambda a: a.reshape(-1, 1)
    pe = O.encoding_coord(col(cpu["margin_x"]), col(cpu["margin_y"]), col(cpu["margin_t"]), 27000.0, 27000.0, 145, 257, 86400.0)
    vals = torch.cat(O.physics_net_decode(meta, pe, cpu["margin_input_data"][0], cpu["forecast_h"], sp), 1)
    ref = torch.nn.functional.smooth_l1_loss(vals, cpu["margin_data"][0], beta=0.1, reduction="none").mean() * 1.0e6
    for pre, dk in (("inter", "inter_data"), ("margin", "margin_input_data")):
        tot, _ = O.place_one_batch(col(cpu[pre + "_x"]), col(cpu[pre + "_y"]), col(cpu[pre + "_t"]), col(cpu[pre + "_f"]),
                                   cpu[dk][0], cpu["forecast_h"], meta, sp)
        ref = ref + tot
    ref.backward()
    assert abs(total.item() - ref.item()) / abs(ref.item()) < 1e-4
    gn = sum(p.grad.norm().item() ** 2 for p in params.values() if p.grad is not None) ** 0.5
    for k, p in net.named_parameters():
        r = params[k].grad
        if r is None:
            continue
        err = (p.grad.cpu() - r).norm().item()
        assert err < 1e-3 * max(r.norm().item(), 1e-6 * gn), (k, err, r.norm().item())


def test_predict_grid_matches_oracle_pipeline():
    import numpy as np
    from oracle import dpn_oracle as O, sampler_oracle as SO
    m = _interface(img=(13, 17))
    g = torch.Generator().manual_seed(8)
    field = torch.randn(1, 159, 2405, generator=g, dtype=torch.float64).cuda()
    fh = torch.full((1, 1, 1), 24.0 / 360.0, dtype=torch.float64).cuda()
    coarse = torch.randn(1, 5, 4, 5, 6, generator=g)            # covers a 13 x 17 fine grid (3 x 4 coarse cells)
    out = m.predict_grid(field, coarse.cuda(), fh, time_ids=[0, 5, 24])
    assert out.shape == (3, 13, 17, 6)
    net = m.physics_net
    meta = net.meta_net(field, fh).detach().cpu()
    sp = O.split_params({k: v.detach().cpu() for k, v in net.named_parameters()})
    for ti, tid in enumerate([0, 5, 24]):
        ys, xs = np.meshgrid(np.arange(13.0), np.arange(17.0), indexing="ij")
        x = xs.reshape(-1) * 27000.0; y = ys.reshape(-1) * 27000.0; t = np.full_like(x, tid * 3600.0)
        cd = torch.from_numpy(SO.trilinear(coarse[0].numpy(), x, y, t))
        col = lambda a: torch.from_numpy(a).reshape(-1, 1)
        pe = O.encoding_coord(col(x), col(y), col(t), 27000.0, 27000.0, 13, 17, 86400.0)
        vals = torch.cat(O.inverse_norm(O.physics_net_decode(meta, pe, cd, fh.cpu(), sp), with_clip=False), 1)
        got = out[ti].reshape(-1, 6).cpu().double()
        assert H.rel(got, vals.detach()) < 1e-4


@pytest.mark.parametrize("K", [6, 1])
def test_pre_encoded_surface_honours_mode(K):
    """PhysicsNet.forward / VariableNet.forward hand the library an already encoded coordinate [N,192] (reference
    model/physics_net.py:41-55, variable_net.py:49) and, for one net, an explicit `ref`.  Round 1 silently ran this surface on the
    CUDA cores whatever `mode` said; now the requested mode runs: both modes agree with the oracle within their own tolerance, the
    results differ in the last bits, and the kernel sequences differ (dpn_last_launch_count)."""
    from deepphysinet_b200 import functional as Fn, testing as T, _native as N
    from deepphysinet_b200.config import PhysicsConsts
    from oracle import dpn_oracle as O
    consts = PhysicsConsts()
    W, pts = T.random_decoder_weights(B=1, N=500, seed=5, device="cuda")
    if K == 1:                                                     # one net with its own skip input, like VariableNet.forward
        W = Fn.DecoderWeights(*[(w[:, 2:3] if n in ("W1", "b1", "W2", "b2", "e") else w[2:3]).clone() for n, w in zip(Fn.DecoderWeights._fields, W)])
    col = lambda k: pts[k][0].double().cpu().reshape(-1, 1)
    pe64 = O.encoding_coord(col("x"), col("y"), col("t"), consts.dx, consts.dy, consts.lat_size, consts.lon_size, consts.pred_t_span)
    cd = pts["coord_data"][0]
    ref_in = cd[:, 2:3].contiguous() if K == 1 else None
    names = Fn.DecoderWeights._fields
    Wb = {n: (w[0] if n in ("W1", "b1", "W2", "b2", "e") else w).double().cpu() for n, w in zip(names, W)}
    if K == 1:                                                     # variable_net.py:67-87 with the explicit skip input
        want = O.decoder_net(pe64, cd.double().cpu(), ref_in.double().cpu(), Wb["W1"][0], Wb["b1"][0], Wb["W2"][0], Wb["b2"][0],
                             Wb["e"][0], O._net_params(Wb, 0))
    else:
        want = torch.cat(O.decode_generated(pe64, cd.double().cpu(), Wb), 1)
    pe32 = pe64.float().cuda()
    got, launches = {}, {}
    for mode in ("fp32", "f16x3"):
        got[mode] = Fn.decoder_values(pe32, cd, W, ref=ref_in, mode=mode).detach().cpu()
        launches[mode] = N.lib().dpn_last_launch_count()
        assert got[mode].shape == (500, K)
        assert H.rel(got[mode], want) < (1e-5 if mode == "fp32" else 2e-5), (mode, H.rel(got[mode], want))
    assert H.rel(got["f16x3"], got["fp32"]) < 2e-5
    assert not torch.equal(got["f16x3"], got["fp32"])              # different arithmetic actually ran
    assert launches["fp32"] != launches["f16x3"], launches         # CUDA-core kernel sequence vs tcgen05 kernel sequence


def test_cross_first_mode_only_changes_the_masked_paths():
    """DPN_MODE_F16X3A re-orders the accumulation of G1 - G3 only when a Jacobian / backward pass follows (the values are continuous
    in the ReLU masks): the values-only forward is bit-identical to f16x3, the PDE call is not - and closer to the oracle in values."""
    from deepphysinet_b200 import functional as Fn, testing as T
    W, pts = T.random_decoder_weights(B=2, N=700, seed=23, device="cuda")
    xyz = (pts["x"], pts["y"], pts["t"])
    o1 = Fn.decoder_values(None, pts["coord_data"], W, xyz=xyz, mode="f16x3")
    o2 = Fn.decoder_values(None, pts["coord_data"], W, xyz=xyz, mode="f16x3a")
    assert torch.equal(o1, o2)
    ref = T.oracle_reference(W, pts)
    a = T.run_library(W, pts, mode="f16x3")
    b = T.run_library(W, pts, mode="f16x3a")
    assert not torch.equal(a["vals"], b["vals"])
    ea = max(T._rel(a["vals"][..., k], ref["vals"][..., k]) for k in range(6))
    eb = max(T._rel(b["vals"][..., k], ref["vals"][..., k]) for k in range(6))
    print("values vs fp64 oracle: f16x3 %.2e, f16x3a %.2e" % (ea, eb))
    assert eb < 3e-7 and eb < ea
