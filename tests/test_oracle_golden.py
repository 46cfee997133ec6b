"""CPU: pins oracle/dpn_oracle.py (and the closed-form twin, and the package's PyTorch encoder /
parameter layout) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import closed_form as CF
from oracle import dpn_oracle as O
from tests import helpers as H


@pytest.mark.parametrize("name", H.CASES)
def test_oracle_matches_reference_golden(name):
    case = H.load_case(name)
    net = H.build_model(case)
    # same seed => same parameters as the reference constructors produced
    chk = np.array([[v.double().sum().item(), v.double().abs().sum().item()] for v in net.state_dict().values()])
    np.testing.assert_allclose(chk, case["param_checksum"], rtol=1e-12, atol=1e-12)

    x, y, t, f, cd, field, fh = H.case_inputs(name)
    np.testing.assert_array_equal(x.numpy(), case["x"])
    np.testing.assert_array_equal(cd.numpy(), case["coord_data"])
    meta = net.meta_net(field, fh)
    np.testing.assert_allclose(meta.detach().flatten()[::997].numpy(), case["meta_sample"], rtol=1e-9, atol=1e-11)

    params = O.split_params(dict(net.named_parameters()))
    total, terms, vals, jac = O.place_one_batch(x, y, t, f, cd, fh, meta, params, return_fields=True, **H.geometry(case))
    np.testing.assert_allclose(total.item(), case["total64"], rtol=1e-10)
    np.testing.assert_allclose(torch.stack(terms).detach().numpy(), case["terms64"], rtol=1e-9)
    np.testing.assert_allclose(vals.numpy(), case["vals64"], rtol=1e-10, atol=1e-12)
    assert H.rel(jac, case["jac64"]) < 1e-10

    total.backward()
    grads = dict(net.named_parameters())
    gnorm = float(np.sqrt((case["grad_norm64"] ** 2).sum()))
    for k, n64 in zip(case["grad_names"], case["grad_norm64"]):
        g = grads[str(k)].grad
        ref = torch.from_numpy(case["g64/" + str(k)])
        got = g if g.numel() <= 4096 else g.flatten()[::997]
        err = (got.reshape(ref.shape) - ref).norm().item()
        # tensors whose gradient is analytically zero (key_projection.bias: softmax shift invariance) hold noise only
        assert err <= max(1e-8 * ref.norm().item(), 1e-12 * gnorm), (k, err, ref.norm().item())
        np.testing.assert_allclose(g.norm().item(), n64, rtol=1e-8, atol=1e-12 * gnorm)


@pytest.mark.parametrize("name", H.CASES)
def test_closed_form_matches_oracle(name):
    """The algorithm the CUDA kernels execute (one value row + one reverse sweep + one combined tangent row
    per point) is the same function as the reference's double-backward graph."""
    case = H.load_case(name)
    net = H.build_model(case)
    x, y, t, f, cd, field, fh = H.case_inputs(name)
    W = H.leaf_weights(net, field, fh)
    geo = H.geometry(case)
    total, terms, vals, jac = O.place_generated(x, y, t, f, cd, W, return_fields=True, **geo)
    total.backward()
    losses, G, vals2, jac2 = CF.pde_fwd_bwd(x, y, t, f, cd, {k: v.detach() for k, v in W.items()},
                                            dx=geo["dx"], dy=geo["dy"], lat_size=geo["lat_size"],
                                            lon_size=geo["lon_size"], t_span=geo["pred_t_span"], with_clip=geo["with_clip"])
    np.testing.assert_allclose(losses.numpy(), torch.stack(terms).detach().numpy(), rtol=1e-10)
    np.testing.assert_allclose(losses.numpy(), case["terms64"], rtol=1e-9)
    assert H.rel(vals2, vals) < 1e-12 and H.rel(jac2, jac) < 1e-11
    for k in G:
        assert H.rel(G[k], W[k].grad) < 1e-10, k


def test_reference_fp32_noise_floor_is_recorded():
    """The 1e-4 target sits near the reference's own fp32-vs-fp64 discrepancy; the fixtures carry that yardstick."""
    case = H.load_case("inter_0p25_n192")
    rel = case["grad_ref32_vs_ref64"] / np.maximum(case["grad_norm64"], 1e-300)
    big = case["grad_norm64"] > 1e-9 * np.sqrt((case["grad_norm64"] ** 2).sum())
    assert np.isfinite(rel[big]).all() and rel[big].max() < 1e-2
