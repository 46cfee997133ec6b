"""Query-point producer (N2): oracle vs scipy's interpn on CPU; CUDA kernel vs oracle on GPU."""
import numpy as np
import pytest
import torch

from oracle import sampler_oracle as SO


def _field(seed=0, Tt=5, Hc=37, Wc=65):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((Tt, Hc, Wc, 6)).astype(np.float32)


def _queries(n, seed=1, lat=145, lon=257):
    rng = np.random.default_rng(seed)
    x = rng.random(n) * (lon - 1) * 27000.0
    y = rng.random(n) * (lat - 1) * 27000.0
    t = rng.integers(0, 25, n) * 3600.0
    # edge cases: exact nodes, the last node of every axis, the origin
    k = min(n, 4)
    x[:k] = [0.0, (lon - 1) * 27000.0, 4 * 27000.0, (lon - 1) * 27000.0][:k]
    y[:k] = [0.0, (lat - 1) * 27000.0, 8 * 27000.0, 0.0][:k]
    t[:k] = [0.0, 24 * 3600.0, 6 * 3600.0, 24 * 3600.0][:k]
    return x, y, t


def test_oracle_matches_scipy_interpn():
    from scipy.interpolate import interpn
    fld = _field()
    x, y, t = _queries(500)
    got = SO.trilinear(fld, x, y, t)
    # the DataArray of physics_dataset.py:477-479 has dims (y, x, t); interp is point-wise along 'z'
    lat = np.arange(37.0); lon = np.arange(65.0); hrs = np.arange(5.0) * 6.0
    pts = np.stack([y / 27000.0 / 4.0, x / 27000.0 / 4.0, t / 3600.0], 1)
    for v in range(6):
        data = np.transpose(fld[..., v].astype(np.float64), (1, 2, 0))       # [y, x, t]
        ref = interpn((lat, lon, hrs), data, pts, method="linear")
        np.testing.assert_allclose(got[:, v], ref, rtol=1e-12, atol=1e-12)


def _out_of_range_queries():
    # x beyond the last node, negative y, t beyond the last slice (hours passed as if they were seconds x 3600 x 10), NaN, one good point
    x = np.array([257 * 27000.0, 10 * 27000.0, 10 * 27000.0, np.nan, 10 * 27000.0])
    y = np.array([10 * 27000.0, -1.0 * 27000.0, 10 * 27000.0, 10 * 27000.0, 10 * 27000.0])
    t = np.array([3600.0, 3600.0, 240 * 3600.0, 3600.0, 3600.0])
    return x, y, t


def test_oracle_out_of_range_is_nan_like_interpn():
    from scipy.interpolate import interpn
    fld = _field()
    x, y, t = _out_of_range_queries()
    got = SO.trilinear(fld, x, y, t)
    lat = np.arange(37.0); lon = np.arange(65.0); hrs = np.arange(5.0) * 6.0
    pts = np.stack([y / 27000.0 / 4.0, x / 27000.0 / 4.0, t / 3600.0], 1)
    data = np.transpose(fld[..., 0].astype(np.float64), (1, 2, 0))
    ref = interpn((lat, lon, hrs), data, pts, method="linear", bounds_error=False, fill_value=np.nan)
    assert np.array_equal(np.isnan(got[:, 0]), np.isnan(ref))
    assert np.isnan(got[:4]).all() and np.isfinite(got[4]).all()
    np.testing.assert_allclose(got[4, 0], ref[4], rtol=1e-12)


@pytest.mark.gpu
def test_cuda_sampler_out_of_range_is_nan():
    from deepphysinet_b200 import functional as Fn
    fld = _field(seed=3)
    x, y, t = _out_of_range_queries()
    tt = lambda a: torch.tensor(a[None], dtype=torch.float32).cuda()
    cd, _ = Fn.sample_field(torch.from_numpy(fld[None]).cuda(), tt(x), tt(y), tt(t))
    cd = cd[0].cpu().numpy()
    assert np.isnan(cd[:4]).all() and np.isfinite(cd[4]).all()
    np.testing.assert_allclose(cd[4], SO.trilinear(fld, x[4:], y[4:], t[4:])[0], rtol=2e-6, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("B,N", [(1, 1), (2, 1000), (1, 4097)])
def test_cuda_sampler_matches_oracle(B, N):
    from deepphysinet_b200 import functional as Fn
    flds = np.stack([_field(seed=10 + b) for b in range(B)])
    qs = [_queries(N, seed=20 + b) for b in range(B)]
    x = torch.tensor(np.stack([q[0] for q in qs]), dtype=torch.float32)
    y = torch.tensor(np.stack([q[1] for q in qs]), dtype=torch.float32)
    t = torch.tensor(np.stack([q[2] for q in qs]), dtype=torch.float32)
    cd, f = Fn.sample_field(torch.from_numpy(flds).cuda(), x.cuda(), y.cuda(), t.cuda())
    for b in range(B):
        ref = SO.trilinear(flds[b], x[b].double().numpy(), y[b].double().numpy(), t[b].double().numpy())
        np.testing.assert_allclose(cd[b].cpu().numpy(), ref, rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(f[b].cpu().numpy(), SO.coriolis(y[b].double().numpy()), rtol=1e-6, atol=1e-12)
