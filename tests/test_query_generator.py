"""Query-point generator (N2): the numpy Philox4x32-10 oracle against the Random123 known-answer vectors and the reference's
distributions on CPU; the CUDA kernel against the oracle (bit-exact) and, fused with the sampler, against the sampler oracle on GPU."""
import numpy as np
import pytest
import torch

from oracle import query_oracle as QO
from oracle import sampler_oracle as SO


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = QO.philox4x32_10(np.array(ctr, dtype=np.uint64), np.array(key, dtype=np.uint64))
        assert tuple(int(v) for v in got) == want, (ctr, [hex(int(v)) for v in got])


def test_distributions_match_the_reference_draws():
    """physics_dataset.py:442-446: x in [0, (W-1) dx), y in [0, (H-1) dy) continuous, t / 3600 integer in [0, 25); :334-338: grid nodes."""
    x, y, t = QO.generate(2, 200000, seed=7)
    assert x.min() >= 0 and x.max() < 256 * 27000.0 and y.min() >= 0 and y.max() < 144 * 27000.0
    assert abs(x.mean() / (256 * 27000.0) - 0.5) < 5e-3 and abs(y.mean() / (144 * 27000.0) - 0.5) < 5e-3
    h = t / 3600.0
    assert np.array_equal(h, np.round(h)) and h.min() == 0 and h.max() == 24
    counts = np.bincount(h.astype(int).ravel(), minlength=25)
    assert counts.min() > 0.9 * counts.mean() and counts.max() < 1.1 * counts.mean()
    assert abs(np.corrcoef(x[0], y[0])[0, 1]) < 0.01 and abs(np.corrcoef(x[0], x[1])[0, 1]) < 0.01
    gx, gy, gt = QO.generate(1, 100000, seed=7, on_grid=True)
    ix, iy = gx / 27000.0, gy / 27000.0
    assert np.array_equal(ix, np.round(ix)) and ix.min() == 0 and ix.max() == 256 and iy.min() == 0 and iy.max() == 144
    # successive steps (offset) draw disjoint blocks of the same stream
    a = QO.generate(1, 64, seed=3, offset=0)[0]
    b = QO.generate(1, 32, seed=3, offset=32)[0]
    assert np.array_equal(a[:, 32:], b)


@pytest.mark.gpu
@pytest.mark.parametrize("on_grid", [False, True])
def test_cuda_generator_is_bit_exact(on_grid):
    from deepphysinet_b200 import functional as Fn
    B, N = 3, 5001
    x, y, t = Fn.generate_queries(B, N, seed=0x1234567890ABCDEF, offset=(1 << 33) + 5, on_grid=on_grid)
    ox, oy, ot = QO.generate(B, N, seed=0x1234567890ABCDEF, offset=(1 << 33) + 5, on_grid=on_grid)
    assert np.array_equal(x.cpu().numpy(), ox) and np.array_equal(y.cpu().numpy(), oy) and np.array_equal(t.cpu().numpy(), ot)


@pytest.mark.gpu
def test_cuda_generator_fused_with_sampler():
    from deepphysinet_b200 import functional as Fn
    B, N = 2, 4097
    rng = np.random.default_rng(5)
    fld = rng.standard_normal((B, 5, 37, 65, 6)).astype(np.float32)
    x, y, t, cd, f = Fn.generate_queries(B, N, seed=11, coarse=torch.from_numpy(fld).cuda())
    ox, oy, ot = QO.generate(B, N, seed=11)
    assert np.array_equal(x.cpu().numpy(), ox) and np.array_equal(t.cpu().numpy(), ot)
    for b in range(B):
        ref = SO.trilinear(fld[b], ox[b].astype(np.float64), oy[b].astype(np.float64), ot[b].astype(np.float64))
        np.testing.assert_allclose(cd[b].cpu().numpy(), ref, rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(f[b].cpu().numpy(), SO.coriolis(oy[b].astype(np.float64)), rtol=1e-6, atol=1e-12)
    # the fused producer equals generator followed by dpn_sample_field
    cd2, f2 = Fn.sample_field(torch.from_numpy(fld).cuda(), x, y, t)
    assert torch.equal(cd, cd2) and torch.equal(f, f2)
