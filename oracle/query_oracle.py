"""CPU oracle of the on-GPU query-point generator - TEST INFRASTRUCTURE ONLY.

The reference draws its query points with numpy's global MT19937 stream inside DataLoader workers
(dataset/physics_dataset.py:442-446 interior, :334-338 margin): continuous x, y uniform over the fine grid and an integer hour
in [0, 25) for interior points, integer grid nodes for margin points.  The GPU generator (csrc/dpn_sampler.cu:query_kernel)
keeps the distributions and replaces the stream by the counter-based Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel
random numbers: as easy as 1, 2, 3", SC'11; the generator behind cuRAND's Philox and torch.cuda's default RNG), restated here
word for word in numpy.  Pinned by the known-answer vectors of the Random123 distribution (tests/test_query_generator.py).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter [..., 4], key [..., 2] (uint32 values) -> [..., 4] uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k = [np.asarray(key[..., i], dtype=np.uint64) for i in range(2)]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return np.stack(c, axis=-1).astype(np.uint32)


def generate(B, N, seed, offset=0, on_grid=False, lat_size=145, lon_size=257, t_steps=25, dx=27000.0, dy=27000.0, dt=3600.0):
    """x, y, t [B,N] float32, bit-identical to dpn_generate_queries."""
    idx = np.arange(N, dtype=np.uint64)[None, :] + np.uint64(offset)
    ctr = np.zeros((B, N, 4), dtype=np.uint64)
    ctr[..., 0] = idx & MASK
    ctr[..., 1] = idx >> np.uint64(32)
    ctr[..., 2] = np.arange(B, dtype=np.uint64)[:, None]
    key = np.zeros((B, N, 2), dtype=np.uint64)
    key[..., 0] = np.uint64(seed) & MASK
    key[..., 1] = np.uint64(seed) >> np.uint64(32)
    r = philox4x32_10(ctr, key).astype(np.uint64)
    if on_grid:
        x = ((r[..., 0] * np.uint64(lon_size)) >> np.uint64(32)).astype(np.float64) * dx
        y = ((r[..., 1] * np.uint64(lat_size)) >> np.uint64(32)).astype(np.float64) * dy
    else:
        x = (r[..., 0] >> np.uint64(8)).astype(np.float64) * (1.0 / 16777216.0) * float(lon_size - 1) * dx
        y = (r[..., 1] >> np.uint64(8)).astype(np.float64) * (1.0 / 16777216.0) * float(lat_size - 1) * dy
    t = ((r[..., 2] * np.uint64(t_steps)) >> np.uint64(32)).astype(np.float64) * dt
    return x.astype(np.float32), y.astype(np.float32), t.astype(np.float32)
