"""CPU oracle of the query-point producer - TEST INFRASTRUCTURE ONLY.

Restates dataset/physics_dataset.py:442-446,477-486,521-526: the normalised coarse field (37 x 65 nodes at 1 degree,
5 slices 6 h apart) is wrapped in an xarray.DataArray with coordinates (y = lat, x = lon, t = hours) and sampled
point-wise with `DataArray.interp(x=..., y=..., t=...)`, i.e. multi-linear interpolation on a regular grid, which
xarray delegates to scipy.interpolate.interpn(method="linear").  xarray is not installed in this image, so the pin
is scipy's interpn itself (tests/test_sampler.py::test_oracle_matches_scipy_interpn); "parity vs xarray" is
therefore one step removed and says so here.
"""
import numpy as np


def trilinear(coarse, x, y, t, dx=27000.0, dy=27000.0, cells_per_coarse=4.0, t_step=6 * 3600.0):
    """coarse [Tt,Hc,Wc,6]; x, y, t [N] (metres, seconds) -> [N,6] float64."""
    coarse = np.asarray(coarse, dtype=np.float64)
    Tt, Hc, Wc, _ = coarse.shape
    gx = np.asarray(x, dtype=np.float64) / dx / cells_per_coarse
    gy = np.asarray(y, dtype=np.float64) / dy / cells_per_coarse
    gt = np.asarray(t, dtype=np.float64) / t_step
    ix = np.clip(np.floor(gx).astype(int), 0, Wc - 2)
    iy = np.clip(np.floor(gy).astype(int), 0, Hc - 2)
    it = np.clip(np.floor(gt).astype(int), 0, Tt - 2)
    wx, wy, wt = (gx - ix)[:, None], (gy - iy)[:, None], (gt - it)[:, None]
    out = np.zeros((gx.shape[0], 6))
    for dt_ in (0, 1):
        for dy_ in (0, 1):
            for dx_ in (0, 1):
                w = (wt if dt_ else 1 - wt) * (wy if dy_ else 1 - wy) * (wx if dx_ else 1 - wx)
                out += w * coarse[it + dt_, iy + dy_, ix + dx_]
    # outside the stack: NaN, as DataArray.interp / interpn(bounds_error=False) fill (no silent extrapolation)
    eps = 1e-9
    inside = ((gx >= -eps) & (gx <= Wc - 1 + eps) & (gy >= -eps) & (gy <= Hc - 1 + eps) & (gt >= -eps) & (gt <= Tt - 1 + eps))
    out[~inside] = np.nan
    return out


def coriolis(y, dy=27000.0, begin_lat=18.0, deg_per_cell=0.25, omega=7.29e-5):
    """dataset/physics_dataset.py:521-526 with lat = begin_lat + y_cells * 0.25 (:445)."""
    lat = begin_lat + np.asarray(y, dtype=np.float64) / dy * deg_per_cell
    return 2 * omega * np.sin(lat / 180 * np.pi)
