"""BASELINE.json configs[2] and configs[3] on one B200 (the headline bench covers configs[1] / [4]):
  C3  fused decoder + Jacobian + residual + backward microbench, B = 1, N = 2^20 .. 2^24 random query points
  C4  dense-grid continuous-time inference: every node of the 145 x 257 grid at 48 hourly leads, values only
Prints one line per (config, mode).   python tools/config_sweep.py [modes]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from deepphysinet_b200 import InterfacePhysics, functional as Fn, testing as T
from deepphysinet_b200.config import DEFAULT_OBS_NORM
import bench as BN

modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["f16x3", "bf16"]
dev = torch.device("cuda:0")


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for logn in (20, 22, 24):
    N = 1 << logn
    W, pts = T.random_decoder_weights(B=1, N=N, seed=1, device=dev)
    leaves = [w.clone().requires_grad_(True) for w in W]
    for mode in modes:
        ms = timed(lambda: Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves), mode=mode), 3)
        print("C3 microbench  N=2^%d B=1  mode %-6s  %8.2f ms  %6.2f M points/s (fwd + Jacobian + residual + bwd)" % (logn, mode, ms, N / ms / 1e3))
    del W, pts, leaves
    torch.cuda.empty_cache()

# N2: the on-GPU trilinear sampler + Coriolis ("feature gather"): 12 B/point in (x, y, t), 28 B/point out (coord_data, f);
# the coarse field stack (5 x 37 x 65 x 6 floats = 289 KB per sample) stays in L2
for Bs, Ns in ((8, 65536), (1, 1 << 24)):
    g = torch.Generator().manual_seed(2)
    coarse_s = (0.5 * torch.randn(Bs, 5, 37, 65, 6, generator=g)).to(dev)
    xs = (torch.rand(Bs, Ns, generator=g) * 256 * 27000.0).to(dev)
    ys = (torch.rand(Bs, Ns, generator=g) * 144 * 27000.0).to(dev)
    ts = (torch.randint(0, 25, (Bs, Ns), generator=g).float() * 3600.0).to(dev)
    ms = timed(lambda: Fn.sample_field(coarse_s, xs, ys, ts), 10)
    print("N2 sampler  B=%d N=%d  %7.3f ms  %7.1f M points/s  %6.0f GB/s of the 40 B/point it must move (measured copy peak 6536)"
          % (Bs, Ns, ms, Bs * Ns / ms / 1e3, 40.0 * Bs * Ns / ms / 1e6))
    del coarse_s, xs, ys, ts

obs = {k: dict(v, norm_type="mean_norm", use_norm=True) for k, v in DEFAULT_OBS_NORM.items()}
torch.manual_seed(0)
model = InterfacePhysics(BN.META_CFG, BN.NET_CFG, obs, None, dict(img_size=(145, 257), dx=27000, dy=27000)).to(dev)
g = torch.Generator().manual_seed(3)
field = torch.randn(1, 159, 2405, generator=g).to(dev)
coarse = (0.5 * torch.randn(1, 5, 37, 65, 6, generator=g)).to(dev)
fh = torch.full((1, 1, 1), 24.0 / 360.0, device=dev)
leads = list(range(48))
npts = 145 * 257 * len(leads)
for mode in modes:
    model.mode = mode
    out = model.predict_grid(field, coarse, fh, leads)
    assert out.shape == (48, 145, 257, 6) and torch.isfinite(out).all()
    ms = timed(lambda: model.predict_grid(field, coarse, fh, leads), 5)
    print("C4 dense-grid inference 145x257 x 48 leads (%d points, encoder + sampler + decoder values)  mode %-6s  %7.2f ms  %6.2f M points/s"
          % (npts, mode, ms, npts / ms / 1e3))
