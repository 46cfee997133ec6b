"""Where the host time of one fused call goes (diagnostic): cProfile over synchronised steps."""
import cProfile, pstats, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from deepphysinet_b200 import functional as Fn, testing as T
W, pts = T.random_decoder_weights(B=8, N=65536, seed=0, device="cuda")
leaves = [w.clone().requires_grad_(True) for w in W]
def step():
    r = Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves))[0]
    torch.cuda.synchronize()
    return r
for _ in range(3): step()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
