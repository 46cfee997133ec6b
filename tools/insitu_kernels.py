"""In-situ durations of the library's kernels inside back-to-back dpn_pde_fwd_bwd calls (torch.profiler / CUPTI: no serialisation,
no cache flush - complements the ncu launch list).   python tools/insitu_kernels.py [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from deepphysinet_b200 import functional as Fn, testing as T
from torch.profiler import profile, ProfilerActivity

mode = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
dev = torch.device("cuda:0")
W, pts = T.random_decoder_weights(B=8, N=65536, seed=1, device=dev)
leaves = [w.clone().requires_grad_(True) for w in W]
f = lambda: Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves), mode=mode)
for _ in range(4):
    f()
torch.cuda.synchronize()
n = 4
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        f()
    torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total / 1e3) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print("%s: %d calls, kernel time %.2f ms per call" % (mode, n, tot / n))
for k, c, t in rows[:6]:
    print("   %-60s %4d launches  %8.3f ms per call  mean %.3f ms" % (k[:60], c, t / n, t / c))
