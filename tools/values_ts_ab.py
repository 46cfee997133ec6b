"""Values-only decoder (dpn_decoder_fwd) on B = 1 x 1 788 720 points (the dense-grid inference of BASELINE configs[3]):
time per call and agreement with the default kernel, for the environment this process was started in.
    python tools/values_ts_ab.py [mode]            # default: pass 1 with the activation tile in tensor memory
    DPN_TS=0 python tools/values_ts_ab.py [mode]   # the shared-memory variant
Writes the outputs of a small case to /tmp/values_<tag>.pt so that two runs can be compared bit for bit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from deepphysinet_b200 import functional as Fn, testing as T

mode = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
tag = "smem" if os.environ.get("DPN_TS") == "0" else "tmem"
dev = torch.device("cuda:0")
W, pts = T.random_decoder_weights(B=2, N=1000, seed=5, device=dev)
o = Fn.decoder_values(None, pts["coord_data"], W, xyz=(pts["x"], pts["y"], pts["t"]), mode=mode)
o32 = Fn.decoder_values(None, pts["coord_data"], W, xyz=(pts["x"], pts["y"], pts["t"]), mode="fp32")
print("[%s] %s small case: max |o - o_fp32| / max|o| = %.3e" % (tag, mode, ((o - o32).abs().max() / o32.abs().max()).item()))
torch.save(o.cpu(), "/tmp/values_%s_%s.pt" % (mode, tag))
other = "/tmp/values_%s_%s.pt" % (mode, "tmem" if tag == "smem" else "smem")
if os.path.exists(other):
    d = (torch.load(other) - o.cpu()).abs().max().item()
    print("[%s] max |o_tmem - o_smem| = %.3e" % (tag, d))
N = 145 * 257 * 48
W, pts = T.random_decoder_weights(B=1, N=N, seed=6, device=dev)
f = lambda: Fn.decoder_values(None, pts["coord_data"], W, xyz=(pts["x"], pts["y"], pts["t"]), mode=mode)
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(8):
    e0.record(); f(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
print("[%s] %s values-only, B=1 x %d points: median %.2f ms (min %.2f)  %.1f M points/s" % (tag, mode, N, ts[len(ts) // 2], ts[0], N / ts[len(ts) // 2] / 1e3))
