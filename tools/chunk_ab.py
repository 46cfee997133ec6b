"""Per-call time of the fused operator for several internal pass sizes (DpnShape.chunk = points per sample per pass).
    python tools/chunk_ab.py [mode] [chunk ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deepphysinet_b200 import functional as Fn, testing as T
mode = sys.argv[1] if len(sys.argv) > 1 else "f16x3"
chunks = [int(c) for c in sys.argv[2:]] or [16384, 32768, 65536]
W, pts = T.random_decoder_weights(B=8, N=65536, seed=0, device="cuda")
leaves = [w.clone().requires_grad_(True) for w in W]
orig = Fn._shape
for rep in range(2):
    for ch in chunks:
        Fn._shape = lambda *a, **kw: orig(*a, **{**kw, "chunk": ch})
        step = lambda: Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves), mode=mode)[0]
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(12):
            step()
        e1.record()
        torch.cuda.synchronize()
        print("mode %s chunk %6d: %.2f ms per call" % (mode, ch, e0.elapsed_time(e1) / 12), flush=True)
Fn._shape = orig
