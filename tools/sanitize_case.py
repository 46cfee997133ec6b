"""One small fused call per split-mode kernel set, for compute-sanitizer (tools/gpu_sanitize.sh): N = 300 query points, B = 2,
f16x3 and bf16 (the PL = 2 and PL = 1 instantiations of pass1_ts_kernel, pass2z_kernel, wgrad2_kernel); prints a checksum."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from deepphysinet_b200 import testing as T
W, pts = T.random_decoder_weights(B=2, N=300, seed=0, device="cuda")
for mode in sys.argv[1:] or ["f16x3", "bf16"]:
    out = T.run_library(W, pts, mode=mode, want_fields=False)
    torch.cuda.synchronize()
    print("sanitize_case %s: total %.6e, |dW2| %.6e" % (mode, out["total"].item(), out["grads"][2].norm().item()))
