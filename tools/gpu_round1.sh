#!/bin/bash
# first GPU visit: fp32 parity suite + tcgen05 plumbing probe + fp32 timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for m in 0 1 2 3 4 5 6 7; do timeout 60 tools/bin/umma_probe $m; done > gpurun_out/probe.txt 2>&1
cat gpurun_out/probe.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/fp32_timing.txt
import torch, time
from deepphysinet_b200 import testing as T, functional as Fn
W, pts = T.random_decoder_weights(B=1, N=65536, seed=0, device="cuda")
leaves = [w.clone().requires_grad_(True) for w in W]
def step():
    tot, terms = Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves), mode="fp32")
    return tot
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print("fp32 mode: N=65536 B=1 %.2f ms/step -> %.3f Mpoints/s" % (ms, 65536 / ms / 1e3))
PY
