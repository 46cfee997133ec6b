#!/bin/bash
# tcgen05 path bring-up: stage-by-stage debug, GPU test suites, quick timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/tc_debug.py 256 1 > gpurun_out/tc_debug.txt 2>&1; echo "tc_debug rc=$?"
head -c 6000 gpurun_out/tc_debug.txt
timeout 900 python -m pytest tests/test_gpu_bf16.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_bf16.txt; cat gpurun_out/pytest_bf16.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_fp32.txt; cat gpurun_out/pytest_fp32.txt
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/timing.txt
import torch
from deepphysinet_b200 import testing as T, functional as Fn
for mode, B, N in (("bf16", 8, 65536), ("bf16", 1, 1 << 20), ("fp32", 8, 65536)):
    W, pts = T.random_decoder_weights(B=B, N=N, seed=0, device="cuda")
    leaves = [w.clone().requires_grad_(True) for w in W]
    def step():
        return Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves), mode=mode)[0]
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%s mode: B=%d N=%d %.2f ms/step -> %.3f Mpoints/s" % (mode, B, N, ms, B * N / ms / 1e3))
PY
