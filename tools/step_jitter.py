"""Per-step device times of the fused operator + nvidia-smi clocks/power while it runs (diagnostic)."""
import os, subprocess, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
from deepphysinet_b200 import functional as Fn, testing as T
mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 60
B, N = 8, 65536
W, pts = T.random_decoder_weights(B=B, N=N, seed=0, device="cuda")
leaves = [w.clone().requires_grad_(True) for w in W]
smi = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,power.limit,temperature.gpu,utilization.gpu,clocks_event_reasons.active",
                        "--format=csv,noheader,nounits", "-lms", "50"], stdout=open("/tmp/smi.csv", "w"))
def step():
    return Fn.pde_residual(pts["x"], pts["y"], pts["t"], pts["f"], pts["coord_data"], Fn.DecoderWeights(*leaves), mode=mode)[0]
for _ in range(3): step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
t0 = time.perf_counter()
ev[0].record()
for i in range(steps):
    step(); ev[i + 1].record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
print("mode %s: per-step ms: first %s ... min %.2f median %.2f max %.2f mean %.2f | cpu issue %.1f ms, wall %.1f ms" %
      (mode, ["%.2f" % m for m in ms[:6]], min(ms), sorted(ms)[len(ms)//2], max(ms), sum(ms)/len(ms), (t1-t0)*1e3, (t2-t0)*1e3))
print("every 5th:", ["%.1f" % m for m in ms[::5]])
smi.terminate(); smi.wait()
cpu = []
for _ in range(10):
    torch.cuda.synchronize(); a = time.perf_counter(); step(); cpu.append((time.perf_counter() - a) * 1e3)
torch.cuda.synchronize()
print("CPU issue time of one call on an idle queue (ms): min %.2f median %.2f" % (min(cpu), sorted(cpu)[5]))
lines = open("/tmp/smi.csv").read().strip().splitlines()
print("smi samples %d; first/last few:" % len(lines)); print("\n".join(lines[:3] + ["..."] + lines[-12:]))
