"""Short training run of the step harness (deepphysinet_b200.TrainStep) with the reference's step sizes - 4096 interior and
20480 margin points, B = 1 - from the same initialisation in two arithmetic modes; prints the loss trajectory of each.
    python tools/train_demo.py [steps] [modes]"""
import copy, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
ge.build()
import bench as BN
from deepphysinet_b200 import InterfacePhysics, TrainStep
from deepphysinet_b200.config import DEFAULT_OBS_NORM
from oracle import dpn_oracle as O           # synthetic query points only (diagnostic tool, not a product path)

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["fp32", "f16x3"]
dev = torch.device("cuda:0")
obs = {k: dict(v, norm_type="mean_norm", use_norm=True) for k, v in DEFAULT_OBS_NORM.items()}
torch.manual_seed(0)
base = InterfacePhysics(BN.META_CFG, BN.NET_CFG, obs, None, dict(img_size=(145, 257), dx=27000, dy=27000)).to(dev)


def batch(seed, n_inter=4096, n_margin=20480):
    g = torch.Generator().manual_seed(seed)
    mk = lambda n: [a.reshape(1, -1).float().to(dev) for a in O.synthetic_points(n, g)[:4]]
    ix, iy, it_, if_ = mk(n_inter)
    mx, my, mt, mf = mk(n_margin)
    return dict(field_data=torch.randn(1, 159, 2405, generator=g).to(dev), forecast_h=torch.full((1, 1, 1), 24.0 / 360.0, device=dev),
                inter_x=ix, inter_y=iy, inter_t=it_, inter_f=if_, inter_data=(0.5 * torch.randn(1, n_inter, 6, generator=g)).to(dev),
                margin_x=mx, margin_y=my, margin_t=mt, margin_f=mf,
                margin_input_data=(0.5 * torch.randn(1, n_margin, 6, generator=g)).to(dev),
                margin_data=(0.5 * torch.randn(1, n_margin, 6, generator=g)).to(dev))


batches = [batch(100 + i % 4) for i in range(4)]           # four "files", cycled
traj = {}
for mode in modes:
    m = copy.deepcopy(base)
    m.mode = mode
    step = TrainStep(m, pde_start_step=3)                 # three data-loss-only steps, then margin + interior PDE + margin PDE
    out = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(steps):
        p = step(batches[i % 4])
        out.append((p["train_loss"].item(), p["grad_norm"].item()))
    torch.cuda.synchronize()
    traj[mode] = out
    print("mode %-6s %d steps in %.2f s (%.1f ms/step incl. .item())" % (mode, steps, time.perf_counter() - t0, 1e3 * (time.perf_counter() - t0) / steps))
print("%4s  " % "step" + "  ".join("%-26s" % ("loss / grad-norm  [%s]" % m) for m in modes) + ("  rel. loss diff" if len(modes) == 2 else ""))
for i in range(steps):
    row = "%4d  " % (i + 1) + "  ".join("%12.5e %12.5e " % traj[m][i] for m in modes)
    if len(modes) == 2:
        a, b = traj[modes[0]][i][0], traj[modes[1]][i][0]
        row += "  %.1e" % (abs(a - b) / abs(a))
    print(row)
