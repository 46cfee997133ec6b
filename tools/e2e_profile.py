"""torch.profiler view of the e2e step (InterfacePhysics.place_one_batch with host tensors): where the GPU time goes."""
import sys, json, torch
sys.path.insert(0, ".")
import bench as Bn
from deepphysinet_b200 import InterfacePhysics
from deepphysinet_b200.config import DEFAULT_LOSS_FACTOR, DEFAULT_OBS_NORM
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
obs = {k: dict(v, norm_type="mean_norm", use_norm=True) for k, v in DEFAULT_OBS_NORM.items()}
torch.manual_seed(0)
model = InterfacePhysics(Bn.META_CFG, Bn.NET_CFG, obs, None, dict(img_size=(145, 257), dx=27000, dy=27000)).to(dev)
B, Np = 8, 65536
g = torch.Generator().manual_seed(100)
hx = (torch.rand(B, Np, generator=g) * 256 * 27000.0).pin_memory()
hy = (torch.rand(B, Np, generator=g) * 144 * 27000.0).pin_memory()
hf = (1e-4 * torch.rand(B, Np, generator=g)).pin_memory()
ht = (torch.randint(0, 25, (B, Np), generator=g).float() * 3600.0).pin_memory()
hcd = (0.5 * torch.randn(B, Np, 6, generator=g)).pin_memory()
hfield = torch.randn(B, 159, 2405, generator=g).pin_memory()
hfh = torch.full((B, 1, 1), 24.0 / 360.0).pin_memory()
crit = torch.nn.MSELoss()
def step():
    model.physics_net.zero_grad(set_to_none=True)
    loss = model.place_one_batch(hx, hy, ht, hf, hfield, hcd, hfh, crit, DEFAULT_LOSS_FACTOR, 0, 0, dev)
    loss.backward()
    return loss.item()
for _ in range(5): step()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, evs[-1].time_range.end
busy = sum(e.time_range.end - e.time_range.start for e in evs)
print("GPU span %.2f ms for 3 steps, kernel-busy %.2f ms (%.0f%%)" % ((t1 - t0) / 1e3, busy / 1e3, 100 * busy / (t1 - t0)))
