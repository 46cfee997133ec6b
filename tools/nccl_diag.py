import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
if rank == 0:
    os.system("nvidia-smi topo -m")
    print("can_access_peer 0->1:", torch.cuda.can_device_access_peer(0, 1))
for mb in (4.4, 22.4, 256):
    n = int(mb * 1e6 / 4)
    x = torch.ones(n, device="cuda")
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_reduce(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0: print("all_reduce %.1f MB: %.3f ms  algbw %.1f GB/s" % (mb, ms, mb / ms))
try:
    import torch.distributed._symmetric_memory as symm
    t = symm.empty(1024, device="cuda")
    hdl = symm.rendezvous(t, dist.group.WORLD)
    if rank == 0: print("symmetric memory OK: world", hdl.world_size, "multicast", getattr(hdl, "multicast_ptr", None))
except Exception as e:
    if rank == 0: print("symmetric memory failed:", repr(e)[:300])
dist.destroy_process_group()
