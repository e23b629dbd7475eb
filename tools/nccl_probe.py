"""2-rank NCCL all-gather bandwidth probe (is NVLink P2P in use on this box?)."""
import os, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (1, 16, 43, 128):
    x = torch.ones(mb * 131072, dtype=torch.float64, device="cuda")
    out = torch.empty(world * x.numel(), dtype=torch.float64, device="cuda")
    for _ in range(3): dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if rank == 0: print(f"all_gather {mb} MiB/rank: {ms:.3f} ms -> {mb * 1.048576 * (world - 1) / ms:.1f} GB/s per rank received", flush=True)
if rank == 0:
    print("can_device_access_peer(0,1):", torch.cuda.can_device_access_peer(0, 1))
dist.destroy_process_group()
