"""Times small all_gathers and prints NCCL's transport choice (development aid)."""
import os, sys, time, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
for nbytes in (1 << 12, 1 << 18, 1 << 20, 1 << 24):
    x = torch.zeros(nbytes // 4, dtype=torch.int32, device="cuda")
    out = torch.zeros(world * nbytes // 4, dtype=torch.int32, device="cuda")
    for _ in range(5):
        dist.all_gather_into_tensor(out, x)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        dist.all_gather_into_tensor(out, x)
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"all_gather {nbytes} B/rank: {e0.elapsed_time(e1)/20*1e3:.1f} us", flush=True)
dist.destroy_process_group()
