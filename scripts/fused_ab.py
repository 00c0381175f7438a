"""A/B of the query-sharded step under torchrun: NCCL all-gather form vs fused peer-window form, same process."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import prgpu_loader
P = prgpu_loader.load(); synth = P.synth
N, Q, k = 10_000_000, 4096, 16
stream = torch.cuda.current_stream().cuda_stream
tree = P.KnnIndex(7, device=local, capacity=N)
tree.add_device(torch.from_numpy(synth.cloud(0, N)).to(dev), N, stream); torch.cuda.synchronize(); tree.build_index()
q = torch.from_numpy(synth.cloud(1000 + rank, Q)).to(dev)
gi = torch.empty((world, Q, k), dtype=torch.int32, device=dev); gd = torch.empty((world, Q, k), dtype=torch.float64, device=dev)
ca = P.sharded.make_comm(P, dist, local, rank, world)
cb = P.sharded.make_comm(P, dist, local, rank, world)
ok = cb.enable_peer_windows(12 * k * Q * world)
def timed(c, reps=30):
    for _ in range(5):
        tree.nearest_k_query_sharded_device(c, q, Q, k, gi, gd, stream=stream)
    torch.cuda.synchronize(); dist.barrier()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps):
        tree.nearest_k_query_sharded_device(c, q, Q, k, gi, gd, stream=stream)
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([b.elapsed_time(e) / reps], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
for rnd in range(3):
    a = timed(ca); f = timed(cb)
    ra = gi.clone(); 
    if rank == 0:
        print(f"round {rnd}: all-gather {a:.4f} ms, fused({ok}) {f:.4f} ms, fused_steps={cb.fused_steps()}", flush=True)
# local search alone
for _ in range(3):
    tree.nearest_k_device(q, Q, k, gi[0], gd[0], stream=stream)
torch.cuda.synchronize(); b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
b.record()
for _ in range(30):
    tree.nearest_k_device(q, Q, k, gi[0], gd[0], stream=stream)
e.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"local search alone {b.elapsed_time(e) / 30:.4f} ms", flush=True)
ca.close(); cb.close(); dist.barrier(); dist.destroy_process_group()
