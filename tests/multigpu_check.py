"""Multi-GPU parity check of the C-ABI sharded entry points (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multigpu_check.py
Every rank holds one shard of the tree (prgpu_knn_set_index_base) and calls
prgpu_knn_nearest_k_sharded_device over the library's own NCCL communicator; rank 0 compares the result
with the CPU oracle over the WHOLE tree (bit for bit), including a duplicate-heavy cloud that forces the
"some shard left queries unproven -> every rank repeats the step" branch, and the sharded checkMotion with
the single-GPU call.  Exit code 0 = all ranks agree with the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import oracle_lib as O
    import prgpu_loader
    P = prgpu_loader.load()
    synth = P.synth
    comm = P.sharded.make_comm(P, dist, local, rank, world)
    stream = torch.cuda.current_stream().cuda_stream
    ok = True
    for name, n, nq, dup in (("uniform", 300_000, 512, 0.0), ("duplicates", 120_000, 384, 0.6), ("tiny", 5_000, 130, 0.1)):
        pts = synth.cloud(20, n, dup_fraction=dup)
        if name == "duplicates":                    # every query has far more than k exact ties
            pts[n // 2:] = pts[:n - n // 2]
        q = synth.cloud(21, nq)
        if name == "duplicates":
            q[:64] = pts[:64]
        b0, b1 = P.sharded.shard_bounds(n, world)[rank]
        tree = P.KnnIndex(7, device=local, capacity=b1 - b0)
        tree.set_index_base(b0)
        tree.add(pts[b0:b1])
        q_dev = torch.from_numpy(q).to(dev)
        for k in (1, 16):
            fi = torch.empty((nq, k), dtype=torch.int32, device=dev)
            fd = torch.empty((nq, k), dtype=torch.float64, device=dev)
            tree.nearest_k_sharded_device(comm, q_dev, nq, k, fi, fd, stream=stream)
            torch.cuda.synchronize()
            hi, hd = tree.nearest_k_sharded(comm, q, k)          # host-buffer form
            oi, od = O.knn(pts, q, k, nthreads=8)
            same = (np.array_equal(fi.cpu().numpy().view(np.uint32), oi) and np.array_equal(fd.cpu().numpy(), od)
                    and np.array_equal(hi, oi) and np.array_equal(hd, od))
            if not same:
                print(f"[rank {rank}] MISMATCH {name} k={k}", flush=True)
            ok = ok and same
        if rank == 0:
            print(f"{name}: n={n} world={world} ok={ok} redo_steps={comm.redo_steps()} "
                  f"stats={tree.search_stats()}", flush=True)
        tree.close()
    # the other partitioning: whole tree on every rank, queries split, every rank receives every answer
    n, nq_local = 400_000, 300
    pts = synth.cloud(30, n, dup_fraction=0.05)
    tree = P.KnnIndex(7, device=local, capacity=n)
    tree.add(pts)
    q_all = synth.cloud(31, nq_local * world)
    for windows in (False, True):                 # all-gather form, then the fused search + exchange over peer memory
        if windows:
            fused_ok = comm.enable_peer_windows(12 * 32 * nq_local * world)
            if rank == 0:
                print(f"peer windows: {'mapped' if fused_ok else 'NOT available (all-gather path stays)'}", flush=True)
        for mode in (0, 3):                       # auto, and the k-d index forced
            tree.set_search_mode(mode)
            for k in (1, 16):
                for rep in range(3):              # (both window halves, flag rows reused)
                    hi, hd = tree.nearest_k_query_sharded(comm, q_all[rank * nq_local:(rank + 1) * nq_local], k)
                oi, od = O.knn(pts, q_all, k, nthreads=8)
                same = np.array_equal(hi.reshape(-1, k), oi) and np.array_equal(hd.reshape(-1, k), od)
                if not same:
                    print(f"[rank {rank}] MISMATCH query-sharded windows={windows} mode={mode} k={k}", flush=True)
                ok = ok and same
    if rank == 0:
        print(f"query-sharded: n={n} world={world} ok={ok} index={tree.index_stats()['indexed_nodes']} "
              f"fused_steps={comm.fused_steps()}", flush=True)
    tree.close()
    # edges: replicated batch, sliced validation, all-gather of the verdicts
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3)
    gw = P.World(dh, link, xyzr, sph, box, device=local)
    for n_e in (100_003, 4_001, 7):
        a, b = synth.edges(2, n_e)
        a_dev, b_dev = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        v1 = torch.empty(n_e, dtype=torch.uint8, device=dev)
        v2 = torch.empty(n_e, dtype=torch.uint8, device=dev)
        gw.check_motion_batch_device(a_dev, b_dev, n_e, 0.01, v1, stream=stream)
        gw.check_motion_batch_sharded_device(comm, a_dev, b_dev, n_e, 0.01, v2, stream=stream)
        torch.cuda.synchronize()
        same = bool(torch.equal(v1, v2))
        if not same:
            print(f"[rank {rank}] MISMATCH check_motion n={n_e}", flush=True)
        ok = ok and same
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("multigpu_check:", "OK" if int(t.item()) == 1 else "FAILED", flush=True)
    sys.exit(0 if int(t.item()) == 1 else 1)


if __name__ == "__main__":
    main()
