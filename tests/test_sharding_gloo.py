"""N>1 host logic on CPU: two gloo ranks run the sharded-kNN step (shard bounds, global index
bases, all-gather of per-shard top-k, ordered merge) with CPU checkers injected for the search and
the merge, and must reproduce the single-process result; query/edge splitting is checked too."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_np(gi, gd, nparts, nq, k, fi, fd):
    gi_, gd_ = gi.numpy().view(np.uint32), gd.numpy()
    for q in range(nq):
        cand = [(gd_[p, q, j], gi_[p, q, j]) for p in range(nparts) for j in range(k)]
        cand.sort()
        for j in range(k):
            fd[q, j] = cand[j][0]
            fi[q, j] = int(np.uint32(cand[j][1]).view(np.int32))


def _worker(rank, world, port, n, nq, k, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as O
    import prgpu_loader
    P = prgpu_loader.load()
    pts = P.synth.cloud(0, n, dup_fraction=0.1)
    q = P.synth.cloud(1, nq)
    b0, b1 = P.sharded.shard_bounds(n, world)[rank]

    def local_search(queries, nq_, k_, idx_out, dist_out):
        i, d = O.knn(pts[b0:b1], queries, k_)
        i = np.where(i == 0xFFFFFFFF, i, i + np.uint32(b0)).astype(np.uint32)   # index_base
        idx_out.copy_(torch.from_numpy(i.view(np.int32)))
        dist_out.copy_(torch.from_numpy(d))

    bufs = dict(idx=torch.empty((nq, k), dtype=torch.int32), dist=torch.empty((nq, k), dtype=torch.float64),
                gidx=torch.empty((world, nq, k), dtype=torch.int32),
                gdist=torch.empty((world, nq, k), dtype=torch.float64),
                fidx=torch.empty((nq, k), dtype=torch.int32), fdist=torch.empty((nq, k), dtype=torch.float64))
    fi, fd = P.sharded.sharded_nearest_k(local_search, _merge_np, dist, world, q, nq, k, bufs)
    if rank == 0:
        ri, rd = O.knn(pts, q, k)
        ok = np.array_equal(fi.numpy().view(np.uint32), ri) and np.array_equal(fd.numpy(), rd)
        with open(out, "w") as f:
            f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_knn_two_ranks(tmp_path):
    import oracle_lib
    oracle_lib.build()
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), 3001, 37, 5, out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_shard_bounds(P):
    for n in (0, 1, 7, 4096, 10 ** 7 + 3):
        for w in (1, 2, 3, 8):
            b = P.sharded.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def _comm_worker(rank, world, port, out):
    """prgpu_comm_init_custom with a gloo all-gather injected: the C-ABI communicator's plumbing (rank/size,
    block placement rank r at r * bytes) on CPU, host-only communicator (device -1)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C
    import prgpu_loader
    P = prgpu_loader.load()

    def allgather(ctx, send, recv, nbytes, stream):
        s = torch.frombuffer((C.c_uint8 * nbytes).from_address(send), dtype=torch.uint8)
        r = torch.frombuffer((C.c_uint8 * (nbytes * world)).from_address(recv), dtype=torch.uint8)
        dist.all_gather_into_tensor(r, s)
        return 0

    comm = P.Comm(-1, rank, world, allgather=allgather)
    ok = P.lib().prgpu_comm_rank(comm.h) == rank and P.lib().prgpu_comm_size(comm.h) == world
    n = 1000
    send = np.full(n, rank + 1, dtype=np.uint8)
    send[:4] = np.frombuffer(np.uint32(0xABCD0000 + rank).tobytes(), dtype=np.uint8)
    recv = np.zeros(n * world, dtype=np.uint8)
    comm.allgather(send, recv, n)
    for r in range(world):
        blk = recv[r * n:(r + 1) * n]
        ok = ok and np.frombuffer(blk[:4].tobytes(), dtype=np.uint32)[0] == 0xABCD0000 + r and (blk[4:] == r + 1).all()
    # the compute entry points refuse a host-only communicator (and there is no GPU here anyway)
    comm.close()
    with open(out + str(rank), "w") as f:
        f.write("ok" if ok else "mismatch")
    dist.barrier()
    dist.destroy_process_group()


def test_c_abi_comm_custom_allgather_two_ranks(tmp_path):
    out = str(tmp_path / "comm")
    mp.spawn(_comm_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out + "0").read() == "ok" and open(out + "1").read() == "ok"
