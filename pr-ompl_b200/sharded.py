"""Host-side sharding logic for the multi-GPU paths: one process per GPU,
torch.distributed for the plumbing.

  * kNN over a big tree: tree nodes are block-sharded over the ranks (`shard_bounds`), every rank
    searches its shard for ALL queries and reports global indices (index_base = shard start); the
    per-shard top-k tables are all-gathered (Q*k*12 bytes per rank) and merged under the same
    (distance, index) order, so the result is identical to the single-GPU one.
  * edge batches / independent planning queries: split with `shard_bounds`, no data-path collective.

The search and the merge are injected callables: the product passes the CUDA ones
(KnnIndex.nearest_k_device / topk_merge_device); the CPU gloo tests pass checkers.
"""


def shard_bounds(n, world):
    """[(begin, end)] of the `world` contiguous shards of n units (sizes differ by at most 1)."""
    return [(r * n // world, (r + 1) * n // world) for r in range(world)]


def sharded_nearest_k(local_search, merge, dist, world, queries, nq, k, bufs):
    """One sharded kNN step.

    local_search(queries, nq, k, idx_out, dist_out): this rank's shard, global indices.
    merge(idx_parts, dist_parts, nparts, nq, k, idx_out, dist_out): ordered merge of `nparts` tables.
    bufs: dict with 'idx','dist' (nq x k), 'gidx','gdist' (world x nq x k), 'fidx','fdist' (nq x k).
    Returns (idx, dist) of the global top-k (the local tables themselves when world == 1).
    """
    local_search(queries, nq, k, bufs["idx"], bufs["dist"])
    if world == 1:
        return bufs["idx"], bufs["dist"]
    # (world, nq, k) viewed as the concatenation (world*nq, k): the layout every backend accepts
    dist.all_gather_into_tensor(bufs["gidx"].view(world * nq, k), bufs["idx"])
    dist.all_gather_into_tensor(bufs["gdist"].view(world * nq, k), bufs["dist"])
    merge(bufs["gidx"], bufs["gdist"], world, nq, k, bufs["fidx"], bufs["fdist"])
    return bufs["fidx"], bufs["fdist"]


def make_comm(P, dist, device, rank, world):
    """prgpu communicator for this rank: rank 0 draws the NCCL unique id (prgpu_comm_unique_id), the 128 bytes
    travel over the host application's own channel -- here torch.distributed -- and every rank calls
    prgpu_comm_init.  The data path afterwards is C ABI only (prgpu_knn_nearest_k_sharded_device ...)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", device))
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(P.Comm.unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return P.Comm(device, rank, world, unique_id=bytes(buf.cpu().numpy().tobytes()))
