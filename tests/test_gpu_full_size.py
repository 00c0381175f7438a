"""Parity at BASELINE.json's full sizes, through size-independent properties and oracle spot checks."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_knn_sweep_full_size(P, O, synth):
    """N=1e7, Q=4096, k=16 and k=1: the tensor-core filter path (mode 2) and the k-d index (mode 0 after
    prgpu_knn_build_index) equal the exact kernel (mode 1) on every query; a spread subset equals the CPU
    oracle; rows are sorted; reported distances are the reference's distance of the reported nodes."""
    import torch
    n, nq = 10_000_000, 4096
    pts = synth.cloud(0, n)
    q = synth.cloud(1, nq)
    t = P.KnnIndex(7, capacity=n)
    t.add(pts)
    sub = np.arange(0, nq, 128)
    for k in (16, 1):
        t.set_search_mode(2)
        fi, fd = t.nearest_k(q, k)
        st = t.search_stats()
        assert st["filter_calls"] > 0
        t.set_search_mode(True)
        ei, ed = t.nearest_k(q, k)
        assert np.array_equal(fi, ei) and np.array_equal(fd, ed)
        t.set_search_mode(0)
        t.build_index()
        ii, id_ = t.nearest_k(q, k)
        ist = t.index_stats()
        assert ist["indexed_nodes"] == n and ist["index_calls"] >= 1
        assert np.array_equal(ii, ei) and np.array_equal(id_, ed)
        oi, od = O.knn(pts, q[sub], k, nthreads=16)
        assert np.array_equal(fi[sub], oi) and np.array_equal(fd[sub], od)
        assert np.all(np.diff(fd, axis=1) >= 0)
        rows = np.arange(0, nq, 37)
        d = np.sqrt(((pts[fi[rows, 0]] - q[rows]) ** 2).sum(1))
        assert np.allclose(d, fd[rows, 0], rtol=1e-13)
    # suffix masks (NNF's "strictly deeper layers") on the large-tree path: tensor-core sample pass + scan
    lo = np.sort(np.random.default_rng(5).integers(0, n - 50_000, nq)).astype(np.uint32)
    lo[-8:] = n - np.arange(1, 9, dtype=np.uint32) * 3      # almost empty suffixes: fewer than k eligible nodes
    t.set_search_mode(False)
    fi, fd = t.nearest_k(q, 5, lo=lo)
    t.set_search_mode(True)
    ei, ed = t.nearest_k(q, 5, lo=lo)
    assert np.array_equal(fi, ei) and np.array_equal(fd, ed)
    valid = fi != 0xFFFFFFFF
    assert np.all(fi[valid] >= np.broadcast_to(lo[:, None], fi.shape)[valid])
    t.close()
    torch.cuda.empty_cache()


def test_check_motion_full_size(P, O, synth, gpu_world, oracle_world):
    """1e6 edges at 0.01 rad: flattened two-phase path; a 16k-edge subset equals the oracle outside the
    contact band; the verdict is independent of batch composition (idempotence on a re-ordered batch)."""
    n = 1_000_000
    a, b = synth.edges(2, n)
    v = gpu_world.check_motion_batch(a, b, 0.01)
    assert 0.45 < v.mean() < 0.65
    sub = np.arange(0, n, 61)
    ov, oc, _ = oracle_world.check_motion_batch(a[sub], b[sub], 0.01, nthreads=16)
    diff = v[sub] != ov
    assert not np.any(diff & (np.abs(oc) >= 1e-6)), np.argwhere(diff)[:5]
    perm = np.random.default_rng(0).permutation(n)[:200_000]
    v2 = gpu_world.check_motion_batch(a[perm], b[perm], 0.01)
    assert np.array_equal(v2, v[perm])
    # small-batch kernel (warp per edge) agrees with the flattened kernels
    v3 = gpu_world.check_motion_batch(a[:3000], b[:3000], 0.01)
    assert np.array_equal(v3, v[:3000])


def test_rrtc_4096_queries(P, O, synth, gpu_world, oracle_world):
    """BASELINE configs[4], all 4096 independent queries: EVERY query reproduces the oracle port's status,
    iteration count, both trees (states and parents) and path bit for bit; every 64th query is also run by the
    reference's own RRTConnect.cpp (compiled unmodified, oracle/_ref) when that library is present."""
    from concurrent.futures import ThreadPoolExecutor
    n = 4096
    q = synth.cloud(4, 8 * n)
    v = gpu_world.is_valid_batch(q)
    qv = q[v == 1]
    starts, goals = np.ascontiguousarray(qv[0:2 * n:2]), np.ascontiguousarray(qv[1:2 * n:2])
    rb = P.RrtcBatch(gpu_world, n, -np.pi, np.pi, extension_type="ext-con", seed=4, max_iters=1000, max_nodes=2048)
    rb.run(starts, goals)
    s = rb.summary()
    assert (s["status"] == 6).mean() > 0.9
    have_ref = O.ref() is not None
    import threading
    fetch_lock = threading.Lock()       # a C-ABI handle is single-threaded

    def check(qi):
        for impl in (("port", "ref") if have_ref and qi % 64 == 0 else ("port",)):
            ref = oracle_world.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, ext=(0, 1), seed=4, query=qi,
                                          max_iters=1000, max_nodes=4096, impl=impl)
            if s["status"][qi] != ref["status"] or s["iterations"][qi] != ref["iterations"]:
                return (qi, impl, "status/iterations")
            with fetch_lock:
                got = rb.fetch(qi, int(s["n_start"][qi]), int(s["n_goal"][qi]), max_path=1024)
            for key in ("start_parent", "goal_parent", "start_states", "goal_states", "path"):
                if not np.array_equal(got[key], ref[key]):
                    return (qi, impl, key)
        return None
    # the oracle calls release the GIL (ctypes); rb.fetch is a small synchronous read-back
    with ThreadPoolExecutor(max_workers=8) as pool:
        bad = [r for r in pool.map(check, range(n)) if r is not None]
    assert not bad, bad[:5]


def test_sharded_c_abi_on_all_visible_gpus():
    """prgpu_comm_* + prgpu_knn_nearest_k_sharded_device + prgpu_check_motion_batch_sharded_device under
    torchrun on every GPU of the box (tests/multigpu_check.py); needs at least 2."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under gpurun --gpus 2)")
    n = 2 if n < 4 else (4 if n < 8 else 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29611",
                        os.path.join(root, "tests", "multigpu_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
