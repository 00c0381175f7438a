"""K1 parity: CUDA nearest / nearestK through the C ABI vs the CPU oracle (bit-exact indices
under the (sqrt-distance, insertion index) order, bit-exact distances)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check(P, O, synth, n, nq, k, dim=7, dup=0.0, lo=None, seed=0):
    pts = synth.cloud(seed, n, dim, dup_fraction=dup)
    q = synth.cloud(seed + 1, nq, dim)
    if dup > 0 and n > 0:
        q[: nq // 2] = pts[np.random.default_rng(5).integers(0, n, nq // 2)]  # zero-distance hits
    t = P.KnnIndex(dim)
    t.add(pts)
    assert t.size() == n
    gi, gd = t.nearest_k(q, k, lo=lo)
    oi, od = O.knn(pts, q, k, lo=lo, nthreads=8)
    assert np.array_equal(gi, oi), (n, nq, k, np.argwhere(gi != oi)[:5])
    assert np.array_equal(gd, od)
    t.close()


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 255, 256, 257, 1000, 10000])
@pytest.mark.parametrize("k", [1, 16])
def test_small_trees(P, O, synth, n, k):
    for nq in (1, 5, 130):
        _check(P, O, synth, n, nq, k)


@pytest.mark.parametrize("k", [1, 5, 16, 33, 64])
def test_k_sweep(P, O, synth, k):
    _check(P, O, synth, 20000, 300, k)


@pytest.mark.parametrize("dim", [1, 2, 3, 6, 8])
def test_dims(P, O, synth, dim):
    _check(P, O, synth, 5000, 200, 4, dim=dim)


def test_duplicates_tie_break(P, O, synth):
    # exact duplicates: ties must resolve to the lower insertion index
    _check(P, O, synth, 30000, 512, 1, dup=0.3)
    _check(P, O, synth, 30000, 512, 16, dup=0.3)
    pts = np.zeros((700, 7))
    t = P.KnnIndex(7)
    t.add(pts)
    idx, dist = t.nearest_k(np.zeros((3, 7)), 5)
    assert np.array_equal(idx, np.tile(np.arange(5, dtype=np.uint32), (3, 1)))
    assert np.all(dist == 0.0)


def test_suffix_mask(P, O, synth):
    # findKnnNodes: candidates = nodes of strictly deeper layers = index >= lo[q]
    n, nq = 6000, 400
    lo = np.sort(np.random.default_rng(9).integers(0, n + 50, nq)).astype(np.uint32)
    _check(P, O, synth, n, nq, 5, lo=lo)
    _check(P, O, synth, n, nq, 1, lo=lo)


def test_batched_sweep_shape(P, O, synth):
    # BASELINE config "batched kNN sweep" at its smallest size: N=1e5 here, Q=4096, k=1/16
    n, nq = 100000, 4096
    pts = synth.cloud(0, n)
    q = synth.cloud(1, nq)
    t = P.KnnIndex(7)
    t.add(pts[: n // 2])
    t.add(pts[n // 2:])  # incremental add keeps insertion order
    sub = np.arange(0, nq, 16)
    for k in (1, 16):
        gi, gd = t.nearest_k(q, k)
        oi, od = O.knn(pts, q[sub], k, nthreads=8)
        assert np.array_equal(gi[sub], oi)
        assert np.array_equal(gd[sub], od)
        # size-independent properties on the full result
        assert np.all(np.diff(gd, axis=1) >= 0)
        d0 = np.sqrt(((pts[gi[:, 0]] - q) ** 2).sum(1))
        assert np.allclose(d0, gd[:, 0], rtol=1e-12)
    assert np.array_equal(t.points(), pts)


def test_empty_and_k_gt_n(P, O, synth):
    t = P.KnnIndex(7)
    idx, dist = t.nearest_k(synth.cloud(1, 4), 3)
    assert np.all(idx == P.INVALID_INDEX) and np.all(np.isinf(dist))
    t.add(synth.cloud(0, 2))
    _check(P, O, synth, 2, 9, 16)
    t.clear()
    assert t.size() == 0


def test_topk_merge_matches_single_pass(P, O, synth):
    import torch
    n, nq, k, shards = 50000, 512, 16, 4
    pts = synth.cloud(0, n)
    q = synth.cloud(1, nq)
    parts_i, parts_d = [], []
    bounds = np.linspace(0, n, shards + 1).astype(int)
    for s in range(shards):
        t = P.KnnIndex(7)
        t.add(pts[bounds[s]:bounds[s + 1]])
        t.set_index_base(int(bounds[s]))
        i, d = t.nearest_k(q, k)
        parts_i.append(i)
        parts_d.append(d)
    pi = torch.from_numpy(np.stack(parts_i).view(np.int32)).cuda()
    pd = torch.from_numpy(np.stack(parts_d)).cuda()
    oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    od = torch.empty((nq, k), dtype=torch.float64, device="cuda")
    P.topk_merge_device(0, pi, pd, shards, nq, k, oi, od, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ri, rd = O.knn(pts, q, k, nthreads=8)
    assert np.array_equal(oi.cpu().numpy().view(np.uint32), ri)
    assert np.array_equal(od.cpu().numpy(), rd)


# ---- the fp32-filter fast path must be indistinguishable from the exact kernel -----------------

def _both_modes(P, O, pts, q, k, lo=None, expect_filter=True):
    t = P.KnnIndex(pts.shape[1])
    t.add(pts)
    gi, gd = t.nearest_k(q, k, lo=lo)
    st = t.search_stats()
    assert (st["filter_calls"] > 0) == expect_filter
    t.set_search_mode(True)
    ei, ed = t.nearest_k(q, k, lo=lo)
    assert np.array_equal(gi, ei) and np.array_equal(gd, ed)
    oi, od = O.knn(pts, q, k, lo=lo, nthreads=8)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    t.close()
    return st


@pytest.mark.parametrize("k", [1, 5, 11, 16])
def test_filter_path_uniform(P, O, synth, k):
    st = _both_modes(P, O, synth.cloud(0, 60000), synth.cloud(1, 700), k)
    assert st["fallback_queries"] == 0          # well separated data: every query is proven


def test_filter_path_not_taken_outside_its_shape(P, O, synth):
    _both_modes(P, O, synth.cloud(0, 60000), synth.cloud(1, 700), 17, expect_filter=False)     # k > 16
    _both_modes(P, O, synth.cloud(0, 60000), synth.cloud(1, 100), 4, expect_filter=False)      # < 128 queries
    _both_modes(P, O, synth.cloud(0, 5000), synth.cloud(1, 300), 4, expect_filter=False)       # small tree
    _both_modes(P, O, synth.cloud(0, 20000, dim=8), synth.cloud(1, 300, dim=8), 4, expect_filter=False)


def test_filter_path_adversarial_ties(P, O, synth):
    rng = np.random.default_rng(3)
    # (a) every point identical: nothing can be proven, everything falls back
    pts = np.tile(synth.cloud(5, 1), (20000, 1))
    st = _both_modes(P, O, pts, synth.cloud(1, 256), 3)
    assert st["fallback_queries"] == 256
    # (b) heavy exact duplication (each point 12x) and queries sitting on points
    base = synth.cloud(6, 2500)
    pts = np.repeat(base, 12, axis=0)[rng.permutation(30000)]
    q = np.vstack([base[:200], synth.cloud(7, 200)])
    _both_modes(P, O, pts, q, 16)
    # (c) clusters of near-duplicates separated by 1e-9 .. 1e-6 (below the fp32 resolution)
    centres = synth.cloud(8, 1000)
    pts = (centres[:, None, :] + rng.uniform(-1, 1, (1000, 20, 7)) * 10.0 ** rng.uniform(-9, -6, (1000, 20, 1))
           ).reshape(-1, 7)
    _both_modes(P, O, pts, np.vstack([centres[:300], synth.cloud(9, 100)]), 8)
    # (d) far from the origin: the error bound grows with |x|^2, the proof gets harder, results stay exact
    pts = synth.cloud(10, 40000) * 0.01 + 250.0
    _both_modes(P, O, pts, synth.cloud(11, 300) * 0.01 + 250.0, 5)
    # (e) blobs much tighter than the TF32 error of the tensor-core scan (2^-9 |q||x| ~ 0.05 here): the
    #     level-1 cut keeps whole blobs, the fp32 re-cut or the exact fallback has to sort them out
    centres = synth.cloud(12, 64)
    pts = (centres[:, None, :] + rng.normal(0.0, 0.02, (64, 600, 7))).reshape(-1, 7)
    q = np.vstack([centres + rng.normal(0.0, 0.02, centres.shape), synth.cloud(13, 192)])
    _both_modes(P, O, pts, q, 16)
    # (f) very large and very small coordinate scales (TF32 keeps the fp32 exponent range)
    for scale in (1e6, 1e-6):
        _both_modes(P, O, synth.cloud(14, 20000) * scale, synth.cloud(15, 256) * scale, 7)


def test_filter_path_suffix_mask_and_dims(P, O, synth):
    n, nq = 50000, 600
    lo = np.sort(np.random.default_rng(9).integers(0, n + 50, nq)).astype(np.uint32)
    _both_modes(P, O, synth.cloud(0, n), synth.cloud(1, nq), 5, lo=lo)
    for dim in (2, 3, 6):
        _both_modes(P, O, synth.cloud(0, 30000, dim=dim), synth.cloud(1, 300, dim=dim), 6)


def test_filter_path_incremental_adds_and_growth(P, O, synth):
    # the fp32 / TF32 mirrors and the pair-wise |x|^2 array must survive odd offsets and reallocation
    pts = synth.cloud(20, 30011)
    q = synth.cloud(21, 300)
    t = P.KnnIndex(7)                      # no capacity hint: grows by doubling
    cuts = [0, 4097, 4098, 4099, 4100, 9001, 20000, 30011]
    for a, b in zip(cuts[:-1], cuts[1:]):
        t.add(pts[a:b])
        if b >= 9001:                      # query between adds, as a planner does
            gi, gd = t.nearest_k(q, 5)
            oi, od = O.knn(pts[:b], q, 5, nthreads=8)
            assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    st = t.search_stats()
    assert st["filter_calls"] >= 3 and st["fallback_queries"] == 0
    t.clear()
    t.add(pts[:10001][::-1].copy())        # stale mirror contents beyond n must not leak
    gi, gd = t.nearest_k(q, 3)
    oi, od = O.knn(pts[:10001][::-1].copy(), q, 3, nthreads=8)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    t.close()


def test_add_device_on_foreign_stream_after_create_and_regrow(P, O, synth):
    """prgpu_knn_add_device runs on the CALLER's stream: the handle's own bookkeeping (zeroing max|x|^2 at
    create, copying the buffers on a regrow) must be ordered against it.  A lost max|x|^2 would collapse
    the filter path's error bounds, lost points would change the answers."""
    import torch
    pts = synth.cloud(5, 60000) * 1.0
    q = synth.cloud(6, 512)
    stream = torch.cuda.Stream()
    t = torch.from_numpy(pts).cuda()
    torch.cuda.synchronize()
    for rep in range(3):
        tree = P.KnnIndex(7, capacity=8192)                      # capacity hint: buffers + max|x|^2 exist
        with torch.cuda.stream(stream):
            tree.add_device(t[:8000], 8000, stream.cuda_stream)           # right after create
            tree.add_device(t[8000:20000], 12000, stream.cuda_stream)     # regrow with an append in flight
            tree.add_device(t[20000:], 40000, stream.cuda_stream)         # and again
        stream.synchronize()
        assert tree.size() == 60000
        assert np.array_equal(tree.points(0, 60000), pts)
        gi, gd = tree.nearest_k(q, 16)                           # >= 8192 nodes, >= 128 queries: filter path
        oi, od = O.knn(pts, q, 16, nthreads=8)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
        st = tree.search_stats()
        assert st["filter_calls"] >= 1 and st["fallback_queries"] == 0
        tree.close()


@pytest.mark.parametrize("n,dim", [(1 << 17, 7), (300_017, 7), (200_000, 3)])
def test_kd_index_equals_brute_force_and_oracle(P, O, n, dim):
    """The sub-linear path (k-d index built on the device at the second batched search of an unchanged
    node set) returns what the brute-force kernels and the CPU oracle return, bit for bit: duplicates, ties,
    queries that ARE tree nodes, queries far outside the nodes' bounding box, k up to 32."""
    rng = np.random.default_rng(n)
    pts = rng.uniform(-np.pi, np.pi, size=(n, dim))
    pts[n // 3: n // 3 + 5000] = pts[:5000]                       # exact duplicates (index tie-break)
    pts[-3000:] = pts[-1] + 1e-9 * rng.normal(size=(3000, dim))   # a tight cluster
    q = rng.uniform(-np.pi, np.pi, size=(100, dim))            # < 128 queries: 1.5e8 pairs make the index pay
    q[:40] = pts[rng.integers(0, n, 40)]
    q[40:48] = 50.0 * rng.normal(size=(8, dim))                   # far outside the bounding box
    q[48:56] = pts[-1]
    tree = P.KnnIndex(dim)
    tree.add(pts)
    for k in (1, 16, 32):
        tree.set_search_mode(2)                                   # brute force, no index
        bi, bd = tree.nearest_k(q, k)
        tree.set_search_mode(3)                                   # the index whenever it can answer
        gi, gd = tree.nearest_k(q, k)
        gi2, gd2 = tree.nearest_k(q, k)
        st = tree.index_stats()
        assert st["indexed_nodes"] == n and st["index_calls"] >= 1, st
        oi, od = O.knn(pts, q, k, nthreads=8)
        assert np.array_equal(bi, oi) and np.array_equal(bd, od)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
        assert np.array_equal(gi2, oi) and np.array_equal(gd2, od)
    # an append makes the index stale: brute force until the second search rebuilds it
    more = rng.uniform(-np.pi, np.pi, size=(1000, dim))
    tree.add(more)
    allp = np.vstack([pts, more])
    assert tree.index_stats()["indexed_nodes"] == 0
    tree.set_search_mode(0)                                       # auto: too few pairs here, brute force answers
    gi, gd = tree.nearest_k(q, 5)
    assert tree.index_stats()["indexed_nodes"] == 0
    tree.set_search_mode(3)
    for _ in range(3):
        gi, gd = tree.nearest_k(q, 5)
        oi, od = O.knn(allp, q, 5, nthreads=8)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    st = tree.index_stats()
    assert st["indexed_nodes"] == n + 1000 and st["builds"] == 2
    assert 0 < st["leaves_visited"]
    tree.close()


def test_kd_index_degenerate_clouds(P, O):
    """All nodes identical / nodes on a line / heavy duplication: the splits degenerate, the answers do not."""
    rng = np.random.default_rng(5)
    n = 1 << 17
    q = rng.uniform(-1, 1, size=(64, 7))
    clouds = [np.tile(rng.uniform(-1, 1, size=(1, 7)), (n, 1)),
              np.outer(np.linspace(-1, 1, n), np.ones(7)),
              np.repeat(rng.uniform(-1, 1, size=(n // 64, 7)), 64, axis=0)]
    for pts in clouds:
        tree = P.KnnIndex(7)
        tree.add(pts)
        tree.set_search_mode(3)
        gi, gd = tree.nearest_k(q, 16)
        assert tree.index_stats()["indexed_nodes"] == n
        oi, od = O.knn(pts, q, 16, nthreads=8)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od)
        tree.close()


@pytest.mark.parametrize("nq", [1, 7, 300, 700, 1500, 3000, 5000])
def test_kd_team_search_every_batch_size(P, O, nq):
    """The team kernel gives a block's 8 warps 1, 2, 4 or 8 queries depending on the batch size, and idle warps take
    subtrees over from busy ones; whatever the split, the answers are the brute-force answers (ties by index,
    duplicates, clustered queries whose searches are unusually long)."""
    rng = np.random.default_rng(100 + nq)
    n = 1 << 18
    pts = rng.uniform(-np.pi, np.pi, size=(n, 7))
    pts[1000:3000] = pts[:2000]                                   # duplicates
    pts[-20000:] = pts[-1] + 0.05 * rng.normal(size=(20000, 7))   # a dense blob: long searches for queries inside it
    q = rng.uniform(-np.pi, np.pi, size=(nq, 7))
    q[: max(1, nq // 5)] = pts[-1] + 0.05 * rng.normal(size=(max(1, nq // 5), 7))
    tree = P.KnnIndex(7)
    tree.add(pts)
    tree.set_search_mode(3)
    for k in (1, 16, 32):
        gi, gd = tree.nearest_k(q, k)
        oi, od = O.knn(pts, q, k, nthreads=8)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od), (nq, k)
    assert tree.index_stats()["indexed_nodes"] == n
    tree.close()


def test_kd_index_is_built_when_it_has_paid_for_itself(P, O, synth):
    """Auto mode: the index is built once the brute-force searches on an unchanged node set have cost about as
    much extra as the build (not at once: a structure that is refilled between searches must never pay for it);
    an append resets the account; answers are the same before and after."""
    n, nq, k = 4_000_000, 4096, 16
    pts = synth.cloud(0, n)
    q = synth.cloud(1, nq)
    t = P.KnnIndex(7, capacity=n + 16)
    t.add(pts)
    first_i, first_d = t.nearest_k(q, k)
    assert t.index_stats()["indexed_nodes"] == 0                # one search does not justify a build
    built_at = None
    for i in range(2, 60):
        gi, gd = t.nearest_k(q, k)
        assert np.array_equal(gi, first_i) and np.array_equal(gd, first_d)
        if t.index_stats()["indexed_nodes"] and built_at is None:
            built_at = i
    assert built_at is not None and 4 <= built_at <= 40, built_at
    t.add(synth.cloud(3, 1))                                    # stale: the account starts again
    t.nearest_k(q, k)
    t.nearest_k(q, k)
    assert t.index_stats()["indexed_nodes"] == 0
    t.build_index()                                             # the explicit call builds at once
    assert t.index_stats()["indexed_nodes"] == n + 1
    gi, gd = t.nearest_k(q[:300], k)
    oi, od = O.knn(np.vstack([pts, synth.cloud(3, 1)]), q[:300], k, nthreads=8)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    t.close()


@pytest.mark.parametrize("scale,offset", [(1.0, 0.0), (1e-6, 250.0), (1e-9, -40.0), (1e6, 0.0), (1e-6, 0.0),
                                          (1e-42, 0.0), (1e36, 0.0), (1e39, 0.0)])
def test_kd_fp32_leaf_filter_changes_nothing(P, O, monkeypatch, scale, offset):
    """The index searches read a leaf's fp32 copy first and evaluate only what that cannot reject from the fp64
    block (csrc/kdtree.cu, kd_T32).  Families chosen against the fp32 step: spacings far below the fp32 resolution
    at the data's offset (every node of a leaf collapses onto a few fp32 values), fp32-subnormal and beyond-fp32
    coordinates, exact duplicates.  Results = the oracle's = the fp64-only form's (PRGPU_KD_F32=0), bit for bit."""
    rng = np.random.default_rng(11)
    n = 1 << 17
    pts = rng.uniform(-np.pi, np.pi, size=(n, 7)) * scale + offset
    pts[5000:6000] = pts[:1000]                                   # duplicates
    q = rng.uniform(-np.pi, np.pi, size=(64, 7)) * scale + offset
    q[:16] = pts[rng.integers(0, n, 16)]
    oracle = {k: O.knn(pts, q, k, nthreads=8) for k in (1, 16)}
    for ppl in ("2", "3"):                                         # leaf capacity 64 / 96 (kd_build picks by fill)
        monkeypatch.setenv("PRGPU_KD_PPL", ppl)
        tree = P.KnnIndex(7)
        tree.add(pts)
        tree.set_search_mode(3)
        for k in (1, 16):
            oi, od = oracle[k]
            for f32 in ("1", "0"):                                 # filtered form (default), fp64-only form
                monkeypatch.setenv("PRGPU_KD_F32", f32)
                gi, gd = tree.nearest_k(q, k)
                assert np.array_equal(gi, oi) and np.array_equal(gd, od), (k, f32, ppl)
        assert tree.index_stats()["indexed_nodes"] == n
        tree.close()


@pytest.mark.parametrize("window", [True, False])
def test_small_calls_through_the_host_window(P, O, monkeypatch, window):
    """A planner drives the structure one call at a time: add(1 node), nearest(1 query).  Those calls hand their
    arguments to the kernels through a mapped host window (no copy calls; small adds are not synchronised -- the next
    search is ordered behind them).  Same answers as the copy path (PRGPU_NO_HOST_WINDOW) and the oracle: 1300 adds
    in a row (the window's ring of 512 slots wraps twice, the tree buffer regrows), then adds and searches interleaved."""
    if not window:
        monkeypatch.setenv("PRGPU_NO_HOST_WINDOW", "1")
    rng = np.random.default_rng(77)
    pts = rng.uniform(-np.pi, np.pi, size=(1500, 7))
    pts[700:720] = pts[100:120]                                   # duplicates: ties by insertion index
    q = rng.uniform(-np.pi, np.pi, size=(64, 7))
    tree = P.KnnIndex(7)
    for i in range(1300):
        tree.add(pts[i:i + 1])
    gi, gd = tree.nearest_k(q, 5)                                 # 64 queries: still a window call
    oi, od = O.knn(pts[:1300], q, 5, nthreads=4)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    for i in range(1300, 1500):
        tree.add(pts[i:i + 1])
        if i % 3 == 0:
            tree.add(pts[i:i + 1] + 1e-3)                         # an extra node now and then (both go to the oracle)
        qq = q[i % 64:i % 64 + 1]
        gi, gd = tree.nearest_k(qq, 3)
        have = tree.points(0, tree.size())
        oi, od = O.knn(have, qq, 3, nthreads=1)
        assert np.array_equal(gi, oi) and np.array_equal(gd, od), i
    # a foreign stream after small adds: the appends are waited for
    import torch
    tree.add(pts[:3] + 0.5)
    have = np.vstack([have, pts[:3] + 0.5])
    s = torch.cuda.Stream()
    qd = torch.from_numpy(q).cuda()
    i_ = torch.empty((64, 4), dtype=torch.int32, device="cuda")
    d_ = torch.empty((64, 4), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    tree.nearest_k_device(qd, 64, 4, i_, d_, stream=s.cuda_stream)
    s.synchronize()
    oi, od = O.knn(have, q, 4, nthreads=4)
    assert np.array_equal(i_.cpu().numpy().astype(np.uint32), oi) and np.array_equal(d_.cpu().numpy(), od)
    tree.close()


def test_two_batches_in_flight_on_two_streams(P, O):
    """Indexed searches use no per-handle scratch: two batches issued on two streams at once (include/prgpu.h,
    "Batches in flight") give what each gives alone."""
    import torch
    rng = np.random.default_rng(5150)
    n = 1 << 18
    pts = rng.uniform(-np.pi, np.pi, size=(n, 7))
    tree = P.KnnIndex(7)
    tree.add(pts)
    tree.set_search_mode(3)
    qs = [rng.uniform(-np.pi, np.pi, size=(1500, 7)) for _ in range(2)]
    want = [O.knn(pts, q, 16, nthreads=8) for q in qs]
    tree.nearest_k(qs[0][:64], 16)                                # builds the index
    assert tree.index_stats()["indexed_nodes"] == n
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    qd = [torch.from_numpy(q).cuda() for q in qs]
    out = [(torch.empty((1500, 16), dtype=torch.int32, device="cuda"),
            torch.empty((1500, 16), dtype=torch.float64, device="cuda")) for _ in range(2)]
    torch.cuda.synchronize()
    for rep in range(6):                                          # alternately, never waiting in between
        j = rep & 1
        tree.nearest_k_device(qd[j], 1500, 16, out[j][0], out[j][1], stream=streams[j].cuda_stream)
    torch.cuda.synchronize()
    for j in range(2):
        assert np.array_equal(out[j][0].cpu().numpy().astype(np.uint32), want[j][0])
        assert np.array_equal(out[j][1].cpu().numpy(), want[j][1])
    tree.close()


def test_pinned_result_buffers_are_written_by_the_search_itself(P, O):
    """prgpu_knn_nearest_k with PINNED host result buffers and an indexed search: the kernel writes the rows into
    the caller's memory (zero copy).  Same rows as with pageable buffers (the copy path) and the oracle."""
    import torch
    rng = np.random.default_rng(99)
    n, nq, k = 1 << 18, 1000, 16
    pts = rng.uniform(-np.pi, np.pi, size=(n, 7))
    q = rng.uniform(-np.pi, np.pi, size=(nq, 7))
    tree = P.KnnIndex(7)
    tree.add(pts)
    tree.set_search_mode(3)
    gi, gd = tree.nearest_k(q, k)                                 # numpy (pageable) buffers; builds the index
    assert tree.index_stats()["indexed_nodes"] == n
    oi, od = O.knn(pts, q, k, nthreads=8)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)
    qp = torch.from_numpy(q).pin_memory()
    ip = torch.full((nq, k), -1, dtype=torch.int32).pin_memory()
    dp = torch.full((nq, k), -1.0, dtype=torch.float64).pin_memory()
    for with_dist in (True, False):
        ip.fill_(-1)
        P.capi._check(P.lib().prgpu_knn_nearest_k(tree.h, P.capi._ptr(qp), nq, k, None, P.capi._ptr(ip),
                                                  P.capi._ptr(dp) if with_dist else None))
        assert np.array_equal(ip.numpy().astype(np.uint32), oi)
    assert np.array_equal(dp.numpy(), od)
    tree.close()
