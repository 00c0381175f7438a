"""CPU tests pinning the oracle: hand-computed known answers for the restated OMPL
primitives, numpy float64 as a second independent check for kNN,
the oracle's RRTConnect port against the REFERENCE's own RRTConnect.cpp (oracle/_ref, when
built) and against golden vectors generated from it (tests/golden/rrtc_golden.npz)."""
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EXTS = {"ext-ext": (0, 0), "ext-con": (0, 1), "con-ext": (1, 0), "con-con": (1, 1)}


def test_distance_interpolate_known_answers(O):
    L = O.lib()
    a = np.zeros(7)
    b = np.ones(7)
    assert L.orc_distance(O._dp(a), O._dp(b), 7) == math.sqrt(7.0)
    out = np.empty(7)
    L.orc_interpolate(O._dp(a), O._dp(b), 0.25, 7, O._dp(out))
    assert np.all(out == 0.25)
    # from + (to-from)*t, not (1-t)*from + t*to: 0.1 + (0.3-0.1)*0.5
    x, y = np.full(7, 0.1), np.full(7, 0.3)
    L.orc_interpolate(O._dp(x), O._dp(y), 0.5, 7, O._dp(out))
    assert np.all(out == 0.1 + (0.3 - 0.1) * 0.5)


def test_knn_vs_numpy_and_tiebreak(O, synth):
    pts = synth.cloud(0, 3000, dup_fraction=0.2)
    q = synth.cloud(1, 64)
    q[:20] = pts[np.arange(20) * 7]
    for k in (1, 5, 16):
        idx, dist = O.knn(pts, q, k)
        d = np.sqrt(((pts[None, :, :] - q[:, None, :]) ** 2).sum(-1))  # different summation order
        for i in range(len(q)):
            dd = np.array([O.lib().orc_distance(O._dp(pts[j]), O._dp(q[i]), 7) for j in idx[i]])
            assert np.array_equal(dd, dist[i])
            order = np.lexsort((np.arange(len(pts)), d[i]))[:k]
            # same set up to float-rounding of numpy's pairwise sum; ties by lowest index
            assert np.allclose(np.sort(d[i][order]), np.sort(dist[i]), rtol=1e-12)
            for a_, b_ in zip(idx[i][:-1], idx[i][1:]):
                da, db = d[i][a_], d[i][b_]
                assert da < db or (abs(da - db) <= 1e-12 * max(da, 1) and (dist[i][list(idx[i]).index(a_)] <
                       dist[i][list(idx[i]).index(b_)] or a_ < b_))
    # exact duplicates: first inserted wins (NearestNeighborsLinear first strict minimum)
    z = np.zeros((10, 7))
    idx, _ = O.knn(z, np.zeros((1, 7)), 3)
    assert list(idx[0]) == [0, 1, 2]
    # suffix mask = findKnnNodes candidates (deeper layers only)
    idx, _ = O.knn(pts, q, 4, lo=np.full(len(q), 2990, dtype=np.uint32))
    assert idx.min() >= 2990
    idx, dist = O.knn(pts, q[:2], 4, lo=np.array([3000, 2999], dtype=np.uint32))
    assert np.all(idx[0] == 0xFFFFFFFF) and np.all(np.isinf(dist[0]))
    assert idx[1, 0] == 2999 and np.all(idx[1, 1:] == 0xFFFFFFFF)


def test_discrete_motion_validator_tested_set(O, oracle_world):
    # nd = ceil(dist / lvs); tested set = {to} U {j/nd, 0<j<nd}  (Appendix B.3)
    a = np.zeros((4, 7))
    b = np.zeros((4, 7))
    b[0, 0] = 0.0        # nd = 0 -> 1 state
    b[1, 0] = 0.005      # nd = 1 -> 1 state
    b[2, 0] = 0.0799     # nd = 8 -> 8 states
    b[3, 0] = 1.0        # nd = 100 -> 100 states
    _, _, nt = oracle_world.check_motion_batch(a, b, 0.01)
    assert list(nt) == [1, 1, 8, 100]
    _, _, nt = oracle_world.check_motion_batch(a, b, 0.01, check_first=True)
    assert list(nt) == [2, 2, 9, 101]


def test_nnf_alpha_accumulation_skips_endpoint(O, oracle_world):
    # Appendix C.9: for (alpha=0; alpha<=1; alpha+=1.0/numQ) skips alpha=1 for numQ = 9, 11, 18, 20, ...
    def tested(numQ):
        n, alpha, step = 0, 0.0, 1.0 / numQ
        while alpha <= 1.0:
            n += 1
            alpha += step
        return n
    skipped = [nq for nq in range(1, 40) if tested(nq) == nq]
    assert skipped[:6] == [9, 11, 18, 20, 21, 25]
    a = np.zeros((3, 7))
    b = np.zeros((3, 7))
    b[1, 0] = 0.05 * 8.5   # numQ = 9  -> 9 states (endpoint skipped)
    b[2, 0] = 0.05 * 9.5   # numQ = 10 -> 11 states
    _, _, nt = oracle_world.check_motion_batch(a, b, 0.05, mode=1)
    assert list(nt) == [1, tested(9), tested(10)] == [1, 9, 11]


def test_world_fk_known_pose(O, oracle_world):
    pose, cen = oracle_world.fk(np.zeros(7))
    assert np.allclose(pose, [[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 1.306]], atol=1e-15)
    q = np.zeros(7)
    q[1] = math.pi / 2  # shoulder pitch: arm horizontal at shoulder height
    pose, _ = oracle_world.fk(q)
    assert abs(pose[2, 3] - 0.36) < 1e-12 and abs(np.hypot(pose[0, 3], pose[1, 3]) - 0.946) < 1e-12


def test_sample_stream_is_counter_based(O):
    L = O.lib()
    a = L.orc_sample_uniform(1, 2, 3, 4, -1.0, 1.0)
    assert a == L.orc_sample_uniform(1, 2, 3, 4, -1.0, 1.0)
    assert a != L.orc_sample_uniform(1, 2, 3, 5, -1.0, 1.0)
    xs = np.array([L.orc_sample_uniform(0, 0, i, 0, -math.pi, math.pi) for i in range(4000)])
    assert -math.pi <= xs.min() and xs.max() < math.pi and abs(xs.mean()) < 0.15


def _same(r1, r2):
    return all(np.array_equal(r1[k], r2[k]) for k in r1)


def test_rrtconnect_port_equals_reference(O, synth, oracle_world):
    """The flat-array port and the reference's own RRTConnect.cpp (over ompl_min) build identical
    trees: same iteration count, parent arrays, states (bit for bit) and solution path."""
    if O.ref() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box); goldens cover it")
    from golden.make_golden import problems
    starts, goals = problems(synth, oracle_world, 6, seed=40)
    n = 0
    for qi in range(6):
        for name, ext in EXTS.items():
            for rng in (0.0, 0.7):
                kw = dict(rng=rng, ext=ext, seed=3, query=qi, max_iters=250)
                r1 = oracle_world.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, impl="port", **kw)
                r2 = oracle_world.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, impl="ref", **kw)
                assert _same(r1, r2), (qi, name, rng)
                n += r1["status"] == 6
    assert n > 20  # most of them are actually solved


@pytest.mark.parametrize("impl", ["port", "ref"])
def test_rrtconnect_goldens(O, oracle_world, impl):
    if impl == "ref" and O.ref() is None:
        pytest.skip("oracle/_ref not built on this box")
    g = np.load(os.path.join(HERE, "golden", "rrtc_golden.npz"))
    for qi in range(8):
        for name, ext in EXTS.items():
            for rng in (0.0, 0.9):
                r = oracle_world.rrtc_solve(g["starts"][qi], g["goals"][qi], -np.pi, np.pi, rng=rng, ext=ext,
                                            seed=11, query=qi, max_iters=400, impl=impl)
                key = "q%d_%s_r%s" % (qi, name, rng)
                assert [r["status"], r["iterations"]] == list(g[key + "_meta"]), key
                assert np.array_equal(r["start_parent"], g[key + "_sp"]), key
                assert np.array_equal(r["goal_parent"], g[key + "_gp"]), key
                assert np.array_equal(r["start_states"], g[key + "_ss"]), key
                assert np.array_equal(r["goal_states"], g[key + "_gs"]), key
                assert np.array_equal(r["path"], g[key + "_path"]), key


def test_rrtconnect_invalid_endpoints(O, synth, oracle_world):
    q = synth.cloud(77, 400)
    v, _ = oracle_world.is_valid_batch(q)
    bad, good = q[v == 0][0], q[v == 1][:2]
    for impl in (["port", "ref"] if O.ref() is not None else ["port"]):
        r = oracle_world.rrtc_solve(bad, good[0], -np.pi, np.pi, max_iters=50, impl=impl)
        assert r["status"] == 1 and len(r["start_parent"]) == 0          # INVALID_START
    r = oracle_world.rrtc_solve(good[0], bad, -np.pi, np.pi, max_iters=50, impl="port")
    assert r["status"] == 4 and len(r["goal_parent"]) == 0               # goal never sampled -> TIMEOUT
    r = oracle_world.rrtc_solve(good[0], good[0], -np.pi, np.pi, max_iters=50, impl="port")
    assert r["status"] == 6                                              # start == goal connects at once


def test_reference_resume_and_clear_semantics(O, synth, oracle_world):
    """What the GPU planner classes are held to, established on the reference itself: a second
    solve() continues the existing trees -- identical to one uninterrupted solve when the first stopped after
    an even number of samples (the tree alternation restarts with the start tree, RRTConnect.cpp:263-268) --
    and clear() + setup() + solve() equals a fresh planner (RRTConnect.cpp:166-176)."""
    ref = O.ref()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    q = synth.cloud(91, 200)
    v, _ = oracle_world.is_valid_batch(q, 8)
    qv = q[v == 1]
    try:
        for i in range(4):
            for ext in ((0, 1), (1, 1)):
                kw = dict(rng=0.6, ext=ext, seed=2, query=i, max_iters=120, impl="ref")
                ref.ref_set_two_phase(0, 0)
                single = oracle_world.rrtc_solve(qv[2 * i], qv[2 * i + 1], -np.pi, np.pi, **kw)
                for first in (2, 10, 40):
                    ref.ref_set_two_phase(first, 0)
                    resumed = oracle_world.rrtc_solve(qv[2 * i], qv[2 * i + 1], -np.pi, np.pi, **kw)
                    assert all(np.array_equal(single[k], resumed[k]) for k in single), (i, ext, first)
                ref.ref_set_two_phase(30, 1)
                cleared = oracle_world.rrtc_solve(qv[2 * i], qv[2 * i + 1], -np.pi, np.pi, **kw)
                assert all(np.array_equal(single[k], cleared[k]) for k in single), (i, ext, "clear")
    finally:
        ref.ref_set_two_phase(0, 0)


def test_kdtree_baseline_equals_linear_scan(O, synth):
    """The k-d tree CPU baseline (oracle/kdtree.cpp) returns orc_knn's answers bit for bit: ties, exact
    duplicates, k > n, clustered data."""
    rng = np.random.default_rng(0)
    for n, nq, dup in [(1, 4, 0.0), (23, 50, 0.0), (5000, 300, 0.2), (40000, 200, 0.02)]:
        pts = synth.cloud(7, n, dup_fraction=dup)
        q = np.vstack([synth.cloud(8, nq), pts[rng.integers(0, n, 8)]])      # some queries ARE tree points
        t = O.KdTree(pts)
        for k in (1, 5, 16):
            ki, kd = t.knn(q, k, nthreads=4)
            li, ld = O.knn(pts, q, k, nthreads=4)
            assert np.array_equal(ki, li) and np.array_equal(kd, ld), (n, k)
    # all points identical, and tight clusters far apart
    same = np.tile(synth.cloud(9, 1), (200, 1))
    t = O.KdTree(same)
    ki, kd = t.knn(synth.cloud(10, 5), 16)
    li, ld = O.knn(same, synth.cloud(10, 5), 16)
    assert np.array_equal(ki, li) and np.array_equal(kd, ld)
    cl = np.repeat(synth.cloud(11, 20), 100, axis=0) + 1e-9 * rng.normal(size=(2000, 7))
    t = O.KdTree(cl)
    qq = synth.cloud(12, 64)
    ki, kd = t.knn(qq, 16)
    li, ld = O.knn(cl, qq, 16)
    assert np.array_equal(ki, li) and np.array_equal(kd, ld)
