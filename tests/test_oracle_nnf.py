"""CPU tests pinning the NNF oracle port (oracle/nnf_port.cpp) by independent means (the comparison
with the compiled reference and its golden vectors is tests/test_oracle_nnf_ref.py):
(1) the reference's documented quirks (linIntSpace truncation), (2) an
independent Python restatement of libstdc++'s minstd_rand0 + uniform_int_distribution for the IK
budget (frechet/src/NNFrechet.cpp:310-343), (3) an independent algorithm for the min-bottleneck
cost (threshold + reachability instead of the DP recurrence) on graphs rebuilt from the port's own
outputs, and (4) structural invariants of buildNNGraph."""
import numpy as np
import pytest


def _problem(O, synth, dh, seed, W, M, D):
    R = (W - 1) * (D + 1) + 1
    q, poses = synth.nnf_reference(seed, 2 * R + 5, dh)
    ids = O.lin_int_space(0, len(q) - 1, R)
    counts = O.nnf_layer_counts(1, R, W, M)
    ik = synth.nnf_ik_states(2, q[ids], counts)
    return poses[ids], counts, ik


def test_lin_int_space_truncation(O, synth):
    assert O.lin_int_space(0, 999, 797)[-1] == 998            # PROBED in SURVEY Appendix C.11
    assert list(O.lin_int_space(0, 9, 4)) == [0, 3, 6, 9]
    for n, p in [(400, 397), (50, 17), (23, 23)]:
        a = O.lin_int_space(0, n - 1, p)
        assert np.array_equal(a, synth.lin_int_space(0, n - 1, p))
        assert a[0] == 0 and np.all(np.diff(a) >= 0) and a[-1] in (n - 1, n - 2)


@pytest.mark.parametrize("seed,W,M,D", [(1, 5, 5, 3), (7, 12, 20, 1), (42, 30, 16, 2)])
def test_ik_budget_matches_libstdcxx_algorithm(O, synth, seed, W, M, D):
    R = (W - 1) * (D + 1) + 1
    c = O.nnf_layer_counts(seed, R, W, M)
    assert c.sum() == W * M and c[0] == M and c[-1] == M
    assert np.array_equal(c, synth.nnf_layer_counts(seed, R, W, M))      # pr-ompl_b200/synth.py restatement


def _bottleneck_by_threshold(ref, res, D, k):
    """Independent min-bottleneck: smallest tau such that the goal is reachable using only TPG cells with
    f <= tau (binary search over the sorted weights + BFS), on the NN graph rebuilt from knn/states."""
    R = len(ref)
    pose = res["pose"]
    V = len(pose)
    n_ik = res["n_ik"]
    f = np.sqrt(((ref[:, None, [3, 7, 11]] - pose[None, :, [3, 7, 11]]) ** 2).sum(-1))      # R x V
    succ = [[] for _ in range(V)]
    counts = res["_counts"]
    M0, ML = int(counts[0]), int(counts[-1])
    for i in range(M0):
        succ[0].append(2 + i)
    for i in range(ML):
        succ[2 + M0 + i].append(1)
    nxt = 2 + n_ik
    order = list(range(M0)) + list(range(M0 + ML, n_ik))
    for i in order:
        for nb in res["knn"][i]:
            if nb == 0xFFFFFFFF:
                break
            prev = 2 + i
            for d in range(D):
                succ[prev].append(nxt)
                prev = nxt
                nxt += 1
            succ[prev].append(int(nb))
    assert nxt == V
    vals = np.unique(f)

    def reachable(tau):
        ok = f <= tau
        if not ok[0, 0]:
            return False
        seen = np.zeros((R, V), dtype=bool)
        stack = [(0, 0)]
        seen[0, 0] = True
        while stack:
            r, v = stack.pop()
            if r == R - 1 and v == 1:
                return True
            cand = [(r, s) for s in succ[v]]
            if r + 1 < R:
                cand += [(r + 1, v)] + [(r + 1, s) for s in succ[v]]
            for rr, ss in cand:
                if ok[rr, ss] and not seen[rr, ss]:
                    seen[rr, ss] = True
                    stack.append((rr, ss))
        return False
    lo, hi = 0, len(vals) - 1
    while lo < hi:
        mid = (lo + hi) // 2
        if reachable(vals[mid]):
            hi = mid
        else:
            lo = mid + 1
    return vals[lo]


@pytest.mark.parametrize("seed,W,M,k,D", [(1, 3, 2, 2, 1), (4, 5, 5, 5, 3), (3, 6, 6, 3, 2)])
def test_bottleneck_cost_vs_independent_algorithm(O, synth, seed, W, M, k, D):
    dh, link, xyzr = synth.default_arm()
    free = O.World(dh, link, xyzr, np.zeros((0, 4)), np.zeros((0, 6)))          # no collisions: one search
    ref, counts, ik = _problem(O, synth, dh, seed, W, M, D)
    res = O.nnf_run(free, ref, counts, ik, k=k, D=D, max_lazy_iters=5)
    assert res["solved"] == 1 and res["lazy_iters"] == 1
    res["_counts"] = counts
    tau = _bottleneck_by_threshold(np.asarray(ref), res, D, k)
    assert res["goal_cost"] == tau                       # same doubles: max/min selections only
    assert res["final_error"] <= res["goal_cost"] + 1e-15   # mFinalError excludes dummy cells (C.17)


def test_nn_graph_structure(O, synth):
    dh, link, xyzr = synth.default_arm()
    free = O.World(dh, link, xyzr, np.zeros((0, 4)), np.zeros((0, 6)))
    W, M, k, D = 6, 5, 3, 2
    ref, counts, ik = _problem(O, synth, dh, 2, W, M, D)
    res = O.nnf_run(free, ref, counts, ik, k=k, D=D)
    n_ik = res["n_ik"]
    assert n_ik == W * M
    # vertex order: layer 0, last layer, interior layers; neighbours only in strictly deeper layers (C.13)
    R = len(ref)
    layer = np.concatenate([[0] * counts[0], [R - 1] * counts[-1]] +
                           [[r] * counts[r] for r in range(1, R - 1)]).astype(int)
    knn = res["knn"]
    for i in range(n_ik):
        for nb in knn[i]:
            if nb != 0xFFFFFFFF:
                assert layer[nb - 2] > layer[i]
    assert np.all(knn[counts[0]:counts[0] + counts[-1]] == 0xFFFFFFFF)      # last layer: no out-edges
    pairs = int((knn != 0xFFFFFFFF).sum())
    assert res["n_vertices"] == 2 + n_ik + pairs * D
    assert res["n_edges"] == counts[0] + counts[-1] + pairs * (D + 1)
    # sub-sampled states are first + (second-first)*(1+i)/(D+1), bit for bit (NNFrechet.cpp:364-369)
    st = res["states"]
    v = 2 + n_ik
    for i in list(range(counts[0])) + list(range(counts[0] + counts[-1], n_ik)):
        for nb in knn[i]:
            if nb == 0xFFFFFFFF:
                break
            for d in range(D):
                t = (1.0 + d) / (D + 1)
                assert np.array_equal(st[v], st[2 + i] + (st[nb] - st[2 + i]) * t)
                v += 1
