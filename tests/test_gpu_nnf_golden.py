"""K3 parity against the REFERENCE's own NNF planner: tests/golden/nnf_golden.npz was produced by the
reference's frechet/src/*.cpp compiled unmodified (tests/golden/make_golden_nnf.py).  The device path
(prgpu_nnf_* through the C ABI, and pr_gpu::NNFrechet through the plugin library) must reproduce, bit for
bit, the findKnnNodes choices and interpolated configurations; poses to 1e-12; the first search's g[goal]
and LazySP's final g[goal] within 1e-5 relative (north star; ~1e-15 achieved); the solved flag; and must
return a valid, collision-free, optimal path (the path itself depends on upstream's hash-set order)."""
import numpy as np
import pytest

import nnf_golden_util as G

pytestmark = pytest.mark.gpu
CASES = G.load()


def _worlds(P, O, synth, c):
    dh, link, xyzr = synth.default_arm()
    sph, box = G.case_world_arrays(synth, c)
    return P.World(dh, link, xyzr, sph, box), O.World(dh, link, xyzr, sph, box)


def _close(a, b, rel=1e-5):
    return abs(a - b) <= rel * max(abs(b), 1e-300)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("name", list(CASES))
def test_device_nnf_matches_reference_goldens(P, O, synth, name, mode):
    c = CASES[name]
    tseed, pseed, W, M, k, D, ns, nb = [int(x) for x in c["params"]]
    gw, ow = _worlds(P, O, synth, c)
    g = P.Nnf(gw, c["ref_sub"], c["ik_call_counts"], c["ik"], num_nn=k, discretization=D)
    g.set_mode(mode)
    R, V, E, n_ik = [int(x) for x in c["sizes"][:4]]
    assert (g.n_vertices, g.n_edges, g.n_ik) == (V, E, n_ik)
    assert np.array_equal(g.knn(), c["knn"])
    st, po = g.vertices()
    assert np.array_equal(st, c["states"])
    assert np.allclose(po, c["pose"], rtol=0, atol=1e-12)
    r = g.solve(max_lazy_iters=4000)
    assert _close(r["goal_cost_trace"][0], float(c["first_goal_cost"]))
    assert _close(r["goal_cost_trace"][0], float(c["first_goal_cost"]), 1e-12)
    assert r["solved"] == int(c["solved"])
    if int(c["solved"]):
        assert _close(r["goal_cost"], float(c["goal_cost"])) and _close(r["goal_cost"], float(c["goal_cost"]), 1e-12)
        assert r["final_error"] <= r["goal_cost"] * (1 + 1e-12)
        G.check_solution_path(c, r["path"], ow)
        if c["final_error"][0] == c["final_error"][1]:
            assert _close(r["final_error"], c["final_error"][0])
    else:
        assert r["goal_cost"] >= 1e300 and float(c["goal_cost"]) == G.DBL_MAX


@pytest.mark.parametrize("name", ["defaults_s4", "defaults_free", "w12m20_seed7", "w20m30"])
def test_plugin_nnfrechet_matches_reference_goldens(P, O, synth, name):
    """The drop-in planner class takes what the reference takes (full reference path, W, M, k, D, seed, IK
    callback): sub-sampling and the IK budget happen in the plugin and must match the reference's."""
    import plugin_lib as PL
    c = CASES[name]
    tseed, pseed, W, M, k, D, ns, nb = [int(x) for x in c["params"]]
    dh, link, xyzr = synth.default_arm()
    sph, box = G.case_world_arrays(synth, c)
    pw = PL.PluginWorld(dh, link, xyzr, sph, box)
    ow = O.World(dh, link, xyzr, sph, box)
    g = pw.nnf_run(W, M, k, D, pseed, c["ref_full"], c["ik"], solve_seconds=60.0)
    assert np.array_equal(g["layer_counts"], c["ik_call_counts"])
    solved = g["status"] == 6   # ompl::base::PlannerStatus::EXACT_SOLUTION
    assert int(solved) == int(c["solved"])
    if solved:
        assert _close(g["goal_cost"], float(c["goal_cost"]), 1e-12)
        G.check_solution_path(c, g["path"], ow)
        assert np.array_equal(g["path_states"], c["states"][g["path"]])


def _problem(O, synth, dh, seed, W, M, D):
    R = (W - 1) * (D + 1) + 1
    q, poses = synth.nnf_reference(seed, 2 * R + 5, dh)
    ids = O.lin_int_space(0, len(q) - 1, R)
    counts = O.nnf_layer_counts(1, R, W, M)
    ik = synth.nnf_ik_states(2, q[ids], counts)
    return poses[ids], counts, ik


def test_nnf_knn_takes_filter_path_and_matches_oracle(P, O, synth):
    """findKnnNodes over >= 8192 IK nodes with Gaussian (not fp32-representable) states: the device takes
    K1's tensor-core filter path; choices must still equal the oracle's full-sort findKnnNodes."""
    dh, link, xyzr = synth.default_arm()
    gw = P.World(dh, link, xyzr, np.zeros((0, 4)), np.zeros((0, 6)))
    ow = O.World(dh, link, xyzr, np.zeros((0, 4)), np.zeros((0, 6)))
    W, M, k, D = 80, 128, 5, 1
    ref, counts, ik = _problem(O, synth, dh, 6, W, M, D)
    assert len(ik) == 10240 and not np.array_equal(ik, ik.astype(np.float32).astype(np.float64))
    g = P.Nnf(gw, ref, counts, ik, num_nn=k, discretization=D)
    stats = g.knn_stats()
    assert stats["filter_calls"] >= 1, stats            # the interior index (9984 nodes) was filtered
    o = O.nnf_run(ow, ref, counts, ik, k=k, D=D, max_lazy_iters=0, nthreads=8)
    assert np.array_equal(g.knn(), o["knn"])
    st, po = g.vertices()
    assert np.array_equal(st, o["states"]) and np.allclose(po, o["pose"], rtol=0, atol=1e-12)
    # and the first search's cost against the independent reachability check
    r = g.solve(max_lazy_iters=1)
    ed, _ = g.edges()
    c = r["goal_cost_trace"][0]
    assert O.nnf_reachable(ref, po, ed, c) and not O.nnf_reachable(ref, po, ed, c, strict=True)


def test_nnf_full_size_costs_vs_independent_reachability(P, O, synth, world_arrays):
    """BASELINE configs[3]: W=200 waypoints, M=1024 IK samples per layer, 7-DOF arm (k=5, D=1): 399 x 1.2e6
    tensor-product cells.  The device's first-search cost and LazySP's final cost must be exactly the
    thresholds at which the goal becomes reachable (independent algorithm, no DP recurrence); every
    knocked-out edge must collide and every path edge be free under the specification's evaluateEdge."""
    dh, link, xyzr, sph, box = world_arrays
    gw = P.World(dh, link, xyzr, sph, box)
    ow = O.World(dh, link, xyzr, sph, box)
    W, M, k, D = 200, 1024, 5, 1
    ref, counts, ik = _problem(O, synth, dh, 4, W, M, D)
    g = P.Nnf(gw, ref, counts, ik, num_nn=k, discretization=D)
    assert g.n_ik == W * M and g.knn_stats()["filter_calls"] >= 1
    st, po = g.vertices()
    ed, _ = g.edges()
    r = g.solve(max_lazy_iters=20000)
    first = r["goal_cost_trace"][0]
    assert O.nnf_reachable(ref, po, ed, first) and not O.nnf_reachable(ref, po, ed, first, strict=True)
    _, fl = g.edges()
    blocked = (fl & 2) != 0
    assert blocked.sum() == r["n_collisions"] and ((fl & 1) != 0).sum() == r["n_evaluated"]
    if r["solved"]:
        final = r["goal_cost"]
        assert O.nnf_reachable(ref, po, ed, final, blocked=blocked)
        assert not O.nnf_reachable(ref, po, ed, final, strict=True, blocked=blocked)
        path = r["path"].astype(np.int64)
        ok, clear, _ = ow.check_motion_batch(st[path[:-1]], st[path[1:]], 0.05, mode=1, nthreads=8)
        assert np.all((ok == 1) | (np.abs(clear) < 1e-6))
        assert G.coupling_cost(ref, po, path) == final
    b = np.where(blocked)[0]
    ok, clear, _ = ow.check_motion_batch(st[ed[b, 0]], st[ed[b, 1]], 0.05, mode=1, nthreads=8)
    assert np.all((ok == 0) | (np.abs(clear) < 1e-6))


# ---- host-callback form: user FK / metric / validity, device for kNN + search ----------
def _rot_metric_weights(ref, pose):
    """weights[v, r] under the test metric: translation distance + 0.25 x Frobenius norm of the rotation
    difference (same operation order as plugin_capi.cpp / ref_nnf_driver.cpp)."""
    V, R = len(pose), len(ref)
    w = np.empty((V, R))
    rot = [0, 1, 2, 4, 5, 6, 8, 9, 10]
    for r in range(R):
        d = pose[:, [3, 7, 11]] - ref[r, [3, 7, 11]]
        t = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2])
        f = np.zeros(V)
        for c in rot:
            e = ref[r, c] - pose[:, c]
            f = f + e * e
        w[:, r] = t + 0.25 * np.sqrt(f)
    return w


@pytest.mark.parametrize("name", ["defaults_s4", "w12m20_seed7", "w20m30"])
def test_host_form_with_translation_weights_equals_device_form(P, O, synth, name):
    """prgpu_nnf_create_graph + set_weights + set_edge_evaluator with the translation metric evaluated on the
    HOST and the oracle's evaluateEdge as the validity callback must walk exactly the LazySP sequence of the
    device form (same costs after every search, same path)."""
    c = CASES[name]
    tseed, pseed, W, M, k, D, ns, nb = [int(x) for x in c["params"]]
    gw, ow = _worlds(P, O, synth, c)
    dev = P.Nnf(gw, c["ref_sub"], c["ik_call_counts"], c["ik"], num_nn=k, discretization=D)
    st, po = dev.vertices()
    h = P.Nnf.host_form(7, c["ik_call_counts"], c["ik"], num_nn=k, discretization=D)
    assert (h.n_vertices, h.n_edges, h.n_ik) == (dev.n_vertices, dev.n_edges, dev.n_ik)
    assert np.array_equal(h.knn(), c["knn"]) and np.array_equal(h.states(), c["states"])
    pose = np.array(po)
    pose[0], pose[1] = c["ref_sub"][0], c["ref_sub"][-1]     # the dummies carry the end poses
    ref = c["ref_sub"]
    d = ref[None, :, [3, 7, 11]] - pose[:, None, [3, 7, 11]]
    w = np.sqrt(d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2])
    h.set_weights(w)
    calls = []

    def evaluate(a, b):
        calls.append(len(a))
        v, _, _ = ow.check_motion_batch(a, b, 0.05, mode=1, want_clearance=False)
        return v
    h.set_edge_evaluator(evaluate)
    rd, rh = dev.solve(max_lazy_iters=4000), h.solve(max_lazy_iters=4000)
    assert rd["solved"] == rh["solved"] == int(c["solved"])
    assert rd["lazy_iters"] == rh["lazy_iters"]
    assert np.allclose(rd["goal_cost_trace"], rh["goal_cost_trace"], rtol=1e-12, atol=0)
    assert np.array_equal(rd["path"], rh["path"])
    assert abs(rd["final_error"] - rh["final_error"]) <= 1e-12
    assert sum(calls) == rh["n_evaluated"] or sum(calls) >= rh["n_evaluated"]


def test_host_form_errors(P, synth):
    h = P.Nnf.host_form(7, [2, 2], np.zeros((4, 7)), num_nn=2, discretization=1)
    with pytest.raises(P.PrgpuError):
        h.solve(max_lazy_iters=1)                       # no weights
    with pytest.raises(P.PrgpuError):
        h.set_weights(np.full((h.n_vertices, 2), -1.0))  # negative weights
    h.set_weights(np.zeros((h.n_vertices, 2)))
    with pytest.raises(P.PrgpuError):
        h.solve(max_lazy_iters=1)                       # no evaluator
    with pytest.raises(P.PrgpuError):
        P.Nnf.host_form(9, [2, 2], np.zeros((4, 9)))     # dof out of range


@pytest.mark.parametrize("name,with_world", [("defaults_s4", False), ("w12m20_seed7", True), ("w20m30", False)])
def test_plugin_host_callbacks_rotation_metric_vs_compiled_reference(P, O, synth, name, with_world):
    """NNFrechet::NNFrechet (drop-in name) with a task metric the device has no model of (rotation term): the
    planner must take the host-callback form -- also when a world was given -- and reproduce the compiled
    reference run with the same callbacks: IK budget, final g[goal], solved flag, a valid path."""
    import plugin_lib as PL
    if O.ref_nnf() is None:
        pytest.skip("oracle/_ref/libref_frechet.so not built")
    c = CASES[name]
    tseed, pseed, W, M, k, D, ns, nb = [int(x) for x in c["params"]]
    dh, link, xyzr = synth.default_arm()
    sph, box = G.case_world_arrays(synth, c)
    pw = PL.PluginWorld(dh, link, xyzr, sph, box)
    ow = O.World(dh, link, xyzr, sph, box)
    ref = O.ref_nnf_run(ow, c["ref_full"], W, M, k, D, c["ik"], seed=pseed, solve_seconds=300.0, metric=1)
    g = pw.nnf_run_host(W, M, k, D, pseed, c["ref_full"], c["ik"], metric=1, with_world=with_world, solve_seconds=120.0)
    assert g["host_form"] == 1
    assert np.array_equal(g["layer_counts"], ref["ik_call_counts"])
    solved = g["status"] == 6
    assert int(solved) == int(ref["solved"])
    assert _close(g["goal_cost"], ref["goal_cost"], 1e-12)
    if solved:
        # a real c0 -> c1 path of the reference's NN graph, collision-free under the specification
        edges = set(map(tuple, ref["edges"].tolist()))
        path = [int(v) for v in g["path"]]
        assert (0, path[0]) in edges and (path[-1], 1) in edges
        assert all((u, v) in edges for u, v in zip(path[:-1], path[1:]))
        assert np.array_equal(g["path_states"], ref["states"][path])
        if len(path) > 1:
            v, _, _ = ow.check_motion_batch(ref["states"][path[:-1]], ref["states"][path[1:]], 0.05, mode=1,
                                            want_clearance=False)
            assert v.all()
        # optimal under the rotation metric: the best coupling of the path costs the reference's g[goal]
        pose = ref["pose"].copy()
        w = _rot_metric_weights(ref["ref"], pose)
        seq = [0] + path + [1]
        f = w[seq].T
        Rr, L = f.shape
        B = np.full((Rr, L), np.inf)
        B[0, 0] = f[0, 0]
        for r in range(Rr):
            for j in range(L):
                if r == 0 and j == 0:
                    continue
                best = min(B[r - 1, j] if r else np.inf, B[r, j - 1] if j else np.inf,
                           B[r - 1, j - 1] if r and j else np.inf)
                B[r, j] = max(best, f[r, j])
        assert _close(B[-1, -1], ref["goal_cost"], 1e-9)


def test_plugin_translation_metric_without_world_takes_host_form(P, O, synth):
    """No setWorld(): the user's FK / metric / validity checker are the only model -- the result must equal the
    device form's (golden) result."""
    import plugin_lib as PL
    c = CASES["w20m30"]
    tseed, pseed, W, M, k, D, ns, nb = [int(x) for x in c["params"]]
    dh, link, xyzr = synth.default_arm()
    sph, box = G.case_world_arrays(synth, c)
    pw = PL.PluginWorld(dh, link, xyzr, sph, box)
    ow = O.World(dh, link, xyzr, sph, box)
    g = pw.nnf_run_host(W, M, k, D, pseed, c["ref_full"], c["ik"], metric=0, with_world=False, solve_seconds=120.0)
    assert g["host_form"] == 1
    assert int(g["status"] == 6) == int(c["solved"])
    assert _close(g["goal_cost"], float(c["goal_cost"]), 1e-9)   # FK through the fp64 device kernel vs goldens
    if g["status"] == 6:
        G.check_solution_path(c, g["path"], ow)
