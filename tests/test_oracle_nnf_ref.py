"""Pins the NNF oracle port (oracle/nnf_port.cpp) against the REFERENCE ITSELF:
 (1) live, where oracle/_ref/libref_frechet.so exists -- the reference's own frechet/src/{NNFrechet,LPAStar,
     MinPriorityQueue,util}.cpp compiled unmodified over ompl_min + the Boost.Graph stand-in of oracle/compat/;
 (2) against tests/golden/nnf_golden.npz, generated from that build by tests/golden/make_golden_nnf.py.
Compared bit for bit: the sub-sampled reference path (linIntSpace truncation), the IK budget per waypoint
(std::default_random_engine draws), NN-graph sizes, findKnnNodes choices, every vertex state and pose, the
first search's bottleneck cost g[goal], the solved flag and the final g[goal] of LazySP.  Which of several
equally good paths LPA* keeps depends on the iteration order of Boost's hash_setS edge lists
(implementation-defined; the two stand-in orders disagree on it), so the path, the evaluated-edge counts and
mFinalError (it excludes dummy cells, NNFrechet.cpp:629-630, and is therefore path-dependent) are checked
through path validity + optimality instead."""
import numpy as np
import pytest

import nnf_golden_util as G

CASES = G.load()
FAST = [n for n in CASES if CASES[n]["sizes"][1] < 3000]


def _port_inputs(O, c):
    tseed, pseed, W, M, k, D, ns, nb = [int(x) for x in c["params"]]
    R = (W - 1) * (D + 1) + 1
    ids = O.lin_int_space(0, len(c["ref_full"]) - 1, R)
    counts = O.nnf_layer_counts(pseed, R, W, M)
    return W, M, k, D, pseed, ids, counts


def _world(O, synth, c):
    dh, link, xyzr = synth.default_arm()
    sph, box = G.case_world_arrays(synth, c)
    return O.World(dh, link, xyzr, sph, box)


@pytest.mark.parametrize("name", list(CASES))
def test_port_matches_reference_goldens(O, synth, name):
    c = CASES[name]
    W, M, k, D, pseed, ids, counts = _port_inputs(O, c)
    ow = _world(O, synth, c)
    # subsampleRefPath / linIntSpace and the IK budget, as the reference computed them
    assert np.array_equal(c["ref_full"][ids], c["ref_sub"])
    assert np.array_equal(counts, c["ik_call_counts"])
    o = O.nnf_run(ow, c["ref_sub"], counts, c["ik"], k=k, D=D, max_lazy_iters=4000, nthreads=8)
    R, V, E, n_ik = [int(x) for x in c["sizes"][:4]]
    assert (len(c["ref_sub"]), o["n_vertices"], o["n_edges"], o["n_ik"]) == (R, V, E, n_ik)
    assert np.array_equal(o["knn"], c["knn"])
    assert np.array_equal(o["states"], c["states"])
    assert np.array_equal(o["pose"], c["pose"])
    assert o["goal_cost_trace"][0] == float(c["first_goal_cost"])
    assert o["solved"] == int(c["solved"])
    assert o["goal_cost"] == float(c["goal_cost"])
    if o["solved"]:
        assert o["final_error"] <= o["goal_cost"]
        G.check_solution_path(c, o["path"], ow)
        if c["final_error"][0] == c["final_error"][1]:      # both edge orders of the reference agree
            assert o["final_error"] == c["final_error"][0]


def test_golden_reference_paths_are_valid(O, synth):
    """The stored reference solution itself passes the same path checks the port and the GPU must pass."""
    for name in FAST:
        c = CASES[name]
        if not int(c["solved"]):
            continue
        st = c["states"]
        ids = []
        for s in c["path_states_order_dependent"]:
            hit = np.where((st[2:] == s).all(axis=1))[0]
            assert len(hit) >= 1
            ids.append(2 + int(hit[0]))
        G.check_solution_path(c, ids, _world(O, synth, c))


@pytest.mark.parametrize("name", FAST)
def test_port_matches_live_reference(O, synth, name):
    if O.ref_nnf() is None:
        pytest.skip("oracle/_ref/libref_frechet.so not built (needs /root/reference at build time)")
    c = CASES[name]
    W, M, k, D, pseed, ids, counts = _port_inputs(O, c)
    ow = _world(O, synth, c)
    for rev in (False, True):
        r = O.ref_nnf_run(ow, c["ref_full"], W, M, k, D, c["ik"], seed=pseed, reverse_order=rev)
        assert np.array_equal(r["ref"], c["ref_sub"]) and np.array_equal(r["ik_call_counts"], counts)
        assert np.array_equal(r["knn"], c["knn"]) and np.array_equal(r["states"], c["states"])
        assert r["goal_cost"] == float(c["goal_cost"]) and r["solved"] == int(c["solved"])
        # TPG size: R x V nodes; per NN edge (u,v) 2R-1 product edges, per vertex R-1 reference steps
        R, V, E = [int(x) for x in c["sizes"][:3]]
        assert r["n_tpg_vertices"] == R * V
        assert r["n_tpg_edges"] == V * (R - 1) + E * (2 * R - 1)


def test_reference_random_problems_vs_port(O, synth):
    """Fresh problems (not in the goldens), reference vs port, including the LazySP end state."""
    if O.ref_nnf() is None:
        pytest.skip("oracle/_ref/libref_frechet.so not built")
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(9, n_spheres=24, n_boxes=24)
    ow = O.World(dh, link, xyzr, sph, box)
    for tseed, pseed, W, M, k, D in [(11, 3, 6, 7, 4, 2), (12, 5, 9, 9, 3, 1), (13, 2, 4, 12, 6, 3), (14, 9, 7, 5, 2, 2)]:
        R = (W - 1) * (D + 1) + 1
        q, poses = synth.nnf_reference(tseed, 3 * R + 1, dh)
        ids = O.lin_int_space(0, len(q) - 1, R)
        counts = O.nnf_layer_counts(pseed, R, W, M)
        ik = synth.nnf_ik_states(tseed, q[ids], counts)
        r = O.ref_nnf_run(ow, poses, W, M, k, D, ik, seed=pseed)
        f = O.ref_nnf_run(ow, poses, W, M, k, D, ik, seed=pseed, search_only=True)
        o = O.nnf_run(ow, poses[ids], counts, ik, k=k, D=D, max_lazy_iters=4000)
        assert np.array_equal(r["ik_call_counts"], counts) and np.array_equal(r["ref"], poses[ids])
        assert np.array_equal(r["knn"], o["knn"]) and np.array_equal(r["states"], o["states"])
        assert np.array_equal(r["pose"], o["pose"])
        assert f["goal_cost"] == o["goal_cost_trace"][0]
        assert (r["solved"], r["goal_cost"]) == (o["solved"], o["goal_cost"])


@pytest.mark.parametrize("name", list(CASES))
def test_reachability_check_agrees_with_reference_costs(O, name):
    """orc_nnf_reachable (used for the full-size GPU test) brackets the reference's own costs exactly."""
    c = CASES[name]
    first = float(c["first_goal_cost"])
    assert O.nnf_reachable(c["ref_sub"], c["pose"], c["edges"], first)
    assert not O.nnf_reachable(c["ref_sub"], c["pose"], c["edges"], first, strict=True)
    blocked = (c["edge_flags_order_dependent"] & 2) != 0
    final = float(c["goal_cost"])
    if int(c["solved"]):
        assert O.nnf_reachable(c["ref_sub"], c["pose"], c["edges"], final, blocked=blocked)
        assert not O.nnf_reachable(c["ref_sub"], c["pose"], c["edges"], final, strict=True, blocked=blocked)
    else:
        assert not O.nnf_reachable(c["ref_sub"], c["pose"], c["edges"], 1e300, blocked=blocked)
