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
    pw = PL.World(dh, link, xyzr, sph, box)
    ow = O.World(dh, link, xyzr, sph, box)
    g = pw.nnf_run(W, M, k, D, pseed, c["ref_full"], c["ik"], solve_seconds=60.0)
    assert np.array_equal(g["layer_counts"], c["ik_call_counts"])
    solved = g["status"] == 6   # ompl::base::PlannerStatus::EXACT_SOLUTION
    assert int(solved) == int(c["solved"])
    if solved:
        assert _close(g["goal_cost"], float(c["goal_cost"]), 1e-12)
        G.check_solution_path(c, g["path"], ow)
        assert np.array_equal(g["path_states"], c["states"][g["path"]])
