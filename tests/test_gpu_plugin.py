"""Drop-in parity through the OMPL plugin API: pr_ompl::RRTConnect (the GPU-backed planner under
the reference's class name), GpuNearestNeighbors, GpuStateValidityChecker and GpuMotionValidator
are driven the way an OMPL application drives the reference and must reproduce its trees."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXTS = {"ext-ext": (0, 0), "ext-con": (0, 1), "con-ext": (1, 0), "con-con": (1, 1)}


@pytest.fixture(scope="module")
def plugin_world(world_arrays):
    import plugin_lib
    return plugin_lib.PluginWorld(*world_arrays)


def _same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in b)


def test_planner_matches_reference_goldens(plugin_world):
    g = np.load(os.path.join(HERE, "golden", "rrtc_golden.npz"))
    for qi in range(8):
        for ext in EXTS:
            for rng in (0.0, 0.9):
                r, pd = plugin_world.rrtc_solve(g["starts"][qi], g["goals"][qi], -np.pi, np.pi, rng=rng,
                                                extension_type=ext, seed=11, query=qi, max_iters=400)
                key = "q%d_%s_r%s" % (qi, ext, rng)
                ref = dict(status=int(g[key + "_meta"][0]), iterations=int(g[key + "_meta"][1]),
                           start_parent=g[key + "_sp"], goal_parent=g[key + "_gp"], start_states=g[key + "_ss"],
                           goal_states=g[key + "_gs"], path=g[key + "_path"])
                assert _same(r, ref), key
                # getPlannerData: one vertex per node, one edge per non-root node (+ the connection)
                n = len(ref["start_parent"]) + len(ref["goal_parent"])
                assert pd[0] == n and pd[2] == 1 and pd[3] == 1
                assert pd[1] == n - 2 + (1 if ref["status"] == 6 else 0)


def test_planner_matches_oracle_port(plugin_world, oracle_world, synth):
    q = synth.cloud(91, 200)
    v, _ = oracle_world.is_valid_batch(q, 8)
    qv = q[v == 1]
    for i in range(6):
        for ext, code in EXTS.items():
            r, _ = plugin_world.rrtc_solve(qv[2 * i], qv[2 * i + 1], -np.pi, np.pi, rng=0.6, extension_type=ext,
                                           seed=2, query=i, max_iters=200)
            ref = oracle_world.rrtc_solve(qv[2 * i], qv[2 * i + 1], -np.pi, np.pi, rng=0.6, ext=code, seed=2,
                                          query=i, max_iters=200)
            assert _same(r, ref), (i, ext)


def test_bad_extension_string(plugin_world):
    with pytest.raises(RuntimeError):
        plugin_world.rrtc_solve(np.zeros(7), np.zeros(7), -1, 1, extension_type="con-bad")


def test_ompl_interfaces(plugin_world, oracle_world, O, synth):
    q = synth.cloud(92, 300)
    v, m, nn = plugin_world.selftest(q)
    ov, oc = oracle_world.is_valid_batch(q, 8)
    assert not np.any((v != ov) & (np.abs(oc) >= 1e-6))
    lvs = np.sqrt(7 * (2 * np.pi) ** 2) * 0.01
    om, ocl, _ = oracle_world.check_motion_batch(q[:-1], q[1:], lvs, nthreads=8)
    assert not np.any((m != om) & (np.abs(ocl) >= 1e-6))
    for i in range(1, 300, 13):
        idx, _ = O.knn(q[:i], q[i:i + 1], 1)
        assert nn[i] == idx[0, 0]


def test_nnfrechet_planner_matches_oracle(O, synth, world_arrays):
    """NNFrechet::NNFrechet (drop-in name) through its public API -- setRefPath / setFKFunc / setIKFunc /
    setDistanceFunc / setup / solve -- vs the oracle NNF port: same IK budget per waypoint (the
    std::default_random_engine draws), same path, same Frechet error."""
    import plugin_lib
    dh, link, xyzr, _, _ = world_arrays
    sph, box = synth.obstacles(3, n_spheres=32, n_boxes=32)
    pw = plugin_lib.PluginWorld(dh, link, xyzr, sph, box)
    ow = O.World(dh, link, xyzr, sph, box)
    for seed, W, M, k, D in [(4, 5, 5, 5, 3), (3, 12, 20, 5, 1), (4, 20, 30, 4, 2), (1, 12, 20, 5, 1)]:
        R = (W - 1) * (D + 1) + 1
        q, poses = synth.nnf_reference(seed, 2 * R + 5, dh)
        ids = O.lin_int_space(0, len(q) - 1, R)                 # subsampleRefPath
        counts = O.nnf_layer_counts(7, R, W, M)                 # planner seed 7
        ik = synth.nnf_ik_states(2, q[ids], counts)
        o = O.nnf_run(ow, poses[ids], counts, ik, k=k, D=D, max_lazy_iters=3000, nthreads=8)
        g = pw.nnf_run(W, M, k, D, 7, poses, ik)                 # full reference path: the planner sub-samples
        assert np.array_equal(g["layer_counts"], counts)
        assert (g["status"] == 6) == bool(o["solved"])
        assert g["lazy_iters"] == o["lazy_iters"]
        assert (g["n_evaluated"], g["n_collisions"]) == (o["n_evaluated"], o["n_collisions"])
        assert np.array_equal(g["path"], o["path"])
        if o["solved"]:
            assert abs(g["final_error"] - o["final_error"]) <= 1e-5 * o["final_error"]
            assert np.array_equal(g["path_states"], o["states"][o["path"]])
