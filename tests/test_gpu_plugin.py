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


# ---- Drop-in A: the reference's own RRTConnect.cpp (unmodified) over the GPU interfaces ----------------

@pytest.fixture(scope="module")
def dropin_world(world_arrays):
    import plugin_lib
    L = plugin_lib.dropin()
    if L is None:
        pytest.skip("oracle/_ref/libdropin_a.so not built (needs /root/reference at build time)")
    return plugin_lib.PluginWorld(*world_arrays, library=L)


@pytest.mark.parametrize("nn_path", [0, 1])
def test_dropin_a_reference_planner_over_gpu_interfaces(dropin_world, nn_path):
    """pr_ompl::RRTConnect exactly as shipped by the reference, with GpuNearestNeighbors<Motion*> obtained
    from SelfConfig::getDefaultNearestNeighbors (nn_path 0) or setNearestNeighbors<NN>() (nn_path 1) and the
    GPU validity checker / motion validator underneath, reproduces the golden trees recorded from the same
    source over the CPU defaults."""
    import plugin_lib
    plugin_lib.dropin().dropin_set_nn_path(nn_path)
    g = np.load(os.path.join(HERE, "golden", "rrtc_golden.npz"))
    for qi in range(8):
        for ext in EXTS:
            for rng in ((0.0, 0.9) if nn_path == 0 else (0.9,)):
                r, pd = dropin_world.rrtc_solve(g["starts"][qi], g["goals"][qi], -np.pi, np.pi, rng=rng,
                                                extension_type=ext, seed=11, query=qi, max_iters=400)
                key = "q%d_%s_r%s" % (qi, ext, rng)
                ref = dict(status=int(g[key + "_meta"][0]), iterations=int(g[key + "_meta"][1]),
                           start_parent=g[key + "_sp"], goal_parent=g[key + "_gp"], start_states=g[key + "_ss"],
                           goal_states=g[key + "_gs"], path=g[key + "_path"])
                assert _same(r, ref), key
    plugin_lib.dropin().dropin_set_nn_path(0)


def test_dropin_a_single_planner_rate_is_reported(dropin_world, oracle_world, synth):
    """One planner instance = one launch + sync per growTree: state the rate instead of hiding it."""
    q = synth.cloud(4, 400)
    v, _ = oracle_world.is_valid_batch(q, 8)
    qv = q[v == 1]
    starts, goals = qv[0:32:2], qv[1:32:2]
    sec, solved, iters = dropin_world.rrtc_bench(starts, goals, -np.pi, np.pi, max_iters=200)
    assert sec > 0 and iters >= len(starts)
    print("drop-in A: %.1f plans/s, %.0f us per sample, %d/%d solved" %
          (len(starts) / sec, 1e6 * sec / iters, solved, len(starts)))


@pytest.mark.parametrize("which", ["plugin", "dropin"])
def test_second_solve_resumes_and_clear_restarts(which, plugin_world, world_arrays, oracle_world, synth):
    """SURVEY §8(f)1 (semantics pinned on the reference in tests/test_oracle.py::
    test_reference_resume_and_clear_semantics): a second solve() continues the device-resident trees; clear()
    empties them (prgpu_knn_clear) and the next solve() equals a fresh planner's."""
    import plugin_lib
    if which == "dropin":
        if plugin_lib.dropin() is None:
            pytest.skip("oracle/_ref/libdropin_a.so not built")
        w = plugin_lib.PluginWorld(*world_arrays, library=plugin_lib.dropin())
    else:
        w = plugin_world
    q = synth.cloud(91, 200)
    v, _ = oracle_world.is_valid_batch(q, 8)
    qv = q[v == 1]
    try:
        for i in range(3):
            for ext, code in (("ext-con", (0, 1)), ("con-con", (1, 1))):
                single = oracle_world.rrtc_solve(qv[2 * i], qv[2 * i + 1], -np.pi, np.pi, rng=0.6, ext=code, seed=2,
                                                 query=i, max_iters=120)
                for first, clear in ((2, False), (40, False), (30, True)):
                    w.set_two_phase(first, clear)
                    r, pd = w.rrtc_solve(qv[2 * i], qv[2 * i + 1], -np.pi, np.pi, rng=0.6, extension_type=ext, seed=2,
                                         query=i, max_iters=120)
                    assert _same(r, single), (which, i, ext, first, clear)
                    n = len(single["start_parent"]) + len(single["goal_parent"])
                    assert pd[0] == n            # getPlannerData after the second solve: every node once
    finally:
        w.set_two_phase(0, False)


def test_batched_planner_class_through_plugin_library(plugin_world, oracle_world, synth):
    """pr_gpu::BatchedRRTConnect (plugin/pr_gpu/BatchedRRTConnect.h) through libprgpu_plugin.so: every query
    equals what one reference planner instance does with CounterSampler(seed, base + i)."""
    q = synth.cloud(4, 600)
    v, _ = oracle_world.is_valid_batch(q, 8)
    qv = q[v == 1]
    n = 48
    starts, goals = qv[0:2 * n:2], qv[1:2 * n:2]
    for ext, code in (("ext-con", (0, 1)), ("con-ext", (1, 0))):
        b = plugin_world.batched_rrtc_solve(starts, goals, -np.pi, np.pi, rng=0.0, extension_type=ext, seed=5,
                                            query_base=100, max_iters=300)
        for i in range(n):
            o = oracle_world.rrtc_solve(starts[i], goals[i], -np.pi, np.pi, rng=0.0, ext=code, seed=5, query=100 + i,
                                        max_iters=300)
            assert (b["status"][i], b["iterations"][i]) == (o["status"], o["iterations"]), (ext, i)
            assert (b["n_start"][i], b["n_goal"][i]) == (len(o["start_parent"]), len(o["goal_parent"]))
            assert np.array_equal(b["paths"][i], o["path"]), (ext, i)
    with pytest.raises(RuntimeError):
        plugin_world.batched_rrtc_solve(starts[:2], goals[:2], -np.pi, np.pi, extension_type="bad")
