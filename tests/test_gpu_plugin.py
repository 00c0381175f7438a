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
