"""T1 parity: the device-resident batched RRTConnect (a warp per query, then speculative 8-warp teams for
the queries still unsolved after 256 samples) must build, node for
node, the trees of the reference planner: same GrowState sequence, hence same iteration count,
parent arrays, states (bit for bit) and solution path.  Checked against the oracle port on
seeded problems and against golden vectors recorded from the reference's own RRTConnect.cpp."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
EXTS = ["ext-ext", "ext-con", "con-ext", "con-con"]
EXT_CODE = {"ext-ext": (0, 0), "ext-con": (0, 1), "con-ext": (1, 0), "con-con": (1, 1)}


def _problems(synth, oracle_world, n, seed):
    q = synth.cloud(seed, 8 * n)
    v, _ = oracle_world.is_valid_batch(q, 8)
    qv = q[v == 1]
    return qv[0:2 * n:2].copy(), qv[1:2 * n:2].copy()


def _compare(rb, summ, qi, ref):
    assert summ["status"][qi] == ref["status"], qi
    assert summ["iterations"][qi] == ref["iterations"], qi
    ns, ng = int(summ["n_start"][qi]), int(summ["n_goal"][qi])
    assert ns == len(ref["start_parent"]) and ng == len(ref["goal_parent"]), qi
    got = rb.fetch(qi, ns, ng)
    for key in ("start_parent", "goal_parent", "start_states", "goal_states", "path"):
        assert np.array_equal(got[key], ref[key]), (qi, key)


@pytest.mark.parametrize("ext", EXTS)
@pytest.mark.parametrize("rng", [0.0, 0.8])
def test_batch_matches_oracle(P, O, synth, gpu_world, oracle_world, ext, rng):
    n = 48
    starts, goals = _problems(synth, oracle_world, n, seed=50)
    rb = P.RrtcBatch(gpu_world, n, -np.pi, np.pi, rng=rng, extension_type=ext, seed=5, query_id_base=100,
                     max_iters=300, max_nodes=4096)
    rb.run(starts, goals)
    summ = rb.summary()
    assert (summ["status"] == 6).sum() > 0
    for qi in range(n):
        ref = oracle_world.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, rng=rng, ext=EXT_CODE[ext], seed=5,
                                      query=100 + qi, max_iters=300, impl="port")
        _compare(rb, summ, qi, ref)


@pytest.mark.parametrize("ext", EXTS)
def test_long_budget_goes_through_the_team_kernel(P, O, synth, gpu_world, oracle_world, ext):
    # 1000 samples: the queries that are not solved by sample 256 are suspended and finished by the team kernel
    # (speculative evaluation of consecutive samples); iteration counts, trees and paths must not change
    n = 40
    starts, goals = _problems(synth, oracle_world, n, seed=77)
    rb = P.RrtcBatch(gpu_world, n, -np.pi, np.pi, extension_type=ext, seed=21, query_id_base=7,
                     max_iters=1000, max_nodes=4096)
    rb.run(starts, goals)
    summ = rb.summary()
    assert (summ["iterations"] > 256).sum() >= 2          # the team path is exercised
    for qi in range(n):
        ref = oracle_world.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, ext=EXT_CODE[ext], seed=21,
                                      query=7 + qi, max_iters=1000, impl="port")
        _compare(rb, summ, qi, ref)


def test_batch_matches_reference_goldens(P, gpu_world):
    g = np.load(os.path.join(HERE, "golden", "rrtc_golden.npz"))
    for ext in EXTS:
        for rng in (0.0, 0.9):
            rb = P.RrtcBatch(gpu_world, 8, -np.pi, np.pi, rng=rng, extension_type=ext, seed=11, query_id_base=0,
                             max_iters=400, max_nodes=4096)
            rb.run(g["starts"], g["goals"])
            summ = rb.summary()
            for qi in range(8):
                key = "q%d_%s_r%s" % (qi, ext, rng)
                ref = dict(status=int(g[key + "_meta"][0]), iterations=int(g[key + "_meta"][1]),
                           start_parent=g[key + "_sp"], goal_parent=g[key + "_gp"], start_states=g[key + "_ss"],
                           goal_states=g[key + "_gs"], path=g[key + "_path"])
                _compare(rb, summ, qi, ref)


def test_batch_invalid_endpoints_and_caps(P, synth, gpu_world, oracle_world):
    q = synth.cloud(77, 400)
    v, _ = oracle_world.is_valid_batch(q)
    bad, good = q[v == 0][:2], q[v == 1][:4]
    starts = np.stack([bad[0], good[0], good[1], good[2]])
    goals = np.stack([good[3], bad[1], good[1], good[3]])
    rb = P.RrtcBatch(gpu_world, 4, -np.pi, np.pi, max_iters=50, max_nodes=64)
    rb.run(starts, goals)
    s = rb.summary()
    assert s["status"][0] == 1 and s["n_start"][0] == 0      # INVALID_START
    assert s["status"][1] == 4 and s["n_goal"][1] == 0       # invalid goal: never sampled -> TIMEOUT
    assert s["status"][2] == 6                               # start == goal
    assert s["iterations"][3] <= 50
    for qi in (2, 3):  # and the oracle agrees on these too
        ref = oracle_world.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, seed=0, query=qi, max_iters=50)
        assert s["status"][qi] == ref["status"] and s["iterations"][qi] == ref["iterations"]
    with pytest.raises(ValueError):
        P.RrtcBatch(gpu_world, 1, -1, 1, extension_type="con-bad")


def test_paths_are_collision_free(P, synth, gpu_world, oracle_world):
    """Size-independent property at a larger batch: every returned path starts at the start, ends
    at the goal, and all of its edges pass checkMotion."""
    n = 512
    starts, goals = _problems(synth, oracle_world, n, seed=60)
    rb = P.RrtcBatch(gpu_world, n, -np.pi, np.pi, extension_type="ext-con", seed=9, max_iters=500,
                     max_nodes=2048)
    rb.run(starts, goals)
    s = rb.summary()
    solved = np.where(s["status"] == 6)[0]
    assert len(solved) > 0.8 * n
    A, B = [], []
    for qi in solved[:64]:
        got = rb.fetch(int(qi), int(s["n_start"][qi]), int(s["n_goal"][qi]))
        path = got["path"]
        assert np.array_equal(path[0], starts[qi]) and np.array_equal(path[-1], goals[qi])
        A.append(path[:-1])
        B.append(path[1:])
    A, B = np.vstack(A), np.vstack(B)
    lvs = np.sqrt(7 * (2 * np.pi) ** 2) * 0.01
    assert np.all(gpu_world.check_motion_batch(A, B, lvs) == 1)


def test_tree_capacity_in_both_kernels(P, O, synth, gpu_world, oracle_world):
    # max_nodes is a capacity of this implementation (the reference's trees are unbounded): a query whose tree fills
    # stops with TIMEOUT and never writes past its buffers; queries that stay below the capacity are unaffected --
    # whether they finish in the warp kernel or, after sample 256, in the team kernel.
    n, cap = 64, 40
    starts, goals = _problems(synth, oracle_world, n, seed=91)
    rb = P.RrtcBatch(gpu_world, n, -np.pi, np.pi, extension_type="ext-con", seed=33, query_id_base=0,
                     max_iters=1000, max_nodes=cap)
    rb.run(starts, goals)
    summ = rb.summary()
    assert summ["n_start"].max() <= cap and summ["n_goal"].max() <= cap
    full = (summ["n_start"] == cap) | (summ["n_goal"] == cap)
    assert np.all(summ["status"][full & (summ["status"] != 6)] == 4)
    checked = 0
    for qi in np.flatnonzero(~full)[:24]:
        ref = oracle_world.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, ext=EXT_CODE["ext-con"], seed=33,
                                      query=int(qi), max_iters=1000, impl="port")
        _compare(rb, summ, int(qi), ref)
        checked += 1
    assert checked >= 8
