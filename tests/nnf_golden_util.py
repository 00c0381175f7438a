"""Helpers shared by the NNF golden-vector tests (CPU port and GPU).  The goldens come from the reference's
own frechet/ sources compiled unmodified (tests/golden/make_golden_nnf.py)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nnf_golden.npz")
DBL_MAX = np.finfo(np.float64).max


def load():
    z = np.load(GOLDEN)
    cases = {}
    for name in z["cases"]:
        name = str(name)
        cases[name] = {key.split("/", 1)[1]: z[key] for key in z.files if key.startswith(name + "/")}
    return cases


def case_world_arrays(synth, c):
    ns, nb = int(c["params"][6]), int(c["params"][7])
    if ns + nb:
        return synth.obstacles(3, n_spheres=ns, n_boxes=nb)
    return np.zeros((0, 4)), np.zeros((0, 6))


def coupling_cost(ref_sub, pose, path):
    """Min over monotone couplings of the reference path with [c0] + path + [c1] of the largest task
    distance (= the bottleneck cost of the best tensor-product path that projects onto `path`)."""
    seq = [0] + [int(v) for v in path] + [1]
    f = np.sqrt(((ref_sub[:, None, [3, 7, 11]] - pose[seq][None, :, [3, 7, 11]]) ** 2).sum(-1))
    R, L = f.shape
    B = np.full((R, L), np.inf)
    B[0, 0] = f[0, 0]
    for r in range(R):
        for j in range(L):
            if r == 0 and j == 0:
                continue
            best = np.inf
            if r > 0:
                best = min(best, B[r - 1, j])
            if j > 0:
                best = min(best, B[r, j - 1])
            if r > 0 and j > 0:
                best = min(best, B[r - 1, j - 1])
            B[r, j] = max(best, f[r, j])
    return B[R - 1, L - 1]


def check_solution_path(c, path, oracle_world):
    """`path` (NN-graph vertex ids) must be a real path c0 -> ... -> c1 of the golden NN graph, free of
    collisions under the specification's evaluateEdge, and optimal: its best coupling costs exactly the
    reference's final goal cost."""
    edges = set(map(tuple, c["edges"].tolist()))
    path = [int(v) for v in path]
    assert (0, path[0]) in edges and (path[-1], 1) in edges
    for u, v in zip(path[:-1], path[1:]):
        assert (u, v) in edges, (u, v)
    st = c["states"]
    if len(path) > 1:
        ok, clear, _ = oracle_world.check_motion_batch(st[path[:-1]], st[path[1:]], 0.05, mode=1)
        assert np.all((ok == 1) | (np.abs(clear) < 1e-6))
    assert coupling_cost(c["ref_sub"], c["pose"], path) == float(c["goal_cost"])
