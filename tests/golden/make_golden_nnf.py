"""Generates tests/golden/nnf_golden.npz by running the REFERENCE's own NNF planner
(frechet/src/{NNFrechet,LPAStar,MinPriorityQueue,util}.cpp compiled unmodified into
oracle/_ref/libref_frechet.so over the stand-ins of oracle/compat/, see oracle/Makefile) on seeded
synthetic problems.  Run in the build container, where /root/reference is mounted:
    python tests/golden/make_golden_nnf.py
The committed .npz is what travels to the GPU box (which has no /root/reference).

Per case the file holds the inputs (full reference path, pre-drawn IK solutions) and, from the reference:
  R, ref_ids-equivalent sub-sampled poses, the IK budget per waypoint (std::default_random_engine draws,
  NNFrechet.cpp:310-343), NN-graph sizes, TPG sizes, findKnnNodes choices, every vertex state and pose,
  the NN edge list, the first search's goal cost and mFinalError (search_only run), and after solve():
  solved flag, final g[goal], mFinalError, evaluated / knocked-out edge counts, the solution path.
Everything up to and including the costs is independent of the stand-in graph's edge iteration order
(checked here against the reversed-order build, libref_frechet_rev.so); evaluated / knocked-out counts
and the path itself depend on which of several equally good paths LPA* keeps, and are stored for
information under *_order_dependent keys."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
import prgpu_loader  # noqa: E402

# (name, trajectory seed, planner seed, W, M, k, D, obstacle spheres, obstacle boxes)
CASES = [
    ("defaults_s4", 4, 1, 5, 5, 5, 3, 32, 32),       # reference defaults, NNFrechet.hpp:262-274
    ("defaults_s1", 1, 1, 5, 5, 5, 3, 32, 32),
    ("defaults_free", 5, 1, 5, 5, 5, 3, 0, 0),       # no obstacles: one search
    ("w8m6_unsolvable", 2, 1, 8, 6, 3, 1, 32, 32),   # every path in collision: goal cost = DBL_MAX
    ("w12m20", 3, 1, 12, 20, 5, 1, 32, 32),
    ("w12m20_seed7", 1, 7, 12, 20, 5, 1, 32, 32),    # planner seed 7: a different IK budget
    ("w10m16_free", 5, 1, 10, 16, 5, 3, 0, 0),
    ("w20m30", 4, 1, 20, 30, 4, 2, 32, 32),
    ("w30m40", 4, 1, 30, 40, 5, 1, 32, 32),
]
ORDER_FREE = ["R", "n_vertices", "n_edges", "n_ik", "n_tpg_vertices", "n_tpg_edges", "solved", "goal_cost"]


def make_case(synth, dh, link, xyzr, case):
    name, tseed, pseed, W, M, k, D, ns, nb = case
    if ns + nb:
        sph, box = synth.obstacles(3, n_spheres=ns, n_boxes=nb)
    else:
        sph, box = np.zeros((0, 4)), np.zeros((0, 6))
    world = O.World(dh, link, xyzr, sph, box)
    R = (W - 1) * (D + 1) + 1
    q, poses = synth.nnf_reference(tseed, 2 * R + 5, dh)
    ids = synth.lin_int_space(0, len(q) - 1, R)
    counts = O.nnf_layer_counts(pseed, R, W, M)   # only to size/generate the IK pool; the reference draws its own
    ik = synth.nnf_ik_states(2, q[ids], counts)
    first = O.ref_nnf_run(world, poses, W, M, k, D, ik, seed=pseed, search_only=True)
    full = O.ref_nnf_run(world, poses, W, M, k, D, ik, seed=pseed)
    first_r = O.ref_nnf_run(world, poses, W, M, k, D, ik, seed=pseed, search_only=True, reverse_order=True)
    full_r = O.ref_nnf_run(world, poses, W, M, k, D, ik, seed=pseed, reverse_order=True)
    for a, b in ((first, first_r), (full, full_r)):
        for key in ORDER_FREE:
            assert a[key] == b[key], (name, key, a[key], b[key])
        for key in ("ref", "ik_call_counts", "knn", "states", "pose", "edges"):
            assert np.array_equal(a[key], b[key]), (name, key)
    out = {
        "params": np.array([tseed, pseed, W, M, k, D, ns, nb], dtype=np.int64),
        "ref_full": poses, "ik": ik,
        "ref_sub": full["ref"], "ik_call_counts": full["ik_call_counts"],
        "sizes": np.array([full["R"], full["n_vertices"], full["n_edges"], full["n_ik"], full["n_tpg_vertices"],
                           full["n_tpg_edges"]], dtype=np.int64),
        "knn": full["knn"], "states": full["states"], "pose": full["pose"], "edges": full["edges"],
        "first_goal_cost": np.float64(first["goal_cost"]),
        "first_final_error": np.array([first["final_error"], first_r["final_error"]]),
        "solved": np.int64(full["solved"]), "goal_cost": np.float64(full["goal_cost"]),
        "final_error": np.array([full["final_error"], full_r["final_error"]]),
        "counts_order_dependent": np.array([[full["n_evaluated"], full["n_collisions"]],
                                            [full_r["n_evaluated"], full_r["n_collisions"]]], dtype=np.int64),
        "path_states_order_dependent": full["path_states"],
        "edge_flags_order_dependent": full["edge_flags"],
    }
    print("%-16s R=%d V=%d E=%d TPG=%d/%d solved=%d first=%.6g final=%.6g err=%.6g/%.6g evaluated=%d/%d (%.1fs+%.1fs)"
          % (name, full["R"], full["n_vertices"], full["n_edges"], full["n_tpg_vertices"], full["n_tpg_edges"],
             full["solved"], first["goal_cost"], full["goal_cost"], full["final_error"], full_r["final_error"],
             full["n_evaluated"], full_r["n_evaluated"], full["setup_seconds"], full["solve_seconds"]))
    return {name + "/" + key: v for key, v in out.items()}


def main():
    assert O.ref_nnf() is not None and O.ref_nnf(True) is not None, "oracle/_ref not built (needs /root/reference)"
    synth = prgpu_loader.load().synth
    dh, link, xyzr = synth.default_arm()
    out = {"cases": np.array([c[0] for c in CASES])}
    for case in CASES:
        out.update(make_case(synth, dh, link, xyzr, case))
    np.savez_compressed(os.path.join(HERE, "nnf_golden.npz"), **out)
    print("wrote nnf_golden.npz with", len(out), "arrays,",
          os.path.getsize(os.path.join(HERE, "nnf_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
