"""Generates tests/golden/rrtc_golden.npz by running the REFERENCE's own planner
(pr_ompl/src/RRTConnect.cpp compiled unmodified into oracle/_ref/libref_rrtconnect.so, see
oracle/Makefile) on seeded synthetic problems.  Run in the build container, where
/root/reference is mounted:   python tests/golden/make_golden.py
The committed .npz is what travels to the GPU box (which has no /root/reference)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle_lib as O  # noqa: E402
import prgpu_loader  # noqa: E402

EXTS = {"ext-ext": (0, 0), "ext-con": (0, 1), "con-ext": (1, 0), "con-con": (1, 1)}


def problems(synth, W, n_pairs, seed=4):
    """Valid (start, goal) pairs: BASELINE 'independent RRTConnect queries', seed 4."""
    q = synth.cloud(seed, 8 * n_pairs)
    v, _ = W.is_valid_batch(q, 8)
    qv = q[v == 1]
    return qv[0:2 * n_pairs:2], qv[1:2 * n_pairs:2]


def main():
    assert O.ref() is not None, "oracle/_ref not built (needs /root/reference)"
    synth = prgpu_loader.load().synth
    dh, link, xyzr = synth.default_arm()
    sph, box = synth.obstacles(3)
    W = O.World(dh, link, xyzr, sph, box)
    starts, goals = problems(synth, W, 8)
    out = {"starts": starts, "goals": goals}
    for qi in range(8):
        for name, ext in EXTS.items():
            for rng in (0.0, 0.9):
                r = W.rrtc_solve(starts[qi], goals[qi], -np.pi, np.pi, rng=rng, ext=ext, seed=11, query=qi,
                                 max_iters=400, impl="ref")
                key = "q%d_%s_r%s" % (qi, name, rng)
                out[key + "_meta"] = np.array([r["status"], r["iterations"]], dtype=np.int64)
                out[key + "_sp"] = r["start_parent"]
                out[key + "_gp"] = r["goal_parent"]
                out[key + "_ss"] = r["start_states"]
                out[key + "_gs"] = r["goal_states"]
                out[key + "_path"] = r["path"]
    np.savez_compressed(os.path.join(HERE, "rrtc_golden.npz"), **out)
    print("wrote rrtc_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
