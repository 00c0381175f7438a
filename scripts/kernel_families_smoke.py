"""The smallest case of every kernel family, for a quick GPU smoke of every kernel (compute-sanitizer is closed on this GPU pool: the call is refused).  No torch: host
buffers through the C ABI only.  Usage: sanitize_cases.py [family ...]   (default: all)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import prgpu_loader

P = prgpu_loader.load()
synth = P.synth
dh, link, xyzr = synth.default_arm()


def knn_exact():      # knn_tile_kernel, topk_merge_kernel, transpose_append_kernel
    t = P.KnnIndex(7)
    t.add(synth.cloud(0, 3000, dup_fraction=0.05))
    for k in (1, 16):
        t.nearest_k(synth.cloud(1, 70), k)
    t.nearest_k(synth.cloud(1, 1), 1)          # the single-query (broadcast) form
    t.close()


def knn_filter():     # knn_sample / collect (mbarrier ring) / threshold / select, fallback gather + scatter
    t = P.KnnIndex(7)
    t.add(synth.cloud(0, 8192 + 77))
    t.set_search_mode(2)                       # brute force: the tensor-core filter path
    t.nearest_k(synth.cloud(1, 128), 16)
    t.close()
    t = P.KnnIndex(7)                          # every point identical: every query takes the exact fallback
    t.add(np.tile(synth.cloud(5, 1), (8192, 1)))
    t.set_search_mode(2)
    t.nearest_k(synth.cloud(1, 128), 3)
    t.close()
    t = P.KnnIndex(7)
    t.add(synth.cloud(0, 8192))
    t.diag_scan(synth.cloud(1, 64), 1e9, cap=8192)
    t.close()


def knn_kd():         # kd_* build kernels + kd_search_kernel
    t = P.KnnIndex(7)
    t.add(synth.cloud(0, 20000))
    t.set_search_mode(3)
    t.nearest_k(synth.cloud(1, 256), 16)
    t.close()


def check_motion():   # check_motion_kernel (small), cm_sample/resolve/gaps/expand/phase_b (>= 4096 edges), is_valid, fk
    sph, box = synth.obstacles(3)
    w = P.World(dh, link, xyzr, sph, box)
    a, b = synth.edges(2, 4096 + 100)
    w.check_motion_batch(a[:200], b[:200], 0.01)
    w.check_motion_batch(a, b, 0.01)
    w.check_motion_batch(a, b, 0.05, mode=P.MODE_NNF_ALPHA)
    w.is_valid_batch(a[:500])
    w.diag_fk_centre_error(a[:500])
    w.close()


def rrtc():           # rrtc_batch_kernel + rrtc_team_kernel (queries that exhaust 256 samples are handed to teams)
    import bench
    sph, box = synth.obstacles(3)
    w = P.World(dh, link, xyzr, sph, box)
    n = 96
    ps, pg = bench.plan_problems(synth, w, n)
    for ext in ("ext-con", "con-con"):
        rb = P.RrtcBatch(w, n, -np.pi, np.pi, extension_type=ext, seed=4, max_iters=400, max_nodes=1024)
        rb.run(ps, pg)
        s = rb.summary()
        rb.close()
    print("rrtc solved", float((s["status"] == 6).mean()))
    w.close()


def nnf():            # nnf_subsample / translations / rowmin / band / band_pack / flow_dp / band_dp / dp / backtracks
    import bench
    sph, box = synth.obstacles(3, n_spheres=16, n_boxes=16)
    w = P.World(dh, link, xyzr, sph, box)
    _, ref, counts, ik = bench.nnf_problem(synth, dh, 12, 24, 1)
    for mode in (0, 1, 2):
        g = P.Nnf(w, ref, counts, ik, num_nn=4, discretization=1)
        g.set_mode(mode)
        r = g.solve(max_lazy_iters=40)
        g.close()
    print("nnf iters", r["lazy_iters"], "solved", r["solved"])
    w.close()


FAMILIES = dict(knn_exact=knn_exact, knn_filter=knn_filter, knn_kd=knn_kd, check_motion=check_motion, rrtc=rrtc, nnf=nnf)
for name in (sys.argv[1:] or list(FAMILIES)):
    FAMILIES[name]()
    print("done", name, flush=True)
