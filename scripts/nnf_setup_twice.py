"""NNF set-up twice in one process (the second is warm): phase trace with PRGPU_NNF_TRACE=1."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, prgpu_loader, bench
P = prgpu_loader.load(); synth = P.synth
dh, link, xyzr = synth.default_arm(); sph, box = synth.obstacles(3, n_spheres=16, n_boxes=16)
gw = P.World(dh, link, xyzr, sph, box)
_, ref, counts, ik = bench.nnf_problem(synth, dh, 200, 1024, 1)
for i in range(3):
    sys.stderr.write("--- create %d\n" % i)
    t0 = time.perf_counter()
    g = P.Nnf(gw, ref, counts, ik, num_nn=5, discretization=1)
    sys.stderr.write("create %.1f ms\n" % ((time.perf_counter() - t0) * 1e3))
    if len(sys.argv) > 1:
        g.solve(max_lazy_iters=40)   # allocates the band tables, as the bench's warm-up instance does
    t0 = time.perf_counter()
    g.close()
    sys.stderr.write("close %.1f ms\n" % ((time.perf_counter() - t0) * 1e3))
