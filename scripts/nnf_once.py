"""A short NNF solve at the benchmark shape (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, prgpu_loader, bench
P = prgpu_loader.load(); synth = P.synth
dh, link, xyzr = synth.default_arm(); sph, box = synth.obstacles(3, n_spheres=16, n_boxes=16)
gw = P.World(dh, link, xyzr, sph, box)
_, ref, counts, ik = bench.nnf_problem(synth, dh, 200, 1024, 1)
g = P.Nnf(gw, ref, counts, ik, num_nn=5, discretization=1)
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
r = g.solve(max_lazy_iters=int(sys.argv[1]))
t1 = time.perf_counter()
print("ok", r["lazy_iters"], "solve_s %.4f" % (t1 - t0), "goal %.9g err %.9g" % (r.get("goal_cost", 0), r.get("final_error", 0)), g.band_info())
