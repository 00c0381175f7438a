"""Tree sizes of solved / unsolved queries of the batched planner benchmark (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, prgpu_loader, bench
P = prgpu_loader.load(); synth = P.synth
dh, link, xyzr = synth.default_arm(); sph, box = synth.obstacles(3)
gw = P.World(dh, link, xyzr, sph, box)
B = 4096
ps, pg = bench.plan_problems(synth, gw, B)
for mi in (1000, 200, 50):
    rb = P.RrtcBatch(gw, B, -np.pi, np.pi, extension_type="ext-con", seed=4, max_iters=mi, max_nodes=2048)
    rb.run(ps, pg)
    t = time.perf_counter(); rb.run(ps, pg); dt = time.perf_counter() - t
    s = rb.summary()
    un = s["status"] != 6
    print(f"max_iters={mi}: {dt*1e3:.2f} ms; unsolved {un.sum()}; solved: iters mean {s['iterations'][~un].mean():.1f} max {s['iterations'][~un].max()} "
          f"nS {s['n_start'][~un].mean():.1f} nG {s['n_goal'][~un].mean():.1f}; unsolved: nS {s['n_start'][un].mean():.1f} nG {s['n_goal'][un].mean():.1f} "
          f"status {np.unique(s['status'][un], return_counts=True)}")
