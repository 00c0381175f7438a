"""One batched RRTConnect run at the benchmark shape (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, prgpu_loader, bench
P = prgpu_loader.load(); synth = P.synth
dh, link, xyzr = synth.default_arm(); sph, box = synth.obstacles(3)
gw = P.World(dh, link, xyzr, sph, box)
B = int(sys.argv[1])
ps, pg = bench.plan_problems(synth, gw, B)
rb = P.RrtcBatch(gw, B, -np.pi, np.pi, extension_type="ext-con", seed=4, max_iters=1000, max_nodes=2048)
for _ in range(2):
    rb.run(ps, pg)
s = rb.summary()
print("ok solved", float((s["status"] == 6).mean()), {k: (v.mean() if hasattr(v, "mean") else v) for k, v in s.items()})
