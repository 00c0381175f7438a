"""Per-call latency of the one-at-a-time calls a host planner makes (Drop-in A): nearest(1), add(1), checkMotion(1)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
dh, link, xyzr = synth.default_arm(); sph, box = synth.obstacles(3)
gw = P.World(dh, link, xyzr, sph, box)
rng = np.random.default_rng(0)
pts = rng.uniform(-np.pi, np.pi, size=(4000, 7)); q = rng.uniform(-np.pi, np.pi, size=(1000, 7))
a, b = synth.edges(5, 1000, max_len=0.3)
tree = P.KnnIndex(7); tree.add(pts[:200])
def t(f, n=1000):
    f(0); t0 = time.perf_counter()
    for i in range(n): f(i)
    return (time.perf_counter() - t0) / n * 1e6
print("nearest(1) on 200 nodes   %.1f us" % t(lambda i: tree.nearest_k(q[i:i+1], 1)))
print("checkMotion(1) 0.3 rad    %.1f us" % t(lambda i: gw.check_motion_batch(a[i:i+1], b[i:i+1], 0.01)))
print("add(1)                    %.1f us" % t(lambda i: tree.add(pts[200+i:201+i])))
print("nearest(1) on 1200 nodes  %.1f us" % t(lambda i: tree.nearest_k(q[i:i+1], 1)))
print("add(1)+nearest(1)         %.1f us" % t(lambda i: (tree.add(pts[1300+i:1301+i]), tree.nearest_k(q[i:i+1], 1))))
print("is_valid(1)               %.1f us" % t(lambda i: gw.is_valid_batch(q[i:i+1])))
