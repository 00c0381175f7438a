"""Ad-hoc timing of the K1/K2 kernels with inputs resident on the device (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import prgpu_loader
P = prgpu_loader.load()
synth = P.synth
dev = torch.device("cuda:0")
st = torch.cuda.current_stream().cuda_stream

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for n in [int(x) for x in sys.argv[1].split(",")]:
    pts = torch.from_numpy(synth.cloud(0, n)).to(dev)
    t = P.KnnIndex(7, capacity=n)
    t.add_device(pts, n, st)
    for nq, k in [(4096, 1), (4096, 16), (1, 1), (32, 1)]:
        q = torch.from_numpy(synth.cloud(1, nq)).to(dev)
        idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
        dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
        ms = timeit(lambda: t.nearest_k_device(q, nq, k, idx, dist, stream=st))
        pairs = n * nq
        print(f"knn n={n} nq={nq} k={k}: {ms:.3f} ms  {nq/ms*1e3:.3e} q/s  {pairs/ms*1e3:.3e} pairs/s  "
              f"tree {n*56/ms*1e3/1e9:.1f} GB/s", flush=True)
    t.close()

dh, link, xyzr = synth.default_arm(); sph, box = synth.obstacles(3)
w = P.World(dh, link, xyzr, sph, box)
ne = int(sys.argv[2])
a, b = synth.edges(2, ne)
A, B = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
v = torch.empty(ne, dtype=torch.uint8, device=dev)
ms = timeit(lambda: w.check_motion_batch_device(A, B, ne, 0.01, v, stream=st))
print(f"check_motion n={ne}: {ms:.3f} ms  {ne/ms*1e3:.3e} edges/s valid={v.float().mean().item():.3f}")
