"""Times one batched kNN shape (development aid): knn_time.py N Q K [reps]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
n, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
pts = torch.from_numpy(synth.cloud(0, n)).to(dev)
t = P.KnnIndex(7, capacity=n); t.add_device(pts, n, st)
if os.environ.get('KNN_EXACT') == '1': t.set_search_mode(True)
q = torch.from_numpy(synth.cloud(1, nq)).to(dev)
idx = torch.empty((nq, k), dtype=torch.int32, device=dev); dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
for _ in range(3):
    t.nearest_k_device(q, nq, k, idx, dist, stream=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(reps):
    t.nearest_k_device(q, nq, k, idx, dist, stream=st)
e1.record(); torch.cuda.synchronize()
print(f"knn n={n} nq={nq} k={k} env={os.environ.get('PRGPU_COLLECT_WAVES')}: {e0.elapsed_time(e1)/reps:.3f} ms", t.search_stats(), flush=True)
