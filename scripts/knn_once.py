"""One batched kNN call at the headline shape (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
n, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
pts = torch.from_numpy(synth.cloud(0, n)).to(dev)
t = P.KnnIndex(7, capacity=n); t.add_device(pts, n, st)
if len(sys.argv) > 4 and sys.argv[4] == "exact": t.set_search_mode(True)
q = torch.from_numpy(synth.cloud(1, nq)).to(dev)
idx = torch.empty((nq, k), dtype=torch.int32, device=dev); dist = torch.empty((nq, k), dtype=torch.float64, device=dev)
for _ in range(3):
    t.nearest_k_device(q, nq, k, idx, dist, stream=st)
torch.cuda.synchronize()
print("ok", t.search_stats())
