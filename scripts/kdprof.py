"""One indexed search at the headline shape (for ncu): N=1e7, Q=4096, k=16."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
dev = torch.device('cuda', 0); stream = torch.cuda.current_stream().cuda_stream
N, Q, k = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 4096, 16
pts = torch.from_numpy(synth.cloud(0, N)).to(dev)
tree = P.KnnIndex(7, capacity=N); tree.add_device(pts, N, stream); torch.cuda.synchronize(); tree.build_index()
q = torch.from_numpy(synth.cloud(1, Q)).to(dev)
i_ = torch.empty((Q, k), dtype=torch.int32, device=dev); d_ = torch.empty((Q, k), dtype=torch.float64, device=dev)
for _ in range(4):
    tree.nearest_k_device(q, Q, k, i_, d_, stream=stream)
torch.cuda.synchronize()
print(tree.index_stats())
