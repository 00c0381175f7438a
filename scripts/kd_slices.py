"""Indexed search time of slices of the 4096-query batch (are there pathological queries?)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
dev = torch.device('cuda', 0); stream = torch.cuda.current_stream().cuda_stream
N, k = 10_000_000, 16
pts = torch.from_numpy(synth.cloud(0, N)).to(dev)
tree = P.KnnIndex(7, capacity=N); tree.add_device(pts, N, stream); torch.cuda.synchronize(); tree.build_index()
qall = torch.from_numpy(synth.cloud(1, 4096)).to(dev)
def t(q):
    Q = len(q)
    i_ = torch.empty((Q, k), dtype=torch.int32, device=dev); d_ = torch.empty((Q, k), dtype=torch.float64, device=dev)
    for _ in range(3):
        tree.nearest_k_device(q, Q, k, i_, d_, stream=stream)
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(10):
        tree.nearest_k_device(q, Q, k, i_, d_, stream=stream)
    e.record(); torch.cuda.synchronize()
    return b.elapsed_time(e) / 10
for size in (1024, 256):
    print(size, [round(t(qall[o:o + size].contiguous()), 3) for o in range(0, 4096, size)], flush=True)
st0 = tree.index_stats()["leaves_visited"]
