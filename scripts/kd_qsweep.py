"""Indexed search time over the batch size (team kernel: queries per block follow the batch size)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
dev = torch.device('cuda', 0); stream = torch.cuda.current_stream().cuda_stream
N, k = 10_000_000, 16
pts = torch.from_numpy(synth.cloud(0, N)).to(dev)
tree = P.KnnIndex(7, capacity=N); tree.add_device(pts, N, stream); torch.cuda.synchronize(); tree.build_index()
for Q in (64, 256, 512, 1024, 1536, 2048, 3072, 4096, 8192):
    q = torch.from_numpy(synth.cloud(1, Q)).to(dev)
    i_ = torch.empty((Q, k), dtype=torch.int32, device=dev); d_ = torch.empty((Q, k), dtype=torch.float64, device=dev)
    for _ in range(3):
        tree.nearest_k_device(q, Q, k, i_, d_, stream=stream)
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(10):
        tree.nearest_k_device(q, Q, k, i_, d_, stream=stream)
    e.record(); torch.cuda.synchronize()
    print(f"Q={Q}: {b.elapsed_time(e) / 10:.3f} ms", flush=True)
