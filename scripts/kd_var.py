"""Times the indexed search of the headline shape (N=1e7, 7-D, k=16) for a few batch sizes; used to compare
build variants of kdtree.cu (the library under test is whatever pr-ompl_b200/libprgpu.so is)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
dev = torch.device('cuda', 0); stream = torch.cuda.current_stream().cuda_stream
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
tag = sys.argv[2] if len(sys.argv) > 2 else ''
pts = torch.from_numpy(synth.cloud(0, N)).to(dev)
tree = P.KnnIndex(7, capacity=N); tree.add_device(pts, N, stream); torch.cuda.synchronize(); tree.build_index()
ref = None
for Q, k in ((4096, 16), (512, 16), (32768, 16), (4096, 1)):
    q = torch.from_numpy(synth.cloud(1, Q)).to(dev)
    i_ = torch.empty((Q, k), dtype=torch.int32, device=dev); d_ = torch.empty((Q, k), dtype=torch.float64, device=dev)
    for _ in range(3):
        tree.nearest_k_device(q, Q, k, i_, d_, stream=stream)
    ts = []
    for rep in range(3):
        b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b.record()
        for _ in range(10):
            tree.nearest_k_device(q, Q, k, i_, d_, stream=stream)
        e.record(); torch.cuda.synchronize()
        ts.append(b.elapsed_time(e) / 10)
    chk = (int(i_.to(torch.int64).sum().item()), float(d_.sum().item()))
    print(f"{tag} N={N} Q={Q} k={k}: {min(ts):.4f} ms (median {sorted(ts)[1]:.4f})  checksum {chk}", flush=True)
st = tree.index_stats(); print(tag, st, flush=True)
