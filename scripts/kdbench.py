import sys,time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, prgpu_loader
P=prgpu_loader.load(); synth=P.synth
dev=torch.device('cuda',0); stream=torch.cuda.current_stream().cuda_stream
for N in (10_000_000, 1_250_000):
    pts=torch.from_numpy(synth.cloud(0,N)).to(dev)
    tree=P.KnnIndex(7,capacity=N); tree.add_device(pts,N,stream); torch.cuda.synchronize(); tree.build_index()
    for Q,k in ((4096,16),(4096,1),(512,16),(64,16),(32768,16)):
        q=torch.from_numpy(synth.cloud(1,Q)).to(dev)
        i_=torch.empty((Q,k),dtype=torch.int32,device=dev); d_=torch.empty((Q,k),dtype=torch.float64,device=dev)
        res={}
        for mode in (2,0):
            tree.set_search_mode(mode)
            for _ in range(3): tree.nearest_k_device(q,Q,k,i_,d_,stream=stream)
            b,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            b.record()
            for _ in range(10): tree.nearest_k_device(q,Q,k,i_,d_,stream=stream)
            e.record(); torch.cuda.synchronize()
            res[mode]=(b.elapsed_time(e)/10, i_.clone(), d_.clone())
        same=torch.equal(res[0][1],res[2][1]) and torch.equal(res[0][2],res[2][2])
        print(f"N={N} Q={Q} k={k}: brute {res[2][0]:.3f} ms, indexed {res[0][0]:.3f} ms ({Q/res[0][0]*1e3:.3g} q/s) same={same}", flush=True)
    st=tree.index_stats(); print(st, "leaves/query", st['leaves_visited']/ max(1,(st['index_calls'])), flush=True)
    tree.close(); del pts
