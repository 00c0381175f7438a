"""One batched checkMotion call at the headline shape (profiling target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, prgpu_loader
P = prgpu_loader.load(); synth = P.synth
ne = int(sys.argv[1])
dev = torch.device("cuda:0"); st = torch.cuda.current_stream().cuda_stream
dh, link, xyzr = synth.default_arm(); sph, box = synth.obstacles(3)
w = P.World(dh, link, xyzr, sph, box)
a, b = synth.edges(2, ne)
A, B = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
v = torch.empty(ne, dtype=torch.uint8, device=dev)
for _ in range(3):
    w.check_motion_batch_device(A, B, ne, 0.01, v, stream=st)
torch.cuda.synchronize()
print("ok", v.float().mean().item())
