"""One LSQ search per mode after warm-up (for ncu): python tools/scan_one.py n nq m k reps"""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
dev = torch.device('cuda')
n, nq, m, k = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
g = torch.Generator(device=dev).manual_seed(0)
B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
nrm = torch.randn(n, device=dev, generator=g) * 3
Q = torch.randn(nq, 128, device=dev, generator=g)
C = torch.randn(m * 256, 128, device=dev, generator=g)
ix = core.Index(core.SCAN_LSQ, B, nrm)
for _ in range(reps): ix.search(Q, C, k)
torch.cuda.synchronize()
