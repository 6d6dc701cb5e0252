import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
dev = torch.device('cuda')
n, nq, d = 1000000, 10000, 128
g = torch.Generator(device=dev).manual_seed(0)
os.environ["RAYUELA_B200_SCANX8"] = "1"
for m in (8, 16):
    B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
    nrm = torch.randn(n, device=dev, generator=g) * 3
    Q = torch.randn(nq, d, device=dev, generator=g)
    C = torch.randn(m * 256, d, device=dev, generator=g)
    ix = core.Index(core.SCAN_LSQ, B, nrm)
    for k in (1, 10, 100, 1000, 4000):
        out = []
        for piggy in (100, 75, 50, 25):
            soft = piggy
            os.environ["RAYUELA_B200_SCAN_PIGGY"] = str(piggy)
            for _ in range(2): ix.search(Q, C, k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): ix.search(Q, C, k)
            e1.record(); torch.cuda.synchronize()
            out.append("piggy=%d: %.2f ms" % (soft, e0.elapsed_time(e1) / 4))
        print("m=%d k=%d  " % (m, k) + "  ".join(out), flush=True)
