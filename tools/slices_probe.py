import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
from rayuela_b200 import core
dev = torch.device('cuda')
n, d, m = 1000000, 128, 8
g = torch.Generator(device=dev).manual_seed(0)
B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
nrm = torch.randn(n, device=dev, generator=g) * 3
C = torch.randn(m * 256, d, device=dev, generator=g)
ix = core.Index(core.SCAN_LSQ, B, nrm)
for nq in (33 * 16, 63 * 16, 100 * 16, 16, 160, 10000):
    Q = torch.randn(nq, d, device=dev, generator=g)
    for k in (1, 100):
        out = []
        for S in ("auto", "1", "2", "3", "4", "6", "8", "12", "16"):
            if S == "auto": os.environ.pop("RAYUELA_B200_SCAN_SLICES", None)
            else: os.environ["RAYUELA_B200_SCAN_SLICES"] = S
            for _ in range(2): ix.search(Q, C, k)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): ix.search(Q, C, k)
            e1.record(); torch.cuda.synchronize()
            out.append("%s:%.2f" % (S, e0.elapsed_time(e1) / 5))
        print("nq=%d k=%d  " % (nq, k) + "  ".join(out), flush=True)
