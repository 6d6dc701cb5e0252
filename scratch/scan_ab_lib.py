"""A/B of two builds of the library in one process-pair: scan timings for the current and the previous linscan.cu."""
import os, subprocess, sys
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
    from rayuela_b200 import _lib
    if sys.argv[2] == "prev":
        _lib.LIB_PATH = "/root/repo/scratch/prev_lib/librayuela_b200.so"
    from rayuela_b200 import core
    dev = torch.device('cuda')
    n, nq, d, m = 1000000, 10000, 128, 8
    g = torch.Generator(device=dev).manual_seed(0)
    B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
    nrm = torch.randn(n, device=dev, generator=g) * 3
    Q = torch.randn(nq, d, device=dev, generator=g)
    C = torch.randn(m * 256, d, device=dev, generator=g)
    ix = core.Index(core.SCAN_LSQ, B, nrm)
    for nqq in (10000, 9472, 528):
      for k in (1, 100, 1000):
        Qs = Q[:nqq]
        for _ in range(3): ix.search(Qs, C, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8): ix.search(Qs, C, k)
        e1.record(); torch.cuda.synchronize()
        print(f"{sys.argv[2]} nq={nqq} k={k}: {e0.elapsed_time(e1)/8:.3f} ms", flush=True)
else:
    for rep in range(1):
        for which in ("prev", "cur"):
            subprocess.run([sys.executable, __file__, "child", which])
