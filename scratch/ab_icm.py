"""A/B timing of the ICM kernel variants (env knobs are read per call)."""
import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
dev = torch.device('cuda')
def timeit(fn, reps):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms_ = [int(x) for x in (sys.argv[1].split(',') if len(sys.argv) > 1 else ['8', '7', '16'])]
variants = sys.argv[2].split(',') if len(sys.argv) > 2 else ['00', '10', '01', '11']   # (USM, PF)
ils = 32
for m in ms_:
    n = 1000000 if m <= 8 else 250000
    X, Q = bench.make_data(n, 100, 128, 1000, dev)
    C = bench.train_codebooks(X[:50000], m, dev)
    B0 = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
    outs = {}
    for v in variants:
        os.environ["RAYUELA_B200_ICM_USM"], os.environ["RAYUELA_B200_ICM_PF"] = v[0], v[1]
        B = B0.clone()
        def step(stats=False):
            B.copy_(B0); return core.encode_icm(X, C, B, ils, 4, 4, True, seed=2024, inplace=True, want_stats=stats)
        ms = timeit(step, 3)
        step(True)
        ex, tot = core.last_icm_steps(); exact = core.last_icm_exact_steps()
        outs[v] = B.clone()
        print(f"icm m={m} n={n} USM={v[0]} PF={v[1]}: {ms:.2f} ms  {n/ms*1e3:,.0f} vectors/s  executed {ex/tot:.3f} of steps, exact-path {exact/max(ex,1):.4f} of executed", flush=True)
    ks = list(outs)
    print("  codes equal across variants:", all(bool(torch.equal(outs[ks[0]], outs[k])) for k in ks[1:]), flush=True)
    del X, Q
