"""A/B timing of this session's kernel variants on one B200 (env knobs are read per call)."""
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

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "icm"):
    n, ils = 1000000, 32
    X, Q = bench.make_data(n, 100, 128, 1000, dev)
    for m in (8, 7):
        C = bench.train_codebooks(X[:50000], m, dev)
        B0 = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
        outs = {}
        for usm in ("0", "1"):
            os.environ["RAYUELA_B200_ICM_USM"] = usm
            B = B0.clone()
            def step():
                B.copy_(B0); core.encode_icm(X, C, B, ils, 4, 4, True, seed=2024, inplace=True)
            ms = timeit(step, 3)
            outs[usm] = B.clone()
            print(f"icm m={m} USM={usm}: {ms:.2f} ms  {n/ms*1e3:,.0f} vectors/s", flush=True)
        print("  codes equal across variants:", bool(torch.equal(outs["0"], outs["1"])), flush=True)
    del X, Q
if what in ("all", "scan"):
    n, nq, d = 1000000, 10000, 128
    g = torch.Generator(device=dev).manual_seed(0)
    for m, envs in ((8, ({"RAYUELA_B200_SCANX8": "0"}, {"RAYUELA_B200_SCANX8": "1"})),
                    (16, ({"RAYUELA_B200_SCAN_V1": "1"}, {"RAYUELA_B200_SCAN_V1": "0"})),
                    (15, ({"RAYUELA_B200_SCAN_V1": "0"},))):
        B = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8, generator=g)
        nrm = torch.randn(n, device=dev, generator=g) * 3
        Q = torch.randn(nq, d, device=dev, generator=g)
        C = torch.randn(m * 256, d, device=dev, generator=g)
        for k in (1, 100, 1000):
            res = []
            for env in envs:
                os.environ.update(env)
                ix = core.Index(core.SCAN_LSQ, B, nrm)
                ms = timeit(lambda: ix.search(Q, C, k), 5)
                res.append(ix.search(Q, C, k))
                ix.free()
                print(f"scan lsq m={m} k={k} {env}: {ms:.3f} ms  {nq/ms*1e3:,.0f} q/s  lookups/clk/SM@1.9GHz {nq*n*m/(ms*1e-3)/148/1.9e9:.2f}", flush=True)
            if len(res) == 2:
                same = bool((res[0][0].view(torch.int32) == res[1][0].view(torch.int32)).all() and (res[0][1] == res[1][1]).all())
                print("  results identical across variants:", same, flush=True)
