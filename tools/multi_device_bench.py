"""One process, several GPUs through the C ABI (rayuela_init): encode 1M x 128 (m=8, ilsiter=32) from pinned host arrays
and search a base-sharded 1M-code index, for 1 .. all visible devices.  Codes / ids must not depend on the device count."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench, rayuela_b200 as rb
from rayuela_b200 import core
ndev = torch.cuda.device_count()
dev = torch.device('cuda', 0)
n, m, d, nq = 1_000_000, 8, 128, 10_000
X, Q = bench.make_data(n, nq, d, 1000, dev)
C = bench.train_codebooks(X[:50000], m, dev)
B0 = torch.randint(0, 256, (n, m), device=dev, dtype=torch.uint8)
Xh, Ch, B0h, Qh = bench.pinned(X.cpu()), bench.pinned(C.cpu()), bench.pinned(B0.cpu()), bench.pinned(Q.cpu())
out, ref_codes, ref_ids = {}, None, None
for D in [k for k in (1, 2, 4, 8) if k <= ndev]:
    rb.init(list(range(D)) if D > 1 else None)
    core.encode_icm(Xh, Ch, B0h, 32, 4, 4, True, seed=2024)                      # warm-up (pools, streams)
    t0 = time.perf_counter()
    for _ in range(3): r = core.encode_icm(Xh, Ch, B0h, 32, 4, 4, True, seed=2024)
    enc = (time.perf_counter() - t0) / 3
    Bh = r["B"]
    rec = sum(Ch.reshape(m, 256, d)[j][Bh[:, j]] for j in range(m)); nrm = (rec * rec).sum(1).astype(np.float32)
    ix = core.Index(core.SCAN_LSQ, Bh, nrm)
    res = {}
    for k in (1, 1000):
        ix.search(Qh, Ch, k)
        t0 = time.perf_counter()
        for _ in range(3): dd, ii = ix.search(Qh, Ch, k)
        res[k] = ((time.perf_counter() - t0) / 3, ii)
    ix.free()
    if ref_codes is None: ref_codes, ref_ids = Bh, {k: v[1] for k, v in res.items()}
    out["devices_%d" % D] = {"encode_ms_host_to_host": enc * 1e3, "vectors_per_sec": n / enc,
                            "codes_equal_single_device": bool(np.array_equal(Bh, ref_codes)),
                            **{"scan_k%d_ms_host_to_host" % k: res[k][0] * 1e3 for k in res},
                            **{"scan_k%d_ids_equal_single_device" % k: bool(np.array_equal(res[k][1], ref_ids[k])) for k in res}}
rb.init(None)
print(json.dumps(out, indent=1))
