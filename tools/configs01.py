"""BASELINE.json configs[0] (PQ 10k plumbing, checked against the oracle / reference C++ end to end) and configs[1]
(OPQ m=8, 1M x 128: quantize_opq + linscan_opq, 10k queries, k = 1 and 1000) on one GPU.  configs[2..4] are bench.py's
headline and its config4 / config5 objects."""
import sys, json, time, torch, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/rayuela.jl_b200')
import bench
from rayuela_b200 import core
from oracle import oracle as orc
dev = torch.device('cuda')
out = {}
def ev(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r
def pq_codebooks(X, m, gen):
    sub = X.shape[1] // m
    return torch.cat([bench.kmeans(X[:50000, i*sub:(i+1)*sub].contiguous(), 256, 8, gen)[0] for i in range(m)]).contiguous()
gen = torch.Generator(device=dev).manual_seed(3)
# config 1: PQ m=8, 10k x 128, 100 queries (plumbing; checked against the oracle end to end)
X, Q = bench.make_data(10000, 100, 128, 1, dev)
Cpq = pq_codebooks(X, 8, gen)
B = core.quantize_pq(X, Cpq, 8)
ix = core.Index(core.SCAN_PQ, B); dd, ii = ix.search(Q, Cpq, 100)
B0 = orc.quantize_pq(X.cpu().numpy(), Cpq.cpu().numpy(), 8)
d0, i0 = orc.ref_linscan(orc.PQ, B0, Q.cpu().numpy(), Cpq.cpu().numpy(), 100)
out['config1_pq_10k'] = dict(codes_equal_oracle=bool(np.array_equal(B.cpu().numpy(), B0)), ids_equal_reference=bool(np.array_equal(ii.cpu().numpy(), i0)),
                             dists_bits_equal=bool(np.array_equal(dd.cpu().numpy().view(np.uint32), d0.view(np.uint32))))
# config 2: OPQ m=8 1M x 128: quantize_opq + linscan_opq 10k queries
X, Q = bench.make_data(1000000, 10000, 128, 1000, dev)
R, _ = torch.linalg.qr(torch.randn(128, 128, device=dev, generator=gen))
RX, RQ = (X @ R).contiguous(), (Q @ R).contiguous()          # R'X in Julia's d-by-n terms
Cpq = pq_codebooks(RX, 8, gen)
t_rot, _ = ev(lambda: (X @ R).contiguous())
t_enc, B = ev(lambda: core.quantize_pq(RX, Cpq, 8))
ix = core.Index(core.SCAN_PQ, B)
gt = bench.exact_nn(X, Q)
for k in (1, 1000):
    t, (dd, ii) = ev(lambda: ix.search(RQ, Cpq, k))
    out[f'config2_opq_scan_k{k}'] = dict(ms=t, qps=10000 / t * 1e3, recall_at_1=float((ii[:, 0].long() == gt).float().mean()),
                                         alg_GBs=10000 * 1e6 * 8 / t / 1e6)
out['config2_opq_encode'] = dict(rotate_ms=t_rot, quantize_pq_ms=t_enc, vectors_per_s=1e6 / ((t_rot + t_enc) * 1e-3))
print(json.dumps(out, indent=1))
