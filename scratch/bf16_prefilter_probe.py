"""Offline feasibility probe (CPU): how many of the 256 candidates of an ICM step survive a bf16-table pre-filter
with a rigorous error window?  exact: ((u + r1) + r2) ... fp32; approx: same chain with bf16(r_k)."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device('cpu')
torch.manual_seed(0)
n, d = 20000, 128
for m in (8, 16):
    X, _ = bench.make_data(n, 10, d, 1000, dev)
    C = bench.train_codebooks(X[:20000].clone(), m, dev).numpy().reshape(m, 256, d)
    X = X.numpy()
    T = np.einsum('jcd,kbd->jkbc', C, C).astype(np.float32) * 2          # T[j][k][b][c]
    def bf16(x):
        u = x.view(np.uint32).astype(np.uint64)
        r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16                    # round to nearest even
        return r.astype(np.uint32).view(np.float32)
    Tb = bf16(T.copy())
    nrm = (C * C).sum(2)
    U = (-2 * np.einsum('jcd,nd->njc', C, X[:2000]) + nrm[None]).astype(np.float32)
    rng = np.random.default_rng(1)
    rowmax = np.abs(T).max(3)               # [j][k][b]
    tabmax = np.abs(T).max((2, 3))          # [j][k]
    # use codes after a few ICM sweeps (realistic): start random, run 2 sweeps exact
    B = rng.integers(0, 256, (2000, m))
    for sweep in range(2):
        for j in range(m):
            ub = U[:, j, :].copy()
            for k in range(m):
                if k != j: ub = ub + T[j, k, B[:, k], :]
            B[:, j] = ub.argmin(1)
    cnt_row, cnt_tab, miss = [], [], 0
    for j in range(m):
        ub = U[:, j, :].copy(); ua = U[:, j, :].copy()
        drow = np.zeros(2000, np.float32); dtab = 0.0
        for k in range(m):
            if k == j: continue
            ub = ub + T[j, k, B[:, k], :]
            ua = ua + Tb[j, k, B[:, k], :]
            drow += rowmax[j, k, B[:, k]] * 2.0 ** -8
            dtab += tabmax[j, k] * 2.0 ** -8
        mag = np.abs(U[:, j, :]).max(1) + drow * 2 ** 8
        drow = drow + mag * 2.0 ** -20; dtab = dtab + mag.max() * 2.0 ** -20
        err = np.abs(ua - ub).max(1)
        assert (err <= drow).all(), (err.max(), drow.min())
        amin = ua.min(1)
        cnt_row.append((ua <= (amin + 2 * drow)[:, None]).sum(1))
        cnt_tab.append((ua <= (amin + 2 * dtab)[:, None]).sum(1))
    cr, ct = np.concatenate(cnt_row), np.concatenate(cnt_tab)
    print("m=%d  survivors (row-max window): mean %.2f  p50 %d  p90 %d  p99 %d  max %d | (table-max window): mean %.2f p90 %d p99 %d"
          % (m, cr.mean(), np.percentile(cr, 50), np.percentile(cr, 90), np.percentile(cr, 99), cr.max(), ct.mean(),
             np.percentile(ct, 90), np.percentile(ct, 99)))
