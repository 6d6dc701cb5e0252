"""Offline feasibility probe (CPU): packed-integer table pre-filter for an ICM step.
q = rint(T/Tmax_j * Q) with Q = floor(65535/(2(m-1))) so that the (m-1)-term sum of offset-encoded values fits 16 bits;
approx(c) = fma(scale, sum q, u).  Count candidates within the rigorous window 2*delta of the approximate minimum."""
import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
import bench
dev = torch.device('cpu')
n, d = 20000, 128
for m in (8, 16, 7):
    X, _ = bench.make_data(n, 10, d, 1000, dev)
    C = bench.train_codebooks(X[:20000].clone(), m, dev).numpy().reshape(m, 256, d)
    X = X.numpy()
    T = np.einsum('jcd,kbd->jkbc', C, C).astype(np.float32) * 2          # T[j][k][b][c]
    for j in range(m): T[j, j] = 0
    nrm = (C * C).sum(2)
    nv = 3000
    U = (-2 * np.einsum('jcd,nd->njc', C, X[:nv]) + nrm[None]).astype(np.float32)
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 65535 // (2 * (m - 1))
    tmax = np.abs(T).max((1, 2, 3))                                       # per j
    scale = (tmax / Q).astype(np.float32)
    q = np.rint(T / scale[:, None, None, None]).astype(np.int32)
    rng = np.random.default_rng(1)
    for label, sweeps in (("random codes", 0), ("after 2 ICM sweeps", 2)):
        B = rng.integers(0, 256, (nv, m))
        for sweep in range(sweeps):
            for j in range(m):
                ub = U[:, j, :].copy()
                for k in range(m):
                    if k != j: ub = ub + T[j, k, B[:, k], :]
                B[:, j] = ub.argmin(1)
        cnt, worst = [], 0.0
        for j in range(m):
            ub = U[:, j, :].copy(); qs = np.zeros((nv, 256), np.int32)
            for k in range(m):
                if k == j: continue
                ub = ub + T[j, k, B[:, k], :]
                qs += q[j, k, B[:, k], :]
            ua = (U[:, j, :].astype(np.float64) + scale[j].astype(np.float64) * qs).astype(np.float32)
            umax = np.abs(U[:, j, :]).max(1)
            delta = (m - 1) * 0.51 * scale[j] + 2.0 ** -20 * (umax + (m - 1) * tmax[j])
            err = np.abs(ua.astype(np.float64) - ub).max(1)
            assert (err <= delta).all(), (err.max(), delta.min())
            worst = max(worst, (err / delta).max())
            amin = ua.min(1)
            cnt.append((ua <= (amin + 2 * delta)[:, None]).sum(1))
        c = np.concatenate(cnt)
        print("m=%d Q=%d %-20s survivors: mean %.3f  P(>1) %.3f  p99 %d  max %d  (err/delta worst %.2f)"
              % (m, Q, label, c.mean(), (c > 1).mean(), np.percentile(c, 99), c.max(), worst))
