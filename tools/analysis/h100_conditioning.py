import sys, pickle
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools/analysis')
import numpy as np
from oracle import restate as rs
exec(open('/root/repo/tools/analysis/h100_riccati.py').read().split("for qi, q in enumerate(cap):")[0])
rng = np.random.default_rng(0)
for qi in (3, 4, 5, 7):
    q = cap[qi]; a = q['args']
    prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
    lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
    Uo = q['U'].T
    mask = np.where(Uo <= lo + 1e-13, 1, np.where(Uo >= hi - 1e-13, 2, 0))
    # oracle sensitivity to 1e-16 relative noise
    for noise in (1e-16, 1e-15):
        A2 = [x * (1 + noise * rng.normal(size=x.shape)) for x in a[5]]
        B2 = [x * (1 + noise * rng.normal(size=x.shape)) for x in a[6]]
        out = rs.qp_exact(a[0], a[1], a[2], a[3], a[4], A2, B2, a[7], a[8], a[9], a[10])
        print('QP %d oracle noise %.0e: |dU| %.3e  first col %.3e kkt %s' % (qi, noise, np.abs(out[1] - q['U']).max(), np.abs(out[1][:, 0] - q['U'][:, 0]).max(), out[3]['kkt']))
    for name, kw in [('fp64', {}), ('longdouble', dict(dtype=np.longdouble))]:
        U, X = riccati(prob, lo, hi, mask, **kw)
        g = prob.gradient(X, U)
        print('  %-12s |U - Uo| max %.3e  first col %.3e  free-grad(rollout X) %.3e |X|max %.3e' % (name, np.abs(U - Uo).max(), np.abs(U[0] - Uo[0]).max(), np.abs(g[mask == 0]).max(), np.abs(X).max()))
