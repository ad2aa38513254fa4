"""Active-set update rules on captured QPs (H=50): rounds needed from the warm start (previous solution shifted)."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import restate as rs
cap = pickle.load(open('/root/repo/tools/analysis/h%s_qps.pkl' % sys.argv[1], 'rb'))

def mask_of(U, lo, hi): return np.where(U <= lo + 1e-13, 1, np.where(U >= hi - 1e-13, 2, 0))

def run(prob, lo, hi, mask, rule, max_rounds=60):
    seen = []
    for rnd in range(max_rounds):
        fixed = mask != 0
        vals = np.where(mask == 1, lo, np.where(mask == 2, hi, 0.0))
        X, U = prob.solve_fixed(fixed, vals)
        g = prob.gradient(X, U)
        gs = max(1.0, np.abs(g).max())
        free = mask == 0
        viol_lo = free & (U < lo - 1e-12); viol_hi = free & (U > hi + 1e-12)
        rel = ((mask == 1) & (g < -1e-10 * gs)) | ((mask == 2) & (g > 1e-10 * gs))
        if not (viol_lo.any() or viol_hi.any() or rel.any()):
            return rnd + 1, U
        nm = mask.copy()
        if rule == 'pd':            # device rule: all at once
            nm[viol_lo] = 1; nm[viol_hi] = 2; nm[rel] = 0
        elif rule == 'add_then_drop':   # add all violated; release only when nothing is violated
            if viol_lo.any() or viol_hi.any():
                nm[viol_lo] = 1; nm[viol_hi] = 2
            else:
                nm[rel] = 0
        elif rule == 'add_then_drop1':  # release only the worst multiplier
            if viol_lo.any() or viol_hi.any():
                nm[viol_lo] = 1; nm[viol_hi] = 2
            else:
                w = np.where(mask == 1, -g, np.where(mask == 2, g, -np.inf)); k = np.unravel_index(np.argmax(w), w.shape); nm[k] = 0
        elif rule == 'pd_then_safe':    # device rule for 4 rounds, then add_then_drop
            if rnd < 4:
                nm[viol_lo] = 1; nm[viol_hi] = 2; nm[rel] = 0
            elif viol_lo.any() or viol_hi.any():
                nm[viol_lo] = 1; nm[viol_hi] = 2
            else:
                nm[rel] = 0
        mask = nm
    return None, U

prev_mask = None
for qi, q in enumerate(cap):
    a = q['args']
    prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
    lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
    Uo = q['U'].T
    mo = mask_of(Uo, lo, hi)
    if prev_mask is not None:
        warm_same = prev_mask.copy()                                 # same step, next SQP iterate
        warm_shift = np.vstack([prev_mask[1:], prev_mask[-1:]])      # next MPC step
        out = []
        for rule in ('pd', 'add_then_drop', 'add_then_drop1', 'pd_then_safe'):
            r1, U1 = run(prob, lo, hi, warm_same, rule)
            r2, U2 = run(prob, lo, hi, warm_shift, rule)
            out.append('%s: same %s shift %s (err %.0e)' % (rule, r1, r2, np.abs(U2 - Uo).max() if r2 else np.nan))
        print('QP %2d pinned %3d | %s' % (qi, (mo != 0).sum(), ' | '.join(out)))
    prev_mask = mo
