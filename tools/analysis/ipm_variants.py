import sys, pickle, glob
sys.path.insert(0, '/root/repo')
import numpy as np, warnings
warnings.simplefilter('ignore')
from oracle import restate as rs
from tools.analysis.ipm_proto import solve, polish

def ipm2(prob, lo, hi, variant, tau=0.995, max_it=80, mu_tol=1e-11):
    H, m = lo.shape
    nofix = np.zeros((H, m), bool); zeros = np.zeros((H, m))
    u = 0.5 * (lo + hi); s1, s2 = u - lo, hi - u
    zl = np.ones((H, m)); zu = np.ones((H, m))
    n = 0; sigma = 0.1
    def maxstep(v, dv):
        neg = dv < 0
        return min(1.0, (tau * (-v[neg] / dv[neg])).min()) if neg.any() else 1.0
    for it in range(max_it):
        mu = (np.sum(s1 * zl) + np.sum(s2 * zu)) / (2 * H * m)
        if mu < mu_tol: break
        Sig = zl / s1 + zu / s2
        if variant == 'mehrotra':
            up, _ = solve(prob, nofix, zeros, Sig, Sig * u); n += 1      # affine: sigma = 0
            du = up - u
            dzl = -zl - zl / s1 * du; dzu = -zu + zu / s2 * du
            ap = min(maxstep(s1, du), maxstep(s2, -du)); ad = min(maxstep(zl, dzl), maxstep(zu, dzu))
            mu_aff = (np.sum((s1 + ap * du) * (zl + ad * dzl)) + np.sum((s2 - ap * du) * (zu + ad * dzu))) / (2 * H * m)
            sigma = (mu_aff / mu) ** 3
            # corrector: complementarity s z + ds dz(aff) = sigma mu
            cl = sigma * mu - du * dzl; cu = sigma * mu + du * dzu
            shift = Sig * u + cl / s1 - cu / s2
            up, _ = solve(prob, nofix, zeros, Sig, shift); n += 1
            du = up - u
            dzl = cl / s1 - zl - zl / s1 * du; dzu = cu / s2 - zu + zu / s2 * du
        else:
            shift = Sig * u + sigma * mu * (1 / s1 - 1 / s2)
            up, _ = solve(prob, nofix, zeros, Sig, shift); n += 1
            du = up - u
            dzl = sigma * mu / s1 - zl - zl / s1 * du; dzu = sigma * mu / s2 - zu + zu / s2 * du
        ap = min(maxstep(s1, du), maxstep(s2, -du)); ad = min(maxstep(zl, dzl), maxstep(zu, dzu))
        if variant == 'equal': ap = ad = min(ap, ad)
        u = u + ap * du; s1, s2 = u - lo, hi - u
        zl = zl + ad * dzl; zu = zu + ad * dzu
        if variant in ('adaptive', 'equal'):
            a = min(ap, ad)
            sigma = 0.1 if a > 0.5 else (0.3 if a > 0.1 else 0.8)
    return u, zl, zu, n, mu

if __name__ == '__main__':
    files = sorted(glob.glob('/root/repo/tools/analysis/h100_dev_m*.pkl') + glob.glob('/root/repo/tools/analysis/h100_m*_q*.pkl') + glob.glob('/root/repo/tools/analysis/h100_fail_m*.pkl'))
    for f in files:
        obj = pickle.load(open(f, 'rb'))
        a = obj if 'dev_m' in f else obj[0]
        prob = rs._SparseQP(np.asarray(a[0]).reshape(-1), *a[1:8])
        lo, hi = rs.qp_bounds(a[2], a[8], a[9], a[10]); lo, hi = lo.T.copy(), hi.T.copy()
        Uo = rs.qp_exact(*a)[1].T
        for v in ('fixed', 'adaptive', 'equal', 'mehrotra'):
            u, zl, zu, n, mu = ipm2(prob, lo, hi, v)
            U, r = polish(prob, lo, hi, u, zl, zu)
            print('%-22s %-9s solves %3d final mu %.1e polish %s err %.1e' % (f.split('/')[-1], v, n, mu, r, np.abs(U - Uo).max()))
