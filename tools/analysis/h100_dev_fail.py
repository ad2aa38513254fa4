"""Rebuild on the CPU the QP a device member failed on (gpurun_out/h100_fail_dev_m*.npz from tools/h100_dump_fail.py) and
run the numpy models of the device algorithm on it."""
import sys, pickle
sys.path.insert(0, '/root/repo')
import numpy as np, warnings
warnings.simplefilter('ignore')
from oracle import restate as rs
from mpc4quantum_b200 import systems
from tools.analysis.ipm_proto import ipm, polish, solve
k = int(sys.argv[1])
d = np.load('/root/repo/gpurun_out/h100_fail_dev_m%d.npz' % k)
step = int(d['step']); H = 100
cfg = systems.config_transmon(1, horizon=H, n_steps=20, discretize=rs.taylor_discretize)
c = 9
Xg = (d['Xg'][:, :c] + 1j * d['Xg'][:, c:]).T      # [c, H+1]
Ug = d['Ug'].T                                      # [m, H]
x = d['xs'][:, step]
us = d['us']
model = rs.BilinearModel(cfg['model'].A, 2, 1)
A_ls, B_ls, D_ls = model.along(Xg, Ug, H)
w0 = step - 1
X_ref = cfg['X_targ'][:, w0:w0 + H + 1]; U_ref = cfg['U_targ'][:, w0:w0 + H]
u_prev = us[:, step - 1]
args = (x, X_ref, U_ref, [cfg['Q']] * H + [cfg['Qf']], [cfg['R']] * H, A_ls, B_ls, D_ls, np.real(u_prev), cfg['sat'], cfg['du'])
growth = np.linalg.norm(np.linalg.multi_dot([rs.realify_op(m) for m in reversed(A_ls)]), 2)
print('member %d step %d: ||prod A|| %.1e, |Xg|max %.1e, |Ug|max %.3f' % (k, step, growth, np.abs(Xg).max(), np.abs(Ug).max()))
try:
    Xo, Uo, obj, info = rs.qp_exact(*args)
    print('oracle solves it: obj %.6f kkt %s' % (obj, info['kkt']))
    Uo = Uo.T
except Exception as e:
    print('oracle fails:', e); Uo = None
prob = rs._SparseQP(np.asarray(args[0]).reshape(-1), *args[1:8])
lo, hi = rs.qp_bounds(args[2], args[8], args[9], args[10]); lo, hi = lo.T.copy(), hi.T.copy()
print('stage-0 box', lo[0], hi[0])
u, zl, zu, n = ipm(prob, lo, hi, verbose=True)
U, r = polish(prob, lo, hi, u, zl, zu)
print('IPM solves %d, polish rounds %s, err vs oracle %s' % (n, r, None if Uo is None else np.abs(U - Uo).max()))
pickle.dump(args, open('/root/repo/tools/analysis/h100_dev_m%d.pkl' % k, 'wb'))
