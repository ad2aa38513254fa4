"""Capture the QPs of the oracle loop at H=100 order 1 (transmon) and study conditioning."""
import sys, pickle, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import restate as rs
from mpc4quantum_b200 import systems

H = int(sys.argv[1]) if len(sys.argv) > 1 else 100
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = systems.config_transmon(1, horizon=H, n_steps=NS, discretize=rs.taylor_discretize)
cap = []
def qp(*a):
    t0 = time.time()
    out = rs.qp_exact(*a)
    cap.append(dict(args=a, X=out[0], U=out[1], obj=out[2], kkt=out[3]['kkt'], t=time.time() - t0))
    return out
plant = rs.ExpmPlant(cfg['experiment'].H0, cfg['experiment'].H1_list)
stats = {}
xs, us, ec = rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt,
                         H, NS, plant, cfg['model'].A, cfg['Q'], cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], qp=qp, stats=stats)
print('ec', ec, 'qp per step', stats['qp_per_step'], 'n qps', len(cap))
print('kkt', [c['kkt'] for c in cap][:10], 'time/qp', np.mean([c['t'] for c in cap]))
pickle.dump(cap, open('/root/repo/tools/analysis/h%d_qps.pkl' % H, 'wb'))
