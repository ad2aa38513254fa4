import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems
from conftest import load_golden
g = load_golden('loop_cnot'); cfg = systems.config_cnot(n_steps=40, horizon=50, ramp_steps=200)
args, kw = systems.mpc_args(cfg)
(xs, us), _, ec = m4q.mpc(*args, **kw)
np.set_printoptions(linewidth=220, precision=1)
print('us gap per step', np.abs(us - g['us']).max(axis=0))
