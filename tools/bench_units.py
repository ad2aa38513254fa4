#!/usr/bin/env python
"""Per-kernel throughput of the stand-alone entry points (GPU box), with the roofline that bounds each:

  m4q_expm_step_batched   plant step rho <- U rho U^+, U = expm(-i H dt): HBM bound (streams H0, H1, u, rho)
  m4q_linearize_batched   A_t, B_t, Delta_t along a guess trajectory: HBM bound (writes A_t [H, c, c] per instance)
  m4q_qp_admm_batched     the horizon QP: fp64 bound (same Riccati kernels as the fused loop)

    python tools/bench_units.py [N]       prints one JSON line per kernel
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch                                                      # noqa: E402
import mpc4quantum_b200 as m4q                                    # noqa: E402
from mpc4quantum_b200 import systems, optimize, _lib              # noqa: E402
from mpc4quantum_b200.experiment import expm_segments            # noqa: E402

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'MEASURED_PEAKS.json'))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'MEASURED_PEAKS.json')) else {}
HBM = PEAKS.get('hbm_gbs', 6650.0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    cfg = systems.config_transmon(1)
    ens, _ = systems.ensemble_transmon(N)
    d, m, c = 3, 2, 9
    rng = np.random.default_rng(0)

    # ---- plant step
    H0 = _lib.dev(ens.H0, np.complex128)
    H1 = _lib.dev(ens.H1, np.complex128)
    u = _lib.dev(rng.uniform(-1.5, 1.5, (N, 1, m)), np.float64)
    rho = _lib.dev(np.tile(cfg['x0'][None], (N, 1)), np.complex128)
    t = timed(lambda: expm_segments(rho, H0, H1, u, cfg['clock'].dt))
    bytes_ = N * (16 * (d * d + m * d * d + 2 * d * d) + 8 * m)
    print(json.dumps({'kernel': 'm4q_expm_step_batched', 'instances': N, 'seconds': t, 'propagations_per_s': N / t,
                      'bound': 'hbm', 'algorithmic_bytes': bytes_, 'achieved_gbs': bytes_ / t / 1e9, 'peak_gbs': HBM,
                      'frac': bytes_ / t / 1e9 / HBM, 'note': 'includes the output allocation of the Python wrapper'}))

    # ---- linearisation
    nb, H = min(N, 1 << 17), 16
    wm = m4q.WrapModel(*cfg['model'].get_discrete(), m, 1)
    X = _lib.dev(rng.standard_normal((nb, c, H + 1)) + 1j * rng.standard_normal((nb, c, H + 1)), np.complex128)
    U = _lib.dev(rng.uniform(-1.5, 1.5, (nb, m, H)), np.float64)
    t = timed(lambda: wm._along(X, U, H))
    bytes_ = nb * (16 * c * (H + 1) + 8 * m * H + 16 * H * (c * c + c * m + c))
    print(json.dumps({'kernel': 'm4q_linearize_batched', 'instances': nb, 'horizon': H, 'seconds': t,
                      'linearisations_per_s': nb / t, 'bound': 'hbm', 'algorithmic_bytes': bytes_,
                      'achieved_gbs': bytes_ / t / 1e9, 'peak_gbs': HBM, 'frac': bytes_ / t / 1e9 / HBM}))

    # ---- batched QP (golden transmon QPs replicated)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'qp.npz'))
    tag = 'transmon'
    k = g['%s_x_init' % tag].shape[0]
    nq = min(N, 1 << 15)
    idx = np.arange(nq) % k
    Hq = g['%s_U' % tag].shape[2]
    Q = np.broadcast_to(np.stack([g['%s_Q' % tag]] * Hq + [g['%s_Qf' % tag]]), (nq, Hq + 1, c, c))
    R = np.broadcast_to(np.stack([g['%s_R' % tag]] * Hq), (nq, Hq, m, m))
    args = [_lib.dev(np.ascontiguousarray(a), dt) for a, dt in (
        (g['%s_x_init' % tag][idx], np.complex128), (g['%s_X_bm' % tag][idx], np.complex128),
        (g['%s_U_bm' % tag][idx], np.float64), (Q, np.complex128), (R, np.float64), (g['%s_A' % tag][idx], np.complex128),
        (g['%s_B' % tag][idx], np.complex128), (g['%s_D' % tag][idx], np.complex128),
        (g['%s_u_prev' % tag][idx], np.float64))]
    sat, du = float(g['%s_sat' % tag]), float(g['%s_du' % tag])
    out = {}

    def run():
        out['r'] = optimize.quad_program_batched(*args, sat, du)
    t = timed(run, reps=3)
    X_, U_, obj, status, iters = out['r']
    err = np.abs(U_.cpu().numpy() - g['%s_U' % tag][idx]).max()
    n, H = 2 * c, Hq
    fac = float(iters.cpu().numpy()[:, 1].sum())
    F_fac = H * (4 * n ** 3 + 6 * n * n * m + 2 * n * m * m + m ** 3)
    F_it = H * (4 * n * n + 8 * n * m)
    flops = fac * (F_fac + 2 * F_it)
    print(json.dumps({'kernel': 'm4q_qp_admm_batched', 'instances': nq, 'horizon': H, 'seconds': t, 'qp_solves_per_s': nq / t,
                      'factorizations_per_qp': fac / nq, 'max_abs_control_error_vs_oracle': err, 'bound': 'fp64',
                      'achieved_tflops': flops / t / 1e12,
                      'note': 'cold start (no warm working set), includes realification of the instance data'}))


if __name__ == '__main__':
    main()
