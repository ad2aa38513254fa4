"""Parity of BASELINE configs 2, 3 and 4 on 64 ensemble members each, against fixtures produced by the REFERENCE's own
``mpc()`` (oracle/make_golden_ens64.py; exact-QP and expm-plant leaves injected, see oracle/refshim.py).

Two tests per config, both at the north_star tolerances with no conditioning-dependent slack:

* closed loop: ``mpc_ensemble`` on the 64 plants -> us within 1e-5, final fidelity within 1e-6, SQP counts per step equal,
  for every member whose closed loop the reference itself reproduces: the fixture records, per member, the gap between
  the reference's ``mpc()`` and its numpy restatement (two CPU evaluations of the same algorithm with the same exact QP
  leaf).  Where that gap is below 1e-8 the loop is well conditioned and the GPU is held to the north_star tolerance; where
  two CPU runs of the reference algorithm already part by more than that (some mismatched qubit / crosstalk plants: the
  39-iteration SQP of step 0 and 20-50 closed-loop steps amplify round-off up to 1e-3, and for one crosstalk member
  even the SQP count of a step differs between the two CPU runs) a closed-loop comparison says nothing about the
  implementation -- those members are pinned by the teacher-forced test below, and here only before amplification
  (first three steps at 1e-7);
* teacher forcing: every MPC step of every member is replayed from the state the REFERENCE run was in when the step
  started (measured state, guesses, previous control) through the host-stepped entry of the fused kernel; the applied
  control of every step must match to 1e-8 and the SQP iteration count exactly.  This pins each of the 64 x S QP
  sequences on its own, without the closed loop amplifying an earlier difference.

The achieved gaps are written to gpurun_out/r2_parity_gaps.json (copied to profiles/ after the run).
"""
import json
import os

import numpy as np
import pytest

import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems, _lib
from mpc4quantum_b200.mpc import ClosedLoopPlan
from conftest import load_golden, ROOT

pytestmark = pytest.mark.gpu

U_TOL, F_TOL, TF_TOL = 1e-5, 1e-6, 1e-8
REPRO_TOL = 1e-8          # reference mpc() vs its numpy restatement: below this the member's closed loop is reproducible
MIN_REPRODUCIBLE = {'qubit': 30, 'transmon': 64, 'crosstalk': 50, 'transmon_o2_h100': 8}
# order-1 model at H = 50: the QPs themselves are ill conditioned (cost-to-go entries ~1e10; two CPU runs of the reference
# algorithm agree to 1e-6 over the closed loop): single steps are held to 1e-6, still 10x inside the north_star
TF_TOL_BY_CONFIG = {'transmon_h50': 1e-6, 'transmon_h100': 1e-6}

CONFIGS = {
    'qubit': (lambda: systems.config_qubit(1), systems.ensemble_qubit, 4096),
    'transmon': (lambda: systems.config_transmon(1), systems.ensemble_transmon, 65536),
    'crosstalk': (lambda: systems.config_crosstalk(0.0), systems.ensemble_crosstalk, 65536),
    # BASELINE config 3 at its longest horizon with the order-2 model: 8 members, closed loop AND teacher forcing
    # (two CPU runs of the reference algorithm agree to 3e-10 over this closed loop)
    'transmon_o2_h100': (lambda: systems.config_transmon(2, horizon=100, n_steps=20), systems.ensemble_transmon, 65536),
}
# teacher forcing only: the order-1 model at H = 50 (cost-to-go entries ~1e10, every step from the third on needs the ADMM
# re-seeding of the working set): 16 members x 20 steps
TF_CONFIGS = dict(CONFIGS, transmon_h50=(lambda: systems.config_transmon(1, horizon=50, n_steps=20),
                                         systems.ensemble_transmon, 65536),
                  # H = 100: ||prod A_t|| ~ 4e13, every QP from the fourth step on goes through the pivoted KKT solve
                  # (csrc/m4q_kkt.cuh); 16 members x 20 steps
                  transmon_h100=(lambda: systems.config_transmon(1, horizon=100, n_steps=20),
                                 systems.ensemble_transmon, 65536),
                  # CNOT state (c = 16, m = 3, H = 50, order-1 model): the 40 steps of tests/golden/loop_cnot.npz, nominal
                  # plant, replayed step by step (oracle/make_golden_cnot.py 40 tf)
                  cnot=(lambda: systems.config_cnot(n_steps=40, horizon=50, ramp_steps=200), None, 1))


def _record(name, **vals):
    out_dir = os.path.join(ROOT, 'gpurun_out')
    if not os.path.isdir(out_dir):
        return
    path = os.path.join(out_dir, 'r2_parity_gaps.json')
    data = {}
    if os.path.exists(path):
        with open(path) as fh:
            data = json.load(fh)
    data.setdefault(name, {}).update(vals)
    with open(path, 'w') as fh:
        json.dump(data, fh, indent=1, sort_keys=True)


@pytest.mark.parametrize('name', list(CONFIGS))
def test_closed_loop_64_members_match_reference(name):
    make, maker, n_total = CONFIGS[name]
    g = load_golden('ens64_' + name)
    cfg = make()
    ens, _ = maker(n_total)
    k = g['us'].shape[0]
    args, kw = systems.mpc_args(cfg)
    kw.pop('progress_bar')
    res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, k), *args[7:], fid_target=cfg['target'], **kw)
    assert (res.exit_code == 0).all() and (res.steps_done == cfg['clock'].n_steps).all()
    du = np.abs(res.us - g['us']).reshape(k, -1).max(axis=1)
    df = np.abs(res.fidelity - g['fidelity'])
    dx = np.abs(res.xs[:, :, -1] - g['xs'][:, :, -1]).reshape(k, -1).max(axis=1)
    cnt_same = (res.qp_count == g['qp_per_step']).all(axis=1)
    repro = (g['restatement_gap'] < REPRO_TOL) & g['restatement_counts_equal']
    _record(name + '_closed_loop', members=int(k), reproducible_members=int(repro.sum()),
            max_us_gap_reproducible=float(du[repro].max()), max_fidelity_gap_reproducible=float(df[repro].max()),
            max_us_gap=float(du.max()), max_fidelity_gap=float(df.max()),
            max_final_state_gap=float(dx.max()), median_us_gap=float(np.median(du)),
            members_over_us_tol=int((du >= U_TOL).sum()), members_over_fid_tol=int((df >= F_TOL).sum()),
            sqp_counts_equal=int(cnt_same.sum()),
            reference_vs_restatement_max_gap=float(g['restatement_gap'].max()),
            per_member_us_gap=[float(x) for x in du], per_member_fidelity_gap=[float(x) for x in df],
            per_member_reference_vs_restatement_gap=[float(x) for x in g['restatement_gap']])
    assert repro.sum() >= MIN_REPRODUCIBLE[name], repro.sum()
    assert (du[repro] < U_TOL).all(), (np.flatnonzero(repro & (du >= U_TOL)), du[repro].max())
    assert (df[repro] < F_TOL).all(), (np.flatnonzero(repro & (df >= F_TOL)), df[repro].max())
    assert cnt_same[repro].all(), np.flatnonzero(repro & ~cnt_same)
    # every member, reproducible or not, before the loop can amplify anything
    assert np.abs(res.us[:, :, :3] - g['us'][:, :, :3]).max() < 1e-7
    assert (res.qp_count[:, :3] == g['qp_per_step'][:, :3]).all()


def _realify_traj(X):
    """[n, c, T] complex -> [n, T, 2c] (time major, [Re | Im]): the layout of the kernel's guess trajectory."""
    Xt = np.transpose(X, (0, 2, 1))
    return np.concatenate([Xt.real, Xt.imag], axis=2)


@pytest.mark.parametrize('name', list(TF_CONFIGS))
def test_teacher_forced_steps_match_reference(name):
    make, maker, n_total = TF_CONFIGS[name]
    g = load_golden('ens64_' + name)
    cfg = make()
    k, S, c, H1 = g['tf_Xg'].shape
    H, m = H1 - 1, cfg['dim_u']
    N = 2 * c
    plan = ClosedLoopPlan(cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'], cfg['model'], cfg['Q'],
                          cfg['R'], cfg['Qf'], cfg['sat'], cfg['du'], d=0, max_iter=100, warm_start=cfg['warm_start'],
                          capacity=k, external_plant=True)
    torch = _lib.torch()
    state = plan.state.view(torch.float64).view(k, -1)
    x0 = _lib.dev(g['tf_x'][:, 0], np.complex128)
    us_gap, cnt_bad = np.zeros((k, S)), np.zeros((k, S), dtype=bool)
    for s in range(S):
        plan.xs[:, :, s] = _lib.dev(g['tf_x'][:, s], np.complex128)
        if s > 0:
            # what the reference loop held when step s started: guesses (mpc.py:271-272) and the previous control
            state[:, :(H + 1) * N] = _lib.dev(_realify_traj(g['tf_Xg'][:, s]).reshape(k, -1), np.float64)
            state[:, (H + 1) * N:(H + 1) * N + H * m] = _lib.dev(
                np.transpose(g['tf_Ug'][:, s], (0, 2, 1)).reshape(k, -1), np.float64)
            plan.us[:, :, s - 1] = _lib.dev(g['tf_us'][:, :, s - 1], np.float64)
        res = plan.run(x0, n=k, step_begin=s, step_end=s + 1)
        assert int((res.exit_code != 0).sum()) == 0, (s, res.exit_code.cpu().numpy())
        us_gap[:, s] = np.abs(res.us[:, :, s].cpu().numpy() - g['tf_us'][:, :, s]).max(axis=1)
        cnt_bad[:, s] = res.qp_count[:, s].cpu().numpy() != g['tf_qp_per_step'][:, s]
    _record(name + '_teacher_forced', members=int(k), steps=int(S), qp_sequences=int(k * S),
            max_us_gap=float(us_gap.max()), sqp_count_mismatches=int(cnt_bad.sum()),
            qp_solves_checked=int(g['tf_qp_per_step'].sum()))
    tol = TF_TOL_BY_CONFIG.get(name, TF_TOL)
    assert us_gap.max() < tol, (us_gap.max(), np.argwhere(us_gap >= tol)[:5])
    assert not cnt_bad.any(), np.argwhere(cnt_bad)[:5]
