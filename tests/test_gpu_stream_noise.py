"""Streaming model updates and measurement noise INSIDE the fused loop.

* streaming: ``OnlineDMDc.fit_iteration`` (model.py:295-313) per member on the device after every MPC step
  (mpc.py:281-285).  Checked against the reference's own streaming loop (tests/golden/streaming.npz, produced by
  oracle/make_golden_streaming.py from the reference's classes) and against a host replay of the reference update rule.
* noise: ``QExperiment.set_sigma`` (experiment.py:193-194, :212) from a counter-based generator.  Parity with the
  reference's global numpy generator can only be statistical; what is pinned exactly is the generator itself (a numpy
  restatement of Philox4x32-10 + Box-Muller below) and where the noise enters the loop.
"""
import numpy as np
import pytest
from scipy.linalg import expm

import mpc4quantum_b200 as m4q
from mpc4quantum_b200 import systems
from mpc4quantum_b200.linearize import WrapModel, krtimes
from conftest import load_golden

pytestmark = pytest.mark.gpu


# ----------------------------------------------------------------------------------------------------------
# numpy restatement of the device generator: Philox4x32-10 (Salmon et al., SC'11), key = seed, counter = (member lo,
# member hi, step, component); two 53-bit uniforms -> Box-Muller
# ----------------------------------------------------------------------------------------------------------
def philox4x32_10(counter, key):
    c = [np.uint64(x) for x in counter]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask, p0 & mask]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return [int(x) for x in c]


def normal_pair(seed, member, step, comp):
    c = philox4x32_10((member & 0xFFFFFFFF, member >> 32, step, comp), (seed & 0xFFFFFFFF, seed >> 32))
    u1 = (((c[1] << 32) | c[0]) >> 11) * 2.0 ** -53 + 2.0 ** -54
    u2 = (((c[3] << 32) | c[2]) >> 11) * 2.0 ** -53
    r = np.sqrt(-2.0 * np.log(u1))
    return r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10."""
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff, 0xffffffff)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _plant_step(member, rho_vec, u, dt):
    d = member.H0.shape[0]
    Ham = member.H0 + sum(ui * h for ui, h in zip(u, member.H1_list))
    U = expm(-1j * Ham * dt)
    return (U @ rho_vec.reshape(d, d) @ U.conj().T).reshape(-1)


def test_measurement_noise_in_the_fused_loop():
    cfg = systems.config_transmon(1, horizon=8, n_steps=10)
    ens, _ = systems.ensemble_transmon(4096)
    args, kw = systems.mpc_args(cfg)
    kw.pop('progress_bar')
    sigma, seed = 1e-3, 0x1234567890ABCDEF
    clean = m4q.mpc_ensemble(args[0], *args[1:6], ens, *args[7:], fid_target=cfg['target'], **kw)
    ens.set_sigma(sigma, seed)
    noisy = m4q.mpc_ensemble(args[0], *args[1:6], ens, *args[7:], fid_target=cfg['target'], **kw)
    again = m4q.mpc_ensemble(args[0], *args[1:6], ens, *args[7:], fid_target=cfg['target'], **kw)
    assert (noisy.exit_code == 0).all()
    assert np.array_equal(noisy.xs, again.xs) and np.array_equal(noisy.us, again.us)        # reproducible from the seed
    assert np.abs(noisy.xs[:, :, 1] - clean.xs[:, :, 1]).max() > 0.1 * sigma                # and actually there
    assert np.array_equal(noisy.us[:, :, 0], clean.us[:, :, 0])                             # step 0 precedes any measurement
    ens.set_sigma(sigma, seed + 1)
    other = m4q.mpc_ensemble(args[0], *args[1:6], ens, *args[7:], fid_target=cfg['target'], **kw)
    assert not np.array_equal(other.xs, noisy.xs)
    # a shard draws what the whole ensemble draws for the same members (the stream is keyed by the GLOBAL index)
    ens.set_sigma(sigma, seed)
    shard = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(1000, 1100), *args[7:], fid_target=cfg['target'], **kw)
    assert np.array_equal(shard.xs, noisy.xs[1000:1100]) and np.array_equal(shard.us, noisy.us[1000:1100])
    # where it enters: xs[step + 1] = plant(xs[step], us[step]) + sigma (n1 + i n2), the plant restarting from the NOISY
    # record (experiment.py:212 feeding mpc.py:259); n1, n2 from the generator restated above
    dt, S = cfg['clock'].dt, cfg['clock'].n_steps
    for k in (0, 77, 4095):
        member = ens.member(k)
        for step in range(S):
            resid = noisy.xs[k, :, step + 1] - _plant_step(member, noisy.xs[k, :, step], noisy.us[k, :, step], dt)
            want = np.array([complex(*normal_pair(seed, k, step, comp)) for comp in range(9)]) * sigma
            assert np.abs(resid - want).max() < 1e-12, (k, step, np.abs(resid - want).max())
    # statistics over the ensemble: every component of every step is N(0, sigma^2) + i N(0, sigma^2)
    resid = np.empty((256, S, 9), dtype=complex)
    for k in range(256):
        member = ens.member(k)
        for step in range(S):
            resid[k, step] = noisy.xs[k, :, step + 1] - _plant_step(member, noisy.xs[k, :, step], noisy.us[k, :, step], dt)
    z = np.concatenate([resid.real.ravel(), resid.imag.ravel()]) / sigma         # 46,080 samples
    assert abs(z.mean()) < 0.03 and abs(z.var() - 1) < 0.03
    assert abs(np.mean(z ** 4) - 3) < 0.15 and abs(np.mean(z ** 3)) < 0.06
    assert abs(np.corrcoef(resid.real.ravel(), resid.imag.ravel())[0, 1]) < 0.03


def test_noise_through_mpc_uses_the_fused_loop():
    """mpc() with experiment.set_sigma(): same call as in the reference; seeded, it is reproducible; the global numpy
    generator provides the seed otherwise (np.random.seed governs it, as it governs the reference's noise)."""
    cfg = systems.config_qubit(1)
    args, kw = systems.mpc_args(cfg)
    cfg['experiment'].set_sigma(1e-3, seed=7)
    (xs1, us1), _, ec1 = m4q.mpc(*args, **kw)
    (xs2, us2), _, ec2 = m4q.mpc(*args, **kw)
    assert ec1 == ec2 == 0 and np.array_equal(xs1, xs2)
    cfg['experiment'].set_sigma(1e-3)
    np.random.seed(3)
    (xs3, _), _, _ = m4q.mpc(*args, **kw)
    np.random.seed(3)
    (xs4, _), _, _ = m4q.mpc(*args, **kw)
    (xs5, _), _, _ = m4q.mpc(*args, **kw)
    assert np.array_equal(xs3, xs4) and not np.array_equal(xs4, xs5) and not np.array_equal(xs3, xs1)
    cfg['experiment'].set_sigma(0)
    g = load_golden('loop_qubit_o1')
    (xs0, us0), _, _ = m4q.mpc(*args, **kw)
    assert np.abs(us0 - g['us']).max() < 1e-5
    with pytest.raises(NotImplementedError, match='closed system'):
        cfg['experiment'].set('c_ops', [np.eye(2)])


def _replay_online_dmdc(model, cfg, xs, us, lift):
    """The reference's update rule (model.py:295-313 through mpc.py:281-285) replayed on the host over a trajectory."""
    wrapped = WrapModel(*model.get_discrete(), cfg['dim_u'], cfg['order'])
    for step in range(us.shape[1]):
        lu = wrapped.lift_u(us[:, step].reshape(-1, 1))
        lx = np.asarray(lift(xs[:, step])).reshape(-1, 1)
        model.fit_iteration(np.asarray(lift(xs[:, step + 1])).reshape(-1, 1), lx, krtimes(lu, lx))
    return model


def test_streaming_fused_matches_the_reference_loop():
    """mpc(streaming=True) with OnlineDMDc now stays in the fused kernel; fixture from the reference's own loop."""
    g = load_golden('streaming')
    cfg = systems.config_qubit_freq(1, n_steps=15)
    model = m4q.OnlineDMDc.from_bootstrap(4, 4, cfg['model'].A.shape[1] - 4, cfg['model'].A.copy(), alpha=1e2)
    args, kw = systems.mpc_args(cfg)
    args = list(args)
    args[7] = model
    (xs, us), model2, ec = m4q.mpc(*args, streaming=True, **kw)
    assert ec == 0 and model2 is model and model._iteration == 15
    assert np.abs(us - g['loop_us']).max() < 1e-5, np.abs(us - g['loop_us']).max()
    assert np.abs(xs - g['loop_xs']).max() < 1e-4
    assert np.abs(model.A - g['loop_A']).max() < 1e-4 and np.abs(model.P - g['loop_P']).max() < 1e-2
    assert np.abs(model.A - g['loop_A0']).max() > 1e-2
    # the device update rule itself, on the device's own trajectory, to round-off
    host = m4q.OnlineDMDc.from_bootstrap(4, 4, cfg['model'].A.shape[1] - 4, cfg['model'].A.copy(), alpha=1e2)
    _replay_online_dmdc(host, cfg, xs, us, cfg['experiment'].lift)
    assert np.abs(model.A - host.A).max() < 1e-9 * max(1.0, np.abs(host.A).max())
    assert np.abs(model.P - host.P).max() < 1e-9 * max(1.0, np.abs(host.P).max())


@pytest.mark.parametrize('config', ['transmon', 'crosstalk'])
def test_streaming_ensemble_members_update_their_own_models(config):
    if config == 'transmon':
        cfg, (ens, _) = systems.config_transmon(2, horizon=10, n_steps=8), systems.ensemble_transmon(65536)
    else:                                          # measure_freq = 2: the updated model predicts every other step
        cfg, (ens, _) = systems.config_crosstalk(0.0, n_steps=10), systems.ensemble_crosstalk(65536)
    c = cfg['model'].A.shape[0]
    dz = cfg['model'].A.shape[1]
    args, kw = systems.mpc_args(cfg)
    kw.pop('progress_bar')
    n = 48
    model = m4q.OnlineDMDc.from_bootstrap(c, c, dz - c, cfg['model'].A.copy(), alpha=10.0)
    res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, n), model, *args[8:], fid_target=cfg['target'],
                           streaming=True, **kw)
    assert (res.exit_code == 0).all()
    assert res.model_A.shape == (n, c, dz) and res.model_P.shape == (n, dz, dz)
    assert np.abs(res.model_A[0] - res.model_A[1]).max() > 1e-6          # every member learns its own plant
    lift = ens.lift
    for k in (0, 17, n - 1):
        host = m4q.OnlineDMDc.from_bootstrap(c, c, dz - c, cfg['model'].A.copy(), alpha=10.0)
        _replay_online_dmdc(host, cfg, res.xs[k], res.us[k], lift)
        assert np.abs(res.model_A[k] - host.A).max() < 1e-9 * max(1.0, np.abs(host.A).max()), k
        assert np.abs(res.model_P[k] - host.P).max() < 1e-9 * max(1.0, np.abs(host.P).max()), k
        # one member alone through mpc(streaming=True) (fused, N = 1): the same trajectory bit for bit
        solo = m4q.OnlineDMDc.from_bootstrap(c, c, dz - c, cfg['model'].A.copy(), alpha=10.0)
        (xs, us), _, ec = m4q.mpc(args[0], *args[1:6], ens.member(k), solo, *args[8:], streaming=True, **kw)
        assert ec == 0 and np.array_equal(us, res.us[k]) and np.array_equal(xs, res.xs[k])
    if cfg['clock'].measure_freq > 1:
        # the controller keeps its captured operators (reference quirk), but the model steps between measurements use
        # the updated ones: the trajectories differ from the non-streaming run from the first model step after an update
        plain = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, n), *args[7:], fid_target=cfg['target'], **kw)
        assert np.abs(plain.xs[:, :, 3] - res.xs[:, :, 3]).max() > 1e-9
        assert np.array_equal(plain.xs[:, :, :2], res.xs[:, :, :2])


def test_fidelity_convention():
    cfg = systems.config_transmon(1, horizon=8, n_steps=6)
    ens, _ = systems.ensemble_transmon(4096)
    args, kw = systems.mpc_args(cfg)
    kw.pop('progress_bar')
    a = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, 64), *args[7:], fid_target=cfg['target'], **kw)
    b = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, 64), *args[7:], fid_target=cfg['target'],
                         fidelity_convention='sqrt', **kw)
    assert np.abs(b.fidelity - np.sqrt(a.fidelity)).max() < 1e-15        # qutip.fidelity = sqrt(<psi|rho|psi>)
    assert np.array_equal(a.us, b.us)


def test_argument_checks_zeroed_tails_and_the_empty_rate_box():
    """shape mismatches raise instead of reading out of bounds; members that stop early leave zeros (not
    uninitialised memory) beyond steps_done; a stage-0 box emptied by the rate bound is exit code 3 (the reference's
    solver reports the problem infeasible, mpc.py:200-203)."""
    from mpc4quantum_b200.mpc import ClosedLoopPlan
    from mpc4quantum_b200 import _lib
    from mpc4quantum_b200.experiment import expm_segments
    cfg = systems.config_transmon(1, horizon=8, n_steps=6)
    ens, _ = systems.ensemble_transmon(4096)
    args, kw = systems.mpc_args(cfg)
    kw.pop('progress_bar')
    with pytest.raises(ValueError, match='dim_u'):           # a one-control ensemble for a two-control problem
        m4q.mpc_ensemble(args[0], *args[1:6], m4q.EnsembleQExperiment(ens.H0[:4], ens.H1[:4, :1]), *args[7:], **kw)
    with pytest.raises(ValueError, match='plant state'):
        m4q.mpc_ensemble(args[0][:4], *args[1:6], ens.slice(0, 4), *args[7:], **kw)
    with pytest.raises(IndexError, match='drive Hamiltonian'):
        bad = m4q.QExperiment(cfg['experiment'].H0, cfg['experiment'].H1_list[:1])
        m4q.mpc(args[0], *args[1:6], bad, *args[7:], **kw)
    with pytest.raises(IndexError, match='controls per segment'):
        expm_segments(np.zeros((2, 9), complex), ens.H0[:2], ens.H1[:2], np.zeros((2, 3, 1)), 0.25)
    # early exit: infidelity threshold met at once -> exit code 1 after one step, zeros beyond
    res = m4q.mpc_ensemble(args[0], *args[1:6], ens.slice(0, 8), *args[7:], fid_target=cfg['target'],
                           exit_infidelity=2.0, **kw)
    assert (res.exit_code == 1).all() and (res.steps_done == 1).all()
    assert np.abs(res.us[:, :, 1:]).max() == 0 and np.abs(res.xs[:, :, 2:]).max() == 0 and (res.qp_count[:, 1:] == 0).all()
    # empty rate box: reference control of step 0 far outside [-sat, sat] with a tight du
    cfg2 = systems.config_transmon(1, horizon=8, n_steps=6)
    U_targ = cfg2['U_targ'].copy()
    U_targ[:, 0] = 10 * cfg2['sat']
    res = m4q.mpc_ensemble(cfg2['x0'], cfg2['dim_u'], cfg2['order'], cfg2['X_targ'], U_targ, cfg2['clock'], ens.slice(0, 8),
                           cfg2['model'], cfg2['Q'], cfg2['R'], cfg2['Qf'], sat=cfg2['sat'], du=0.1, fid_target=cfg2['target'])
    assert (res.exit_code == 3).all() and (res.steps_done == 0).all()
