#!/usr/bin/env python
"""Benchmark of the closed-loop MPC hot path (BASELINE.json: closed-loop MPC trajectories/s over an ensemble of
perturbed transmon plants).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload transmon_h16] [--members 65536]
    python bench.py --impl reference ...      # the CPU path (oracle port of the reference) on the host cores

One "step" = one pass of the whole closed loop (all MPC steps of mpc.py:128-304) over this rank's shard of the
ensemble.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the flop model behind `roofline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'closed-loop MPC trajectories/s'
UNIT = 'trajectories/s'


# ----------------------------------------------------------------------------------------------------------
SYSTEM_NAMES = {'transmon': '3-level transmon', 'qubit': 'ideal qubit', 'crosstalk': 'two qubits with ZZ crosstalk',
                'not_gate': 'NOT-gate synthesis on the qubit process vector',
                'transmon_exact': '3-level transmon, exact-discretisation model'}


def workload(name, discretize=None):
    from mpc4quantum_b200 import systems
    if name.startswith('transmon_h') or name.startswith('transmon_o2_h') or name.startswith('transmon_models_h'):
        order = 2 if name.startswith('transmon_o2_h') else 1
        H = int(name.split('_h')[-1])
        cfg = systems.config_transmon(order, horizon=H, n_steps=20, discretize=discretize)
        # transmon_models_hH: every member also controls with its own perturbed MODEL (discretised per member)
        cfg['per_member_models'] = name.startswith('transmon_models_h')
        return cfg, systems.ensemble_transmon
    if name.startswith('transmon_exact_h'):    # exact-discretisation model mode (expm + Frechet derivative per stage)
        return systems.config_transmon_exact(horizon=int(name.split('_h')[-1]), n_steps=20), systems.ensemble_transmon
    if name == 'qubit':
        return systems.config_qubit(1, discretize=discretize), systems.ensemble_qubit
    if name == 'crosstalk':
        return systems.config_crosstalk(0.0, discretize=discretize), systems.ensemble_crosstalk
    if name == 'not_gate':      # gate synthesis on process vectors (c = 16, m = 1, H = 15, S = 50)
        return systems.config_not_gate(1, discretize=discretize), systems.ensemble_not_gate
    raise SystemExit('unknown workload %s' % name)


def flop_model(cfg, counters, qp_count):
    """Algorithmic real flops of a batch of trajectories from the device counters.

    counters [n, 4] = ADMM iterations, Riccati factorisations, polish rounds, QP solves.
    """
    exact = hasattr(cfg['model'], 'generators')
    c = cfg['model'].dim_x if exact else cfg['model'].A.shape[0]
    n, m = 2 * c, cfg['dim_u']
    p = m if exact else cfg['model'].A.shape[1] // c - 1
    H, S = cfg['clock'].horizon, cfg['clock'].n_steps
    d = cfg['experiment'].H0.shape[0]
    F_lin = H * (12 * p * c * c + 4 * p * c * m + 4 * c * m)
    if exact:   # per stage: expm by 16 Horner + ~3 squaring products (8 c^3 each), ~22 Taylor terms of (1 + 2m) complex
        #           mat-vecs for the Frechet vectors (term count depends on the generator norm; 22 at ||G dt|| ~ 1.5)
        F_lin = H * (19 * 8 * c ** 3 + 22 * (1 + 2 * m) * 8 * c * c)
    F_fac = H * (4 * n ** 3 + 6 * n * n * m + 2 * n * m * m + m ** 3)
    F_it = H * (4 * n * n + 8 * n * m)
    F_ls = 6 * (2 * c * (H + 1) + 2 * m * H)
    F_pl = 8 * d ** 3 * (18 + 2)
    admm, factor, polish, solves = [counters[:, i].astype(np.float64).sum() for i in range(4)]
    ls_calls = float(qp_count[:, :2].sum()) if cfg['warm_start'] else solves
    total = solves * F_lin + factor * F_fac + (admm + 2 * polish) * F_it + ls_calls * F_ls + counters.shape[0] * S * F_pl
    return total, dict(F_lin=F_lin, F_fac=F_fac, F_it=F_it, F_ls=F_ls, F_pl=F_pl)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(',')]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        sm = [float(r[0]) for r in self.rows]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [nm for i, nm in enumerate(names) if any(r[3 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons,
                'power_w_max': max(float(r[2]) for r in self.rows), 'samples': len(self.rows)}


# ----------------------------------------------------------------------------------------------------------
# CPU path: the oracle port of the reference loop (oracle/restate.py), one process per core
# ----------------------------------------------------------------------------------------------------------
def _cpu_member(job):
    name, k, n_total = job
    os.environ['OMP_NUM_THREADS'] = '1'
    from oracle import restate as rs
    cfg, maker = workload(name, discretize=rs.taylor_discretize)
    ens, _ = maker(n_total)
    mem = ens.member(k)
    lift, proj = (rs.lift_coupled, rs.proj_coupled) if cfg.get('kind') == 'coupled' else (rs.lift_identity, rs.lift_identity)
    plant = rs.ProcessPlant(mem.H0, mem.H1_list) if cfg.get('kind') == 'process' else \
        rs.ExpmPlant(mem.H0, mem.H1_list, lift, proj)
    stats = {}
    A_full = getattr(cfg['model'], 'A', None)
    if cfg.get('per_member_models'):
        from mpc4quantum_b200 import systems
        L, _ = systems.transmon_model_liouvillians(n_total)
        A_full = rs.taylor_discretize(list(L[k]), cfg['clock'].dt, cfg['order'])
    exact = rs.ExactModel(list(cfg['model'].generators), cfg['clock'].dt) if cfg['name'] == 'transmon_exact' else None
    xs, us, ec = rs.mpc_loop(cfg['x0'], cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'].dt,
                             cfg['clock'].horizon, cfg['clock'].n_steps, plant, A_full, cfg['Q'], cfg['R'],
                             cfg['Qf'], cfg['sat'], cfg['du'], warm_start=cfg['warm_start'],
                             measure_freq=cfg['clock'].measure_freq, stats=stats, model=exact)
    return float(np.real(np.vdot(cfg['target'], xs[:, -1]))), int(sum(stats['qp_per_step']))


def cpu_pass(name, members, n_total, pool):
    t0 = time.perf_counter()
    res = pool.map(_cpu_member, [(name, k, n_total) for k in members])
    dt = time.perf_counter() - t0
    return dt, res


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm (oracle port: reference loop semantics with exact QP and
    expm plant leaves, see oracle/restate.py) on all host cores.  Each step = a bounded sample of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = host_cores()
    per_step = max(cores, 8)
    n_total = args.members * args.gpus
    with mp.get_context('spawn').Pool(cores) as pool:
        cpu_pass(args.workload, range(min(cores, 4)), n_total, pool)          # import / page-in warm-up
        for w in range(args.warmup):
            cpu_pass(args.workload, range(w * per_step, (w + 1) * per_step), n_total, pool)
        t_total, qps = 0.0, 0
        for s in range(args.steps):
            dt, res = cpu_pass(args.workload, range(s * per_step, (s + 1) * per_step), n_total, pool)
            t_total += dt
            qps += sum(r[1] for r in res)
    value = args.steps * per_step / t_total
    sample = '%d members per step (first members of the same seeded ensemble), %d steps' % (per_step, args.steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s, %d perturbed plants per GPU' % (args.workload, args.members),
                   'note': 'CPU arm times a bounded sample and reports trajectories/s of the host'},
        'qp_solves_per_s': qps / t_total,
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample,
                         'what': 'reference loop semantics (mpc.py:128-304) restated in numpy with exact active-set QP '
                                 'and scipy expm plant leaves; cvxpy/OSQP/qutip are not installable offline'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='transmon_h16')
    ap.add_argument('--members', type=int, default=65536, help='ensemble members per GPU (weak scaling)')
    ap.add_argument('--cpu-seconds', type=float, default=20.0, help='budget of the cpu_baseline leg')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--admm-first', action='store_true',
                    help='always run an ADMM block before the active-set rounds (m4q_qp_settings.admm_first)')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import mpc4quantum_b200 as m4q
    from mpc4quantum_b200 import _lib
    from mpc4quantum_b200.ensemble import shard_bounds, fidelity_histogram, allreduce_histogram

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d' % args.gpus

    cfg, maker = workload(args.workload)
    n_total = args.members * world
    lo, hi = shard_bounds(n_total, rank, world)
    n = hi - lo
    ens_all, _ = maker(n_total)                      # same seeded draw on every rank; each keeps its block
    ens = ens_all.slice(lo, hi)
    model = cfg['model']
    if cfg.get('per_member_models'):
        from mpc4quantum_b200 import systems
        model = systems.ensemble_transmon_models(n_total, order=cfg['order'])[0].slice(lo, hi)
    margs = (cfg['dim_u'], cfg['order'], cfg['X_targ'], cfg['U_targ'], cfg['clock'], model, cfg['Q'], cfg['R'],
             cfg['Qf'], cfg['sat'], cfg['du'])
    plan = m4q.ClosedLoopPlan(*margs, d=ens.d, lift_mode=ens.lift_mode, warm_start=cfg['warm_start'],
                              fid_target=cfg['target'], capacity=n,
                              settings=_lib.qp_settings(admm_first=int(args.admm_first)))
    geom = plan.launch_info()

    # ---- resident inputs (value) and pinned host inputs (e2e)
    H0_h = torch.from_numpy(np.ascontiguousarray(ens.H0)).pin_memory()
    H1_h = torch.from_numpy(np.ascontiguousarray(ens.H1)).pin_memory()
    x0_plant = cfg['u0'] if cfg.get('kind') == 'process' else cfg['x0']     # gate synthesis: the propagator itself
    x0_h = torch.from_numpy(np.ascontiguousarray(x0_plant.reshape(1, -1))).pin_memory()
    H0_d, H1_d, x0_d = H0_h.cuda(), H1_h.cuda(), x0_h.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')      # > 126 MB L2
    hist = torch.zeros(256, dtype=torch.int64, device='cuda')
    S, m = cfg['clock'].n_steps, cfg['dim_u']
    fid_out = torch.empty(n, dtype=torch.float64).pin_memory()
    ec_out = torch.empty(n, dtype=torch.int32).pin_memory()
    us_out = torch.empty((n, m, S), dtype=torch.float64).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def resident_pass():
        res = plan.run(x0_d, H0_d, H1_d, n=n, x0_shared=True)
        hist.zero_()
        fidelity_histogram(res.fidelity, 0.0, 1.0, 256, hist)
        if world > 1:
            allreduce_histogram(hist)
        return res

    def e2e_pass():
        h0 = H0_h.cuda(non_blocking=True)
        h1 = H1_h.cuda(non_blocking=True)
        x0 = x0_h.cuda(non_blocking=True)
        res = plan.run(x0, h0, h1, n=n, x0_shared=True)
        fid_out.copy_(res.fidelity, non_blocking=True)
        ec_out.copy_(res.exit_code, non_blocking=True)
        us_out.copy_(res.us, non_blocking=True)
        torch.cuda.synchronize()
        return res

    def timed(fn, steps):
        """Device time of `steps` passes, L2 flushed between passes (flush not timed); max over ranks."""
        total_ms, kern_ms = 0.0, []
        for _ in range(steps):
            flush.fill_(1)
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            sync_all()
            total_ms += e0.elapsed_time(e1)
            kern_ms.append(e0.elapsed_time(e1))
        t = torch.tensor([total_ms], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), kern_ms

    for _ in range(args.warmup):
        res = resident_pass()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    total_ms, per_pass = timed(resident_pass, args.steps)
    sampler.stop_flag.set()
    sampler.join()
    res = resident_pass()
    sync_all()
    counters = res.counters.cpu().numpy()
    qp_count = res.qp_count.cpu().numpy()
    exit_codes = res.exit_code.cpu().numpy()
    fid = res.fidelity.cpu().numpy()

    # ---- the dominant kernel alone (it is the whole pass but for the table build and the histogram)
    kern_ms = []
    for _ in range(max(2, min(args.steps, 3))):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.run(x0_d, H0_d, H1_d, n=n, x0_shared=True)
        e1.record()
        torch.cuda.synchronize()
        kern_ms.append(e0.elapsed_time(e1))
    kernel_ms = float(np.mean(kern_ms))

    # ---- fp64 roofline denominator, measured on this part now
    scratch = torch.zeros(8, dtype=torch.float64, device='cuda')
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    ctas, iters = sms * 8, 1 << 16
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().m4q_fp64_fma_probe(ctas, iters, _lib.ptr(scratch), _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 2.0 * 16 * iters * 256 * ctas / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    fp64_peak = best
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().m4q_fp64_dmma_probe(ctas, iters >> 2, _lib.ptr(scratch), _lib.stream_ptr()))
        e1.record()
        torch.cuda.synchronize()
        best = max(best, 512.0 * 8 * (iters >> 2) * 8 * ctas / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    dmma_peak = best

    e2e_pass()                                    # untimed: first use of the pinned staging buffers / allocator blocks
    sync_all()
    e2e_steps = max(2, min(args.steps, 3))
    e2e_ms, e2e_per_pass = timed(e2e_pass, e2e_steps)

    flops, per_unit = flop_model(cfg, counters, qp_count)
    traj_total = n_total * args.steps
    value = traj_total / (total_ms * 1e-3)
    qp_total = float(counters[:, 3].sum())
    if world > 1:
        t = torch.tensor([qp_total, flops], dtype=torch.float64, device='cuda')
        dist.all_reduce(t)
        qp_total_all, flops_all = t.tolist()
    else:
        qp_total_all, flops_all = qp_total, flops
    achieved = flops / (kernel_ms * 1e-3) / 1e12
    in_bytes = H0_h.numel() * 16 + H1_h.numel() * 16 + x0_h.numel() * 16
    out_bytes = fid_out.numel() * 8 + ec_out.numel() * 4 + us_out.numel() * 8
    hbm_alg = in_bytes + res.xs.numel() * 16 + out_bytes + counters.nbytes + qp_count.nbytes

    # DRAM traffic of the dominant kernel: from the committed ncu capture (profiles/), scaled to this launch's members
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r1_final_ncu_traffic.json')) as fh:
            tr = json.load(fh)
        if args.workload == 'transmon_h16':
            traffic = (tr['dram_bytes_read'] + tr['dram_bytes_write']) * n / tr['members']
            traffic_src = ('ncu dram__bytes_read.sum + dram__bytes_write.sum of one %d-member launch '
                           '(profiles/r1_final_ncu_traffic.json), scaled by members' % tr['members'])
    except (OSError, KeyError, ValueError):
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': '%s: %s (c=%d, m=%d), horizon %d, %d MPC steps, %d perturbed plants per GPU '
                               '(seed 20220113), tight QP mode (%s)' % (args.workload, SYSTEM_NAMES.get(cfg['name'], cfg['name']),
                                                                   (cfg['model'].dim_x if hasattr(cfg['model'], 'generators') else cfg['model'].A.shape[0]), cfg['dim_u'],
                                                                   cfg['clock'].horizon, S, args.members,
                                                                   'ADMM block first' if args.admm_first else
                                                                   'warm active set first, ADMM fallback'),
                   'members_total': n_total, 'l2': 'flushed between timed passes (256 MB write, untimed)',
                   'launch': geom},
        'qp_solves_per_s': qp_total_all * args.steps / (total_ms * 1e-3),
        'qp_solves_per_trajectory': qp_total / n,
        'admm_iterations_per_qp': float(counters[:, 0].sum() / max(qp_total, 1)),
        'factorizations_per_qp': float(counters[:, 1].sum() / max(qp_total, 1)),
        'exit_codes': {str(k): int((exit_codes == k).sum()) for k in np.unique(exit_codes)},
        'fidelity': {'min': float(fid.min()), 'median': float(np.median(fid)), 'max': float(fid.max())},
        'e2e': {'value': n_total * e2e_steps / (e2e_ms * 1e-3), 'unit': UNIT, 'ms_per_pass': e2e_per_pass,
                'h2d_bytes_per_step': int(in_bytes),
                'd2h_bytes_per_step': int(out_bytes)},
        'gpu_launches': 3 * args.steps,     # build_tables + mpc_kernel + hist_kernel per pass
        'roofline': {'bound': 'fp64', 'achieved': achieved, 'peak': fp64_peak, 'unit': 'TFLOP/s',
                     'frac': achieved / fp64_peak, 'traffic': traffic, 'traffic_source': traffic_src,
                     'peak_source': 'm4q_fp64_fma_probe measured in this run (MEASURED_PEAKS.json has no fp64 figure)',
                     'dmma_probe_tflops': dmma_peak,
                     'kernel': 'mpc_kernel', 'kernel_ms': kernel_ms, 'flops_per_launch': flops,
                     'flops_per_trajectory': flops / n, 'flop_model': per_unit,
                     'hbm': {'algorithmic_bytes': int(hbm_alg), 'achieved_gbs': hbm_alg / (kernel_ms * 1e-3) / 1e9}},
        'clocks': sampler.summary(),
    }
    if rank == 0 and not args.no_cpu and world == 1:
        import multiprocessing as mp
        cores = host_cores()
        with mp.get_context('spawn').Pool(cores) as pool:
            cpu_pass(args.workload, range(min(cores, 4)), n_total, pool)
            t0 = time.perf_counter()
            k_done, qps, cpu_fid = 0, 0, []
            while time.perf_counter() - t0 < args.cpu_seconds and k_done < n:
                dt, r = cpu_pass(args.workload, range(k_done, k_done + cores), n_total, pool)
                cpu_fid += [x[0] for x in r]
                qps += sum(x[1] for x in r)
                k_done += cores
            cpu_t = time.perf_counter() - t0
        line['cpu_baseline'] = {
            'value': k_done / cpu_t, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': 'first %d members of the same ensemble, %.1f s' % (k_done, cpu_t),
            'qp_solves_per_s': qps / cpu_t,
            'max_abs_fidelity_gap_vs_gpu': float(np.abs(np.array(cpu_fid) - fid[:k_done]).max())}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
