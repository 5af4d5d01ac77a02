#!/usr/bin/env python
# -*- coding: utf-8 -*-
''' Benchmark of the SONIC lookup-generation path (BASELINE.json metric: lookup grid points/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1]
                    [--scaling weak|strong]

    One "step" = one pass of the hot path (initial deflection -> batched integrator with the
    periodic-convergence test -> fused cycle averaging) over one lookup grid.  The workload at
    every N is BASELINE config 2, the full RS 4-D lookup (3 a x 7 f x 51 A x 158 Q = 169 218 ODE
    points, fs = 1: one grid point per ODE point), which fits one GPU.

    Own arm (default): per rank, inputs are uploaded once (sonic_plan_create), then
      * `value`   : K launches of the resident plan, timed with CUDA events on the launching
                    stream, barrier + synchronize on both sides, max over ranks;
      * `e2e`     : K calls of the public API `computeAStimLookup` with host arrays (allocation,
                    host->device copies, kernels, device->host copies inside the timed region);
      * `roofline`: algorithmic FP64 flops of the integrator kernel (SURVEY.md 8(d) weights, the
                    right-hand-side evaluations counted in-kernel) / its CUDA-event duration,
                    against the FP64 FMA peak measured on this device in the same run;
      * `cpu_baseline` (N = 1): the CPU oracle (a port of the reference path around scipy's LSODA,
                    oracle/sonic_oracle.py) on all host cores, on a bounded systematic sample of
                    the same grid.
    Reference arm (--impl reference): the CPU oracle port on all host cores, each step one pass
    over the bounded sample; rank 0 only.

    Under torchrun (N > 1) every rank drives its own GPU; there is no data-path collective.
    --scaling strong (default): ONE RS 4-D table sharded over the ranks (whole trajectory groups dealt
    round-robin in cost order, `computeAStimLookup(mpi=True)` for the end-to-end figure, one final
    all_gather); the replica figure (every rank regenerates a full table) is reported next to it as
    `config.weak`.  --scaling weak makes the replica figure the headline instead.

    Extra keys of the own arm: `parity` (the literal north_star figures of this build, computed live
    against the reference-generated fixtures under tests/golden), `multi_neuron` (the four cortical
    tables through one integrator launch, N = 1).
'''

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'lookup grid points/s (RS full 4-D lookup)'
UNIT = 'points/s'

# ---- algorithmic flop weights, SURVEY.md 8(d): add/sub/mul 1, FMA 2, div 10, exp 50, log 60,
# ---- sin 50, pow 200
W_RHS = 532          # one right-hand-side evaluation (bls.py:681-718)
W_CM = 82            # one capacitance sample (bls.py:334-345)
W_V = 14             # one (sample, fs) membrane potential
W_STEP = 190         # per accepted step: Pascal prediction (45), history update (36), weights (36),
#                      error norm / tests (33), order-selection powers amortised over nq+1 steps (40)
W_CORR = 64          # per corrector iteration: residual (12), 3x3 chord solve (42), norm + rate (10)
W_JAC = 200          # per Jacobian: U and ng difference columns (2 div + 50), Z column (27), norm (33), 3x3 LU (70)
W_SAMPLE = 46        # per output sample: Nordsieck interpolation (30), quotient (10), accumulators (6)


def rates_weight(name):
    ''' Algorithmic flops of one evaluation of all rate constants of a neuron, same weights,
        counted on the rate expressions the device functions are generated from. '''
    from pysonic_b200.neurons import NEURON_SPECS, Gate
    total = 0
    for k in NEURON_SPECS[name]['kin']:
        exprs = list(k.pre) + ([k.xinf, k.tau] if isinstance(k, Gate) else [k.expr])
        for e in exprs:
            nexp = e.count('exp(')
            nvt = e.count('vtrap(')
            total += 50 * nexp + nvt * (50 + 2 * 10 + 2)
            total += 10 * e.count('/') + e.count('*') + e.count('+') + e.count('-')
        if isinstance(k, Gate):
            total += 2 * 10 + 1   # alpha = xinf / tau, beta = (1 - xinf) / tau
    return total


def workload(name):
    ''' Grid definitions of SURVEY.md 8(d) / BASELINE.md 3. '''
    import pysonic_b200 as ps
    pn = ps.getPointNeuron('RS')
    if name == 'c1':
        return dict(neuron='RS', a=np.array([32e-9]), f=np.array([500e3]),
                    A=np.insert(np.logspace(np.log10(0.1), np.log10(600), 19), 0, 0.) * 1e3,
                    Q=np.linspace(-107e-5, 50e-5, 50), fs=np.array([1.0]),
                    label='C1: RS, a=32 nm, f=500 kHz, 20 A x 50 Q, fs=1 (1 000 points)')
    Qmin, Qmax = pn.Qbounds
    return dict(neuron='RS', a=np.array([16e-9, 32e-9, 64e-9]),
                f=np.array([20., 100., 500., 1e3, 2e3, 3e3, 4e3]) * 1e3,
                A=np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3,
                Q=np.arange(Qmin, Qmax + 1e-5, 1e-5), fs=np.array([1.0]),
                label='C2: RS full 4-D lookup, a 16/32/64 nm x 7 f (20 kHz-4 MHz) x 51 A (0-600 kPa) '
                      'x 158 Q, fs=1 (169 218 points)')


def flatten(w):
    ia, f, A, Q = np.meshgrid(np.arange(w['a'].size), w['f'], w['A'], w['Q'], indexing='ij')
    return ia.ravel().astype(np.int32), f.ravel(), A.ravel(), Q.ravel()


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# ---------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_sample(w, cores, budget_s):
    ''' Systematic sample of the cost-sorted grid: every `stride`-th point, so that the sample
        has the cost profile of the whole grid.  Sized for about `budget_s` seconds on `cores`
        processes (mean reference cost 0.68 core-seconds per point on C2, 0.37 on C1). '''
    from pysonic_b200.parallel import predicted_log_cost
    ia, f, A, Q = flatten(w)
    n = ia.size
    mean_cost = 0.68 if n > 1000 else 0.37
    m = int(min(n, max(2 * cores, round(budget_s * cores / mean_cost))))
    order = np.argsort(-predicted_log_cost(w['a'][ia], f, A, Q), kind='stable')
    stride = max(n // m, 1)
    idx = order[stride // 2::stride][:m]
    jobs = [(w['neuron'], float(w['a'][ia[i]]), float(f[i]), float(A[i]), w['fs'], float(Q[i])) for i in idx]
    desc = (f'{len(jobs)} of {n} points: every {stride}th point of the cost-sorted grid '
            f'(systematic sample, all a/f/A/Q strata), {cores} processes')
    return jobs, desc


def cpu_pass(jobs, cores):
    ''' One pass of the oracle over the sample; returns wall seconds. '''
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import multiprocessing as mp
    import sonic_oracle as so
    t0 = time.perf_counter()
    if cores > 1:
        with mp.get_context('fork').Pool(cores) as pool:
            pool.map(so._point, jobs, chunksize=1)
    else:
        for j in jobs:
            so._point(j)
    return time.perf_counter() - t0


def reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    w = workload(args.workload)
    cores = host_cores()
    jobs, desc = cpu_sample(w, cores, min(args.cpu_budget, 12.0))   # K + W passes: keep each short
    for _ in range(args.warmup):
        cpu_pass(jobs, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(jobs, cores)
    dt = time.perf_counter() - t0
    value = len(jobs) * args.steps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': args.scaling or 'strong', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic (deterministic grid from the reference formulas)',
        'config': {'workload': w['label'], 'sample': desc},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class Clocks:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        self.thr.join(timeout=2)
        sm, smax, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(nm)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': smax,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def live_parity(ps, lkp_c2=None, info_c2=None, w_c2=None, device=0):
    ''' The literal north_star figures of this build against the reference-generated fixtures
        (tests/golden, made by tests/golden/make_goldens.py from the unmodified reference): BASELINE
        config 1 in full, and -- when the full RS 4-D table is at hand -- its 10 710 nodes of the dense
        C2 fixture. '''
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from parity import parity_stats
    gold = os.path.join(ROOT, 'tests', 'golden')
    out = {'tolerance': 'V and rates within 1e-4 relative (1e-9 absolute); per-entry bound max(1e-4, 5 x the '
                        "entry's own deviation under +-2 ulp re-runs of the reference)"}

    def variants(stem):
        return [np.load(os.path.join(gold, stem + t))
                for t in ('_ulp_up.npz', '_ulp_dn.npz', '_ulp_up2.npz', '_ulp_dn2.npz', '_ulp_up3.npz', '_ulp_dn3.npz')
                if os.path.isfile(os.path.join(gold, stem + t))]

    g = np.load(os.path.join(gold, 'c1_RS_32nm_500kHz.npz'))
    keys = [str(k) for k in g['keys']]
    # (this rank alone, on its own device: no collective, the other ranks are not here)
    lkp, info = ps.computeAStimLookup(ps.getPointNeuron('RS'), g['a'], g['f'], g['A'], g['fs'], g['Q'],
                                      return_info=True, loglevel=10, device=device, shard=False)
    out['c1_full_1000_points'] = parity_stats(lkp.tables, info['ncycles'], g, variants('c1_RS_32nm_500kHz'), keys)
    big = os.path.join(gold, 'c2_RS_big.npz')
    if lkp_c2 is not None and os.path.isfile(big) and len(variants('c2_RS_big')) >= 2:
        g = np.load(big)
        iQ = [int(np.argmin(np.abs(w_c2['Q'] - x))) for x in g['Q']]
        if np.array_equal(w_c2['Q'][iQ], g['Q']) and np.array_equal(w_c2['A'], g['A']):
            sub = {k: lkp_c2[k][:, :, :, iQ] for k in keys}
            out['c2_dense_10710_points'] = parity_stats(sub, info_c2['ncycles'][:, :, :, iQ], g,
                                                        variants('c2_RS_big'), keys)
    return out


def gpu_arm(args):
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    w = workload(args.workload)
    scaling = args.scaling or 'strong'

    # the CPU baseline leg runs first, before any CUDA context exists in this process (fork)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        jobs, desc = cpu_sample(w, cores, args.cpu_budget)
        dt = cpu_pass(jobs, cores)
        cpu = {'value': len(jobs) / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc,
               'seconds': dt}

    import torch
    import torch.distributed as dist
    import pysonic_b200 as ps
    from pysonic_b200 import _lib
    from pysonic_b200.parallel import predicted_log_cost, shard_indices, trajectory_groups

    _lib.load()
    if _lib.device_count() < 1:
        raise SystemExit('bench.py: no CUDA device; the engine has no CPU path')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    pn = ps.getPointNeuron(w['neuron'])
    nrates = len(pn.rates)
    bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
    ia_all, f_all, A_all, Q_all = flatten(w)
    n_grid = ia_all.size
    peak = _lib.fp64_peak(local_rank)          # FP64 FMA peak of this device, TFLOP/s
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def measure(mode, steps, warmup):
        ''' Resident plan of this rank's share of the work (the whole grid, or its shard of it) on a
            torch stream; W + K launches, CUDA events around the K timed ones, max over ranks. '''
        ia, f, A, Q = ia_all, f_all, A_all, Q_all
        if mode == 'strong' and world > 1:
            idx = shard_indices(predicted_log_cost(w['a'][ia], f, A, Q), rank, world,
                                trajectory_groups(ia, f, A, Q))
            ia, f, A, Q = ia[idx], f[idx], A[idx], Q[idx]
        plan = _lib.Plan(local_rank, bls, pn.neuron_id, nrates, ia, f, A, Q, w['fs'])
        plan.set_stream(stream.cuda_stream)

        def step():
            with torch.cuda.stream(stream):
                flush.fill_(1)
            plan.launch()

        for _ in range(warmup):
            step()
        barrier()
        clocks = Clocks(local_rank)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        for _ in range(steps):
            step()
        ev[1].record(stream)
        barrier()
        clk = clocks.stop()
        ms_region = max_over_ranks(ev[0].elapsed_time(ev[1]))
        st = plan.stats()                      # counters + per-kernel event times of the last launch
        plan.destroy()
        n_job = n_grid * world if (mode == 'weak' and world > 1) else n_grid
        return {'ms_per_step': ms_region / steps, 'value': n_job / (ms_region / steps * 1e-3), 'stats': st,
                'clocks': clk, 'n_job': n_job, 'n_local': ia.size}

    head = measure(scaling, args.steps, args.warmup)
    other = None
    if world > 1:
        other = measure('weak' if scaling == 'strong' else 'strong', max(1, args.steps - 1), 1)
    st = head['stats']
    ms_int = st['ms_integrate']
    n_job, n_local = head['n_job'], head['n_local']

    # ---- roofline of the dominant kernel (the integrator), this rank ----
    n_corr = st['n_rhs'] - 3 * st['n_jac'] - st['n_cycles']
    # n_rhs counts like LSODA (3 per finite-difference Jacobian); the kernel evaluates one full
    # right-hand side per Jacobian (Z column) and forms the U and ng columns from exact
    # differences (inside W_JAC), so full evaluations = n_rhs - 2 n_jac
    n_full = st['n_rhs'] - 2 * st['n_jac']
    flops_int = (n_full * W_RHS + st['n_steps'] * W_STEP + n_corr * W_CORR + st['n_jac'] * W_JAC +
                 st['n_cycles'] * 999 * W_SAMPLE)
    achieved = flops_int / (ms_int * 1e-3) * 1e-12
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.isfile(tpath):
        with open(tpath) as fh:
            traffic = json.load(fh).get(args.workload, {}).get('integrate_dram_bytes_per_launch')
    roofline = {
        'kernel': 'sonic_integrate_kernel', 'bound': 'fp64', 'achieved': achieved, 'peak': peak,
        'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic,
        'peak_source': 'FP64 FMA microbenchmark (sonic_fp64_peak) on this device in this run; '
                       'MEASURED_PEAKS.json has no FP64 entry',
        'algorithmic_flops_per_launch': flops_int, 'kernel_ms': ms_int,
        'kernel_share_of_step': ms_int / (st['ms_z0'] + st['ms_integrate'] + st['ms_average']),
        'rhs_evaluations': n_full, 'rhs_evaluations_lsoda_count': st['n_rhs'], 'steps': st['n_steps'], 'jacobians': st['n_jac'],
        'cycles': st['n_cycles'],
        'hbm_note': 'algorithmic HBM bytes/point ~ 8.2 kB (one 1000-sample cycle profile written and '
                    're-read) + 150 B of inputs/outputs against >= 1e7 flops: not HBM-bound',
    }

    # ---- end to end through the public API, host buffers ----
    barrier()
    h2d = n_local * (4 + 3 * 8 + 4) + w['fs'].size * 8 + len(bls) * 64
    d2h = n_local * ((1 + nrates) * w['fs'].size * 8 + 4 + 4 + 8)

    def e2e_step():
        if scaling == 'strong' and world > 1:
            return ps.computeAStimLookup(pn, w['a'], w['f'], w['A'], w['fs'], w['Q'], mpi=True,
                                         loglevel=10, return_info=True)
        return ps.computeAStimLookup(pn, w['a'], w['f'], w['A'], w['fs'], w['Q'], loglevel=10,
                                     device=local_rank, shard=False, return_info=True)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        lkp, info = e2e_step()
    torch.cuda.synchronize()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = n_job * args.steps / dt_e2e
    finite = all(np.isfinite(v).all() for v in lkp.tables.values())

    # ---- extras (rank 0, outside every timed region) ----
    parity = multi = None
    if rank == 0 and not args.no_extras:
        parity = live_parity(ps, lkp if args.workload == 'c2' else None, info, w, device=local_rank)
    if world == 1 and not args.no_extras and args.workload == 'c2':
        names = ['RS', 'FS', 'LTS', 'IB']
        pns = [ps.getPointNeuron(x) for x in names]
        Qs = [np.arange(p_.Qbounds[0], p_.Qbounds[1] + 1e-5, 1e-5) for p_ in pns]
        ps.computeAStimLookups(pns, w['a'], w['f'], w['A'], w['fs'], Qs, loglevel=10, device=local_rank)
        t0 = time.perf_counter()
        lk4, i4 = ps.computeAStimLookups(pns, w['a'], w['f'], w['A'], w['fs'], Qs, loglevel=10,
                                         device=local_rank, return_info=True)
        dt4 = time.perf_counter() - t0
        npts = sum(int(np.prod(x['V'].shape)) for x in lk4)
        same = all(np.array_equal(lk4[0][k], lkp[k]) for k in ['V'] + pn.rates)
        multi = {'neurons': names, 'grid_points': npts, 'seconds_e2e': dt4, 'points_per_s_e2e': npts / dt4,
                 'kernel_ms_integrate': i4['stats']['ms_integrate'],
                 'time_vs_one_table': dt4 / (dt_e2e / args.steps),
                 'rs_table_bit_identical_to_single_neuron_run': bool(same),
                 'api': 'pysonic_b200.computeAStimLookups -> sonic_lookup_run_multi (one integrator launch, '
                        'one averaging launch per neuron)'}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    par = {'weak': f'{world} ranks, one full table per rank (replicas), no collective',
           'strong': f'{world} ranks, ONE table sharded (trajectory groups dealt round-robin in cost order), '
                     'no data-path collective, one final all_gather of the tables'}
    config = {
        'workload': w['label'], 'ode_points_per_step': n_job, 'grid_points_per_step': n_job * w['fs'].size,
        'parallelism': '1 GPU' if world == 1 else par[scaling],
        'l2': 'explicit 256 MB flush write before every step (and 1.35 GB of cycle profiles per step)',
    }
    if other is not None:
        config['weak' if scaling == 'strong' else 'strong'] = {
            'value': other['value'], 'unit': UNIT, 'ms_per_step': other['ms_per_step'],
            'parallelism': par['weak' if scaling == 'strong' else 'strong']}
    line = {
        'metric': METRIC, 'value': head['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True,
        'scaling': scaling, 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic (deterministic grid from the reference formulas, no RNG)',
        'config': config,
        'roofline': roofline,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': dt_e2e / args.steps * 1e3, 'api': 'pysonic_b200.computeAStimLookup -> '
                'sonic_lookup_run / sonic_points_run (C ABI, host buffers)', 'tables_finite': bool(finite)},
        'gpu_launches': 4 * args.steps,
        'kernel_ms': {'z0': st['ms_z0'], 'integrate': st['ms_integrate'], 'average': st['ms_average']},
        'clocks': head['clocks'],
    }
    if parity is not None:
        line['parity'] = parity
    if multi is not None:
        line['multi_neuron'] = multi
    if cpu is not None:
        line['cpu_baseline'] = cpu
    print(json.dumps(line), flush=True)


def main():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=3)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='own', choices=['own', 'reference'])
    p.add_argument('--workload', default='c2', choices=['c1', 'c2'])
    p.add_argument('--scaling', default=None, choices=['weak', 'strong'],
                   help='N > 1: strong (default) = one table sharded over the ranks, weak = one table per rank')
    p.add_argument('--no-extras', action='store_true', help='skip the parity and multi-neuron legs')
    p.add_argument('--cpu-budget', type=float, default=20.0, help='seconds of CPU work per sample pass')
    p.add_argument('--no-cpu-baseline', action='store_true')
    args = p.parse_args()
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
