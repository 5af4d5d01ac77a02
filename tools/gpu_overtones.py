# -*- coding: utf-8 -*-
''' GPU box: the charge-overtone lookup of run_lookups.py --novertones 1 for RS at one radius and
    frequency (51 A x 158 Q x 5 AQ1 x 5 phiQ1 = 201 450 ODE points): the throughput-bound regime
    of the integrator.  Writes gpurun_out/overtones_<tag>.json. '''
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pysonic_b200 as ps  # noqa: E402
from pysonic_b200 import _lib  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'x'
w = bench.workload('c2')
pn = ps.getPointNeuron('RS')
peak = _lib.fp64_peak(0)
res = []
for rep in range(2):
    t0 = time.perf_counter()
    lkp, info = ps.computeAStimLookup(pn, np.array([32e-9]), np.array([500e3]), w['A'], np.array([1.0]), w['Q'],
                                      novertones=1, return_info=True, loglevel=10)
    dt = time.perf_counter() - t0
    st = info['stats']
    n_full = st['n_rhs'] - 2 * st['n_jac']
    n_corr = st['n_rhs'] - 3 * st['n_jac'] - st['n_cycles']
    flops = (n_full * bench.W_RHS + st['n_steps'] * bench.W_STEP + n_corr * bench.W_CORR + st['n_jac'] * bench.W_JAC +
             st['n_cycles'] * 999 * bench.W_SAMPLE)
    r = {'rep': rep, 'ode_points': int(info['ncycles'].size), 'dims': list(lkp['V'].shape), 'wall_s': dt,
         'points_per_s': info['ncycles'].size / dt, 'ms_integrate': st['ms_integrate'], 'ms_average': st['ms_average'],
         'n_rhs_lsoda': st['n_rhs'], 'n_full_rhs': n_full, 'n_steps': st['n_steps'],
         'algorithmic_tflops': flops / (st['ms_integrate'] * 1e-3) * 1e-12, 'fp64_peak_tflops': peak,
         'roofline_frac': flops / (st['ms_integrate'] * 1e-3) * 1e-12 / peak,
         'finite': bool(all(np.isfinite(v).all() for v in lkp.tables.values())),
         'status_counts': {int(k): int(v) for k, v in zip(*np.unique(info['status'], return_counts=True))},
         'ncycles_hist': np.bincount(info['ncycles'].ravel()).tolist()}
    print(json.dumps(r), flush=True)
    res.append(r)
with open(os.path.join(ROOT, 'gpurun_out', f'overtones_{tag}.json'), 'w') as fh:
    json.dump(res, fh, indent=1)
