# -*- coding: utf-8 -*-
''' GPU box: run BASELINE configs 3-5 at full size through the public API and report
    size-independent properties + timing (gpurun_out/configs_<tag>.json). '''
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'x'
A51 = np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3
F7 = np.array([20., 100., 500., 1e3, 2e3, 3e3, 4e3]) * 1e3


def qdefault(pn, step=1e-5):
    Qmin, Qmax = pn.Qbounds
    return np.arange(Qmin, Qmax + step, step)


cases = [('C3 STN fs sweep', 'STN', [32e-9], [500e3], A51, None, np.arange(1, 101) * 1e-2, 1e-5)]
cases += [(f'C4 {n}', n, [32e-9], F7, A51, None, [1.0], 1e-5) for n in ('FHnode', 'SWnode', 'MRGnode', 'SUseg')]
cases += [(f'C5 {n}', n, [32e-9], [20e3, 500e3, 4e6], np.logspace(np.log10(50), np.log10(600), 26) * 1e3, None, [1.0], 5e-6)
          for n in ('RE', 'TC')]
res = []
for label, name, a, f, A, Q, fs, qstep in cases:
    pn = ps.getPointNeuron(name)
    Q = qdefault(pn, qstep) if Q is None else Q
    a, f, A, fs = (np.asarray(x, float) for x in (a, f, A, fs))
    t0 = time.perf_counter()
    lkp, info = ps.computeAStimLookup(pn, a, f, A, fs, Q, return_info=True, loglevel=10)
    dt = time.perf_counter() - t0
    nc, st = info['ncycles'], info['status']
    finite = all(np.isfinite(v).all() for v in lkp.tables.values())
    sgn = np.sign(np.where(np.abs(Q) < 1e-12, 0., Q))[None, None, None, :, None]
    r = {'case': label, 'ode_points': int(nc.size), 'entries': int(lkp['V'].size), 'tables': len(lkp.tables) - 1,
         'wall_s': dt, 'ode_points_per_s': nc.size / dt, 'entries_per_s': lkp['V'].size / dt,
         'kernel_ms': {k: info['stats'][k] for k in ('ms_z0', 'ms_integrate', 'ms_average')},
         'finite': bool(finite), 'status_counts': {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))},
         'ncycles_min_max': [int(nc.min()), int(nc.max())], 'max_rhs_per_point': None,
         'V_sign_ok': bool(np.all(np.sign(lkp['V']) * sgn >= 0)),
         'rates_nonnegative': bool(all(np.all(lkp[k] >= 0) for k in pn.rates))}
    print(json.dumps(r), flush=True)
    res.append(r)
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
with open(os.path.join(ROOT, 'gpurun_out', f'configs_{tag}.json'), 'w') as fh:
    json.dump(res, fh, indent=1)
