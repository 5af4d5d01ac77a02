# Workloads off the nodes of the measured cost table (a = 16/32/64 nm, f = 20k...4M): schedule quality when the
# queue order and lane budgets come from the interpolated estimate.  Prints one JSON line per workload.
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import pysonic_b200 as ps

pn = ps.getPointNeuron('RS')
A = np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3
Q = np.arange(pn.Qbounds[0], pn.Qbounds[1] + 1e-5, 1e-5)
out = []
for label, a, f in [('on-node reference: a=32 nm, f=500 kHz', [32e-9], [500e3]),
                    ('off-node: a=50 nm (LJ fit computed), f=700 kHz', [50e-9], [700e3]),
                    ('off-node: a=22.6/45.3 nm, f=50/300/700/1500 kHz', [22.6e-9, 45.3e-9], [50e3, 300e3, 700e3, 1.5e6])]:
    a, f = np.array(a), np.array(f)
    ps.computeAStimLookup(pn, a, f, A, np.array([1.0]), Q, loglevel=10)
    t0 = time.perf_counter()
    lkp, info = ps.computeAStimLookup(pn, a, f, A, np.array([1.0]), Q, loglevel=10, return_info=True)
    dt = time.perf_counter() - t0
    st = info['stats']
    npts = lkp['V'].size
    tp = lkp['tcomp'][..., 0]
    rec = {'workload': label, 'points': npts, 'seconds': dt, 'points_per_s': npts / dt, 'kernel_ms': st['ms_integrate'],
           'n_rhs': st['n_rhs'], 'longest_point_s': float(tp.max()), 'kernel_over_longest_point': st['ms_integrate'] * 1e-3 / float(tp.max()),
           'status_words': sorted(int(x) for x in np.unique(info['status'])), 'finite': bool(all(np.isfinite(v).all() for v in lkp.tables.values()))}
    print(json.dumps(rec), flush=True)
