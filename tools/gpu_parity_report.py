# -*- coding: utf-8 -*-
''' The literal north_star parity figures of the current build against every reference-generated grid
    fixture under tests/golden (made by tests/golden/make_goldens.py from the unmodified reference), with the
    list of points that exceed their per-entry bound.  Run on the GPU box:

        python tools/gpu_parity_report.py gpurun_out/parity_r02.json
'''
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pysonic_b200 as ps  # noqa: E402
from parity import entry_bound, grid_err, parity_stats, self_noise  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')
out = {'tolerance': 'V and rates within 1e-4 relative (1e-9 absolute); per-entry bound max(1e-4, 5 x the envelope of the '
                    "entry's deviation over the reference's own +-2 / +-4 ulp re-runs)", 'grids': {}}
stems = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, 'c?_*.npz')) if '_ulp_' not in p)
for stem in stems:
    g = np.load(os.path.join(GOLD, stem + '.npz'))
    variants = [np.load(os.path.join(GOLD, stem + t)) for t in ('_ulp_up.npz', '_ulp_dn.npz', '_ulp_up2.npz', '_ulp_dn2.npz', '_ulp_up3.npz', '_ulp_dn3.npz')
                if os.path.isfile(os.path.join(GOLD, stem + t))]
    if len(variants) < 2:
        continue
    keys = [str(k) for k in g['keys']]
    pn = ps.getPointNeuron(str(g['neuron']))
    lkp, info = ps.computeAStimLookup(pn, g['a'], g['f'], g['A'], g['fs'], g['Q'], return_info=True, loglevel=10)
    st = parity_stats(lkp.tables, info['ncycles'], g, variants, keys)
    st['reference_reruns'] = len(variants)
    err = grid_err(lkp.tables, g, keys)
    env = self_noise(g, variants, keys)
    viol = np.argwhere((err > entry_bound(env)).any(axis=-1))
    st['violations'] = [{'a_nm': float(g['a'][i] * 1e9), 'f_kHz': float(g['f'][j] * 1e-3), 'A_kPa': float(g['A'][k] * 1e-3),
                         'Q_nCcm2': float(g['Q'][l] * 1e5), 'err': float(err[i, j, k, l].max()), 'self': float(env[i, j, k, l].max()),
                         'ncycles': int(info['ncycles'][i, j, k, l]), 'ncycles_ref': int(g['ncycles'][i, j, k, l]),
                         'ncycles_reruns': [int(v['ncycles'][i, j, k, l]) for v in variants]}
                        for i, j, k, l in viol[:60]]
    out['grids'][stem] = st
    print(f"{stem:24s} points {st['points']:6d}  within 1e-4: {st['frac_points_within_1e-4']:.4f} (reference itself {st['reference_self_frac_points_within_1e-4']:.4f})"
          f"  strict coverage {st['strict_1e-4_coverage']:.3f}  bound violations {st['points_violating_entry_bound']:3d}"
          f"  ncycles identical {st['ncycles_identical']:.4f} (self {st['reference_self_ncycles_identical']:.4f}; A>=10kPa {st['ncycles_identical_A_ge_10kPa']})", flush=True)
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'parity_r02.json')
with open(path, 'w') as fh:
    json.dump(out, fh, indent=1)
