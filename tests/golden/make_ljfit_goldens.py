# -*- coding: utf-8 -*-
''' Golden vectors for the intermolecular-pressure parameters: runs the UNMODIFIED
    reference's BilayerSonophore.computePMparams / findDeltaEq / LJfitPMavg (bls.py:410-506) on a writable
    copy of the package (the reference caches its results in PySONIC/core/bls_lookups.json, and
    /root/reference is read-only) for (radius, resting charge) pairs that are absent from that cache,
    plus a few samples of PMavg itself (bls.py:390-404).  Build container only.

        python tests/golden/make_ljfit_goldens.py
'''
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refshim  # noqa: E402

tmp = tempfile.mkdtemp(prefix='pysonic_copy_')
shutil.copytree('/root/reference/PySONIC', os.path.join(tmp, 'PySONIC'))
_refshim.REF_ROOT = tmp
_refshim.load_reference()
from PySONIC.core import BilayerSonophore  # noqa: E402

PAIRS = [(20e-9, -71.9e-5), (45e-9, -71.9e-5), (50e-9, -89.5e-5), (100e-9, -58e-5), (12.5e-9, -140e-5),
         (32e-9, -30e-5), (27e-9, 0.0)]
out = {'fits': [], 'pmavg': []}
for a, Qm0 in PAIRS:
    b = BilayerSonophore(a, 1e-2, Qm0)
    rec = {'a': a, 'Qm0': Qm0, 'Delta': b.Delta, **b.LJ_approx}
    if Qm0 != 0.0:
        D_eq, Pnet = b.findDeltaEq(Qm0)
        rec['Delta_eq'] = D_eq
        rec['Pnet_eq'] = Pnet
    # the reference's own reproducibility: the same fit with the quadrature values perturbed at
    # rounding level (relative 1e-13), twice; the parameters sit in a flat valley for some radii
    orig = b.v_PMavg
    noise = 0.0
    for seed in (1, 2):
        rng = np.random.default_rng(seed)
        b.v_PMavg = lambda Z, R, S: orig(Z, R, S) * (1 + 1e-13 * rng.standard_normal(len(Z)))
        LJ2, _, _ = b.LJfitPMavg()
        noise = max(noise, max(abs(LJ2[k] - b.LJ_approx[k]) / abs(b.LJ_approx[k]) for k in LJ2))
    b.v_PMavg = orig
    rec['self_noise'] = float(noise)
    rec = {k: float(v) for k, v in rec.items()}
    out['fits'].append(rec)
    print(rec, flush=True)
    Zs = np.concatenate([np.linspace(b.Zmin * 0.9, -1e-11, 7), np.linspace(1e-11, 2 * a, 9)])
    for Z in Zs:
        out['pmavg'].append({'a': a, 'Delta': b.Delta, 'Z': float(Z),
                             'PMavg': float(b.PMavg(Z, b.curvrad(Z), b.surface(Z)))})
with open(os.path.join(HERE, 'ljfit.json'), 'w') as fh:
    json.dump(out, fh, indent=1)
shutil.rmtree(tmp)
