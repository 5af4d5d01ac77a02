# -*- coding: utf-8 -*-
''' GPU box: robustness sweep of the C ABI outside the BASELINE grids (random radii from the
    parameter table, f 10 kHz - 10 MHz, A up to 2 MPa, |Q| up to 300 nC/cm2): every point must come
    back with finite tables or a failure status, and the call must return. '''
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps  # noqa: E402

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = 256
for name, radii in (('RS', [16e-9, 32e-9, 64e-9, 100e-9]), ('SWnode', [32e-9])):
    pn = ps.getPointNeuron(name)
    for a in radii:
        try:
            nbls = ps.NeuronalBilayerSonophore(a, pn)
        except ValueError:
            continue
        f = 10 ** rng.uniform(4, 7, n)
        A = np.where(rng.random(n) < 0.1, 0., 10 ** rng.uniform(2, np.log10(2e6), n))
        Q = rng.uniform(-300e-5, 300e-5, n)
        t0 = time.perf_counter()
        out, ncyc, status, tp, nrhs, st = nbls.effvars_batch(f, A, Q, [0.5, 1.0])
        dt = time.perf_counter() - t0
        ok = (status & ~np.uint32(3)) == 0
        finite = np.isfinite(out[:, ok]).all()
        nanbad = np.isnan(out[:, ~ok]).all() if (~ok).any() else True
        print(f'{name} a={a*1e9:.0f} nm: {dt:.2f} s, status counts '
              f'{dict(zip(*[x.tolist() for x in np.unique(status, return_counts=True)]))}, max rhs {nrhs.max()}, '
              f'finite(ok)={finite}, nan(failed)={nanbad}', flush=True)
        assert finite and nanbad
print('fuzz ok')
