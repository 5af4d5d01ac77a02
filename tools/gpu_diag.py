# -*- coding: utf-8 -*-
''' GPU box: per-point timing diagnostics of the integrator on a workload (default C2):
    saves tpoint / nrhs / ncycles / status per point to gpurun_out/diag_<tag>.npz. '''
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pysonic_b200 as ps  # noqa: E402
from pysonic_b200 import _lib  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else 'c2'
wl = sys.argv[2] if len(sys.argv) > 2 else 'c2'
if wl in ('c1', 'c2'):
    w = bench.workload(wl)
else:
    # any neuron name: the default 4-D grid of run_lookups.py restricted to a = 32 nm (BASELINE C4)
    w = bench.workload('c2')
    pn_ = ps.getPointNeuron(wl)
    Qmin, Qmax = pn_.Qbounds
    w.update(neuron=wl, a=np.array([32e-9]), Q=np.arange(Qmin, Qmax + 1e-5, 1e-5))
pn = ps.getPointNeuron(w['neuron'])
bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
ia, f, A, Q = bench.flatten(w)
plan = _lib.Plan(0, bls, pn.neuron_id, len(pn.rates), ia, f, A, Q, w['fs'])
for rep in range(2):
    t0 = time.perf_counter()
    plan.launch()
    plan.sync()
    print('launch', rep, 'wall %.3f s' % (time.perf_counter() - t0))
out, ncyc, status, tpoint, nrhs = plan.fetch()
st = plan.stats()
print(st)
print('status counts', np.unique(status, return_counts=True))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', f'diag_{tag}.npz'), ia=ia, f=f, A=A, Q=Q, ncyc=ncyc,
                    status=status, tpoint=tpoint, nrhs=nrhs, V=out[0, :, 0])
