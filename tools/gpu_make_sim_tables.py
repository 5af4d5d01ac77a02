# -*- coding: utf-8 -*-
''' GPU box: engine-built lookup tables on the grids of the reference-built fixtures used for the spike-count
    parity check (tools/spike_parity.py): BASELINE config 1 (RS) and tests/golden/sim_tab_<neuron>.npz.
    Pickles are written to gpurun_out/tables/<neuron>.pkl in the reference's format. '''
import glob
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps  # noqa: E402

gold = os.path.join(ROOT, 'tests', 'golden')
out = os.path.join(ROOT, 'gpurun_out', 'tables')
os.makedirs(out, exist_ok=True)
for path in [os.path.join(gold, 'c1_RS_32nm_500kHz.npz')] + sorted(glob.glob(os.path.join(gold, 'sim_tab_*.npz'))):
    g = np.load(path)
    name = str(g['neuron'])
    lkp = ps.computeAStimLookup(ps.getPointNeuron(name), g['a'], g['f'], g['A'], g['fs'], g['Q'], loglevel=10)
    lkp.toPickle(os.path.join(out, f'{name}.pkl'))
    print(name, lkp, flush=True)
