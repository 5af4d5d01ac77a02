# -*- coding: utf-8 -*-
''' Golden `refs` / table shapes of the reference's computeAStimLookup (scripts/run_lookups.py:22-175)
    for the grid-shaping rules (overtone down-sampling :64-79, test reduction :82-83, overtone axes
    :105-128).  The reference's Batch is replaced by a stub that returns zeros, so only the front-end
    logic runs.  Build container only.

        python tests/golden/make_lookup_refs.py
'''
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refshim import load_reference  # noqa: E402

load_reference()
import PySONIC.core as core  # noqa: E402
from PySONIC.neurons import getPointNeuron  # noqa: E402

import _refshim  # noqa: E402
_refshim._anymod('PySONIC.parsers')   # argparse front end (pulls the plotting package): not needed here
spec = importlib.util.spec_from_file_location('ref_run_lookups', '/root/reference/scripts/run_lookups.py')
rl = importlib.util.module_from_spec(spec)
spec.loader.exec_module(rl)


class FakeBatch:
    def __init__(self, func, queue):
        self.func, self.queue = func, queue

    def __call__(self, **kwargs):
        out = []
        for item in self.queue:
            args = item[0] if isinstance(item, tuple) else item
            nfs = np.size(args[1])
            out.append(([{'V': 0.0} for _ in range(nfs)], 0.0))
        return out

    createQueue = staticmethod(core.Batch.createQueue)
    printQueue = staticmethod(lambda q, nmax=20: None)


rl.Batch = FakeBatch

pn = getPointNeuron('RS')
Adef = np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3
Qdef = np.arange(pn.Qbounds[0], pn.Qbounds[1] + 1e-5, 1e-5)
cases = {
    'overtones1_default': dict(a=[32e-9], f=[500e3], A=Adef, fs=[1.0], Q=Qdef, novertones=1, test=False),
    'overtones1_test': dict(a=[32e-9], f=[500e3], A=Adef, fs=[1.0], Q=Qdef, novertones=1, test=True),
    'overtones2_small': dict(a=[32e-9], f=[500e3], A=Adef[:4], fs=[1.0], Q=Qdef[:3], novertones=2, test=False),
    'plain_test': dict(a=[16e-9, 32e-9, 64e-9], f=[20e3, 4e6], A=Adef, fs=[1.0], Q=Qdef, novertones=0, test=True),
    'fs_span': dict(a=[32e-9], f=[500e3], A=Adef[::10], fs=list(np.arange(1, 101)[::20] * 1e-2), Q=Qdef[::40],
                    novertones=0, test=False),
}
out = {}
for name, c in cases.items():
    lkp = rl.computeAStimLookup(pn, np.array(c['a']), np.array(c['f']), np.array(c['A']), np.array(c['fs']),
                                np.array(c['Q']), novertones=c['novertones'], test=c['test'])
    out[name] = {'args': {k: (list(map(float, v)) if isinstance(v, (list, np.ndarray)) else v) for k, v in c.items()},
                 'refs': {k: list(map(float, v)) for k, v in lkp.refs.items()},
                 'table_keys': list(lkp.tables.keys()), 'shape': list(lkp.tables['V'].shape)}
    print(name, {k: len(v) for k, v in out[name]['refs'].items()}, out[name]['shape'])
# error behaviour: spanning several radii with overtones
try:
    rl.computeAStimLookup(pn, np.array([16e-9, 32e-9]), np.array([500e3]), Adef, np.array([1.0]), Qdef, novertones=1)
    out['overtones_multi_a_error'] = None
except AssertionError as e:
    out['overtones_multi_a_error'] = str(e)
with open(os.path.join(HERE, 'lookup_refs.json'), 'w') as fh:
    json.dump(out, fh)
