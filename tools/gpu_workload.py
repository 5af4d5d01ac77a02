# one lookup of a BASELINE config 3-5 shape through the public API (for ncu captures): STN = C3 (fs sweep),
# a fibre = C4 at 32 nm, RE / TC = C5
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps
name = sys.argv[1]
pn = ps.getPointNeuron(name)
A51 = np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3
Qmin, Qmax = pn.Qbounds
if name == 'STN':
    args = ([32e-9], [500e3], A51, np.arange(1, 101) * 1e-2, np.arange(Qmin, Qmax + 1e-5, 1e-5))
elif name in ('RE', 'TC'):
    args = ([32e-9], [20e3, 500e3, 4e6], np.logspace(np.log10(50), np.log10(600), 26) * 1e3, [1.0], np.arange(Qmin, Qmax + 5e-6, 5e-6))
else:
    args = ([32e-9], np.array([20., 100., 500., 1e3, 2e3, 3e3, 4e3]) * 1e3, A51, [1.0], np.arange(Qmin, Qmax + 1e-5, 1e-5))
a, f, A, fs, Q = (np.asarray(x, float) for x in args)
lkp, info = ps.computeAStimLookup(pn, a, f, A, fs, Q, return_info=True, loglevel=10)
print(name, lkp, info['stats'])
