import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
w = bench.workload('c2'); pn = ps.getPointNeuron('RS')
# throughput-bound slice of C2: 500 kHz and above only (no long chains), every amplitude and charge
t0 = time.perf_counter()
lkp, info = ps.computeAStimLookup(pn, w['a'], w['f'][2:], w['A'], np.array([1.0]), w['Q'], return_info=True, loglevel=10)
print('points', info['ncycles'].size, 'wall %.2f' % (time.perf_counter() - t0), info['stats'])
