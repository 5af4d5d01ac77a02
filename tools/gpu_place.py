import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
from pysonic_b200 import _lib
w = bench.workload('c2'); pn = ps.getPointNeuron('RS')
bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
ia, f, A, Q = bench.flatten(w)
plan = _lib.Plan(0, bls, pn.neuron_id, len(pn.rates), ia, f, A, Q, w['fs'])
for rep in range(4):
    t0 = time.perf_counter(); plan.launch(); plan.sync(); print('launch', rep, '%.3f s' % (time.perf_counter() - t0), flush=True)
