# per-point timing of the C2 grid for a build / setting: npz with tpoint, nrhs, ncycles
import os, sys, numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
from pysonic_b200 import _lib
tag = sys.argv[1]
w = bench.workload('c2'); pn = ps.getPointNeuron('RS')
bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
ia, f, A, Q = bench.flatten(w)
plan = _lib.Plan(0, bls, pn.neuron_id, 8, ia, f, A, Q, w['fs'])
plan.launch(); plan.sync(); plan.launch(); plan.sync()
out, ncyc, st, tp, nrhs = plan.fetch()
stats = plan.stats()
print(tag, 'integrate ms', stats['ms_integrate'], 'sum tp', tp.sum(), 'max tp', tp.max(), 'mean us/rhs', tp.sum() / nrhs.sum() * 1e6)
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', f'c2diag_{tag}.npz'), tp=tp, nrhs=nrhs, ia=ia, f=f, A=A, Q=Q)
