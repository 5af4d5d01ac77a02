# integrator time of the C2 grid (3 launches) for the current build / environment
import os, sys, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
from pysonic_b200 import _lib
w = bench.workload('c2'); pn = ps.getPointNeuron('RS')
bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
ia, f, A, Q = bench.flatten(w)
plan = _lib.Plan(0, bls, pn.neuron_id, 8, ia, f, A, Q, w['fs'])
ms = []
for _ in range(3):
    plan.launch(); plan.sync(); ms.append(round(plan.stats()['ms_integrate'], 1))
print(json.dumps({'tag': sys.argv[1] if len(sys.argv) > 1 else '', 'c2_ms': ms}))
