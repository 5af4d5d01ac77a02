# tick latency of lone lanes for n chains spread over the device (n <= warps: one chain per warp), in the
# register-resident run (SONIC_NESTED unset) or the staged tick (SONIC_NESTED=0)
import json, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps
from pysonic_b200 import _lib
pn = ps.getPointNeuron('RS')
bls32 = [ps.NeuronalBilayerSonophore(32e-9, pn).abi_params()]
res = {'nested_env': os.environ.get('SONIC_NESTED', '')}
for n in (148, 296, 592, 1184):
    A = np.full(n, 600e3); f = np.full(n, 500e3); Q = np.linspace(-107e-5, 50e-5, n) + 1.2345e-7
    plan = _lib.Plan(0, bls32, pn.neuron_id, 8, np.zeros(n, np.int32), f, A, Q, np.array([1.0]))
    plan.launch(); plan.sync(); plan.launch(); plan.sync()
    out, ncyc, st, tp, nrhs = plan.fetch()
    r = tp / nrhs * 1e6
    res[f'n{n}'] = {'us_per_rhs_longest': float(r[int(np.argmax(tp))]), 'median': float(np.median(r)), 'max': float(r.max()), 'ms': plan.stats()['ms_integrate']}
    plan.destroy()
print(json.dumps(res))
