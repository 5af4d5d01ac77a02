# per-point timing of one rank's share of C2 under `world`-way sharding: the longest chains, their tick time and their
# position in the plan's cost order
import os, sys, numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
from pysonic_b200 import _lib
from pysonic_b200.parallel import predicted_log_cost, shard_indices, trajectory_groups
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w = bench.workload('c2'); pn = ps.getPointNeuron('RS')
bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
ia, f, A, Q = bench.flatten(w)
cost = predicted_log_cost(w['a'][ia], f, A, Q); grp = trajectory_groups(ia, f, A, Q)
idx = shard_indices(cost, 0, world, grp)
plan = _lib.Plan(0, bls, pn.neuron_id, len(pn.rates), ia[idx], f[idx], A[idx], Q[idx], w['fs'])
plan.launch(); plan.sync(); plan.launch(); plan.sync()
out, ncyc, st, tp, nrhs = plan.fetch()
print('integrate ms', plan.stats()['ms_integrate'])
c = cost[idx]
rank = np.empty(len(c), int); rank[np.argsort(-c, kind='stable')] = np.arange(len(c))
o = np.argsort(-tp)[:24]
for i in o:
    print(f'tp {tp[i]:.3f} s  nrhs {nrhs[i]:8d}  {tp[i] / nrhs[i] * 1e6:5.2f} us/rhs  predicted {np.exp(c[i]):9.0f}  cost rank {rank[i]:6d}  f {f[idx][i]:.0f} A {A[idx][i]:.0f} Q {Q[idx][i]:.2e}')
