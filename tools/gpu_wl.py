# integrator time of a few workloads for the current build / environment: C2, one rank's share of C2 under 2- and 8-way
# sharding, FHnode (C4), TC (C5), the on-node 8058-point grid; with a hash of the integrator's results
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
from pysonic_b200 import _lib
from pysonic_b200.parallel import predicted_log_cost, shard_indices, trajectory_groups
which = sys.argv[2].split(',') if len(sys.argv) > 2 else ['c2', 'shard2', 'shard8', 'FHnode', 'TC', 'node']
res = {'tag': sys.argv[1] if len(sys.argv) > 1 else ''}
def run(name, pn, a, ia, f, A, Q, reps=2):
    bls = [ps.NeuronalBilayerSonophore(float(x), pn).abi_params() for x in a]
    plan = _lib.Plan(0, bls, pn.neuron_id, len(pn.rates), ia, f, A, Q, np.array([1.0]))
    ms = []
    for _ in range(reps):
        plan.launch(); plan.sync(); ms.append(round(plan.stats()['ms_integrate'], 1))
    out, ncyc, st, tp, nrhs = plan.fetch()
    res[name] = {'ms': ms, 'longest_s': round(float(tp.max()), 3), 'hash': hashlib.sha256(ncyc.tobytes() + nrhs.tobytes() + st.tobytes()).hexdigest()[:10]}
    plan.destroy()
pn = ps.getPointNeuron('RS')
w = bench.workload('c2'); ia, f, A, Q = bench.flatten(w)
if 'c2' in which: run('c2', pn, w['a'], ia, f, A, Q, 3)
cost = predicted_log_cost(w['a'][ia], f, A, Q); grp = trajectory_groups(ia, f, A, Q)
for world in (2, 8):
    if f'shard{world}' in which:
        idx = shard_indices(cost, 0, world, grp)
        run(f'shard{world}', pn, w['a'], ia[idx], f[idx], A[idx], Q[idx])
A51 = np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3
F7 = np.array([20., 100., 500., 1e3, 2e3, 3e3, 4e3]) * 1e3
def grid(pn, fs_, As, qstep):
    Qmin, Qmax = pn.Qbounds
    Qs = np.arange(Qmin, Qmax + qstep, qstep)
    F, AA, QQ = np.meshgrid(fs_, As, Qs, indexing='ij')
    return np.zeros(F.size, np.int32), F.ravel(), AA.ravel(), QQ.ravel()
if 'FHnode' in which:
    pn2 = ps.getPointNeuron('FHnode'); run('FHnode', pn2, [32e-9], *grid(pn2, F7, A51, 1e-5), reps=1)
if 'TC' in which:
    pn3 = ps.getPointNeuron('TC'); run('TC', pn3, [32e-9], *grid(pn3, [20e3, 500e3, 4e6], np.logspace(np.log10(50), np.log10(600), 26) * 1e3, 5e-6), reps=1)
if 'node' in which:
    run('node', pn, [32e-9], *grid(pn, [500e3], A51, 1e-5), reps=2)
print(json.dumps(res))
