# Kernel probe for a build of the library (PYSONIC_B200_LIB=<.so> python tools/gpu_kprobe.py [tag]):
# lone-lane tick latency, C1 and C2 integrator times, and hashes of the results (bit-identity between builds).
import hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
from pysonic_b200 import _lib

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(_lib.LIB_PATH)
pn = ps.getPointNeuron('RS')
res = {'tag': tag, 'lib': os.path.basename(_lib.LIB_PATH)}
# lone lanes: 148 long chains of C1, one per SM
bls32 = [ps.NeuronalBilayerSonophore(32e-9, pn).abi_params()]
for n in (148, 592):
    A = np.full(n, 600e3); f = np.full(n, 500e3); Q = np.linspace(-107e-5, 50e-5, n) + 1.2345e-7
    plan = _lib.Plan(0, bls32, pn.neuron_id, 8, np.zeros(n, np.int32), f, A, Q, np.array([1.0]))
    plan.launch(); plan.sync(); plan.launch(); plan.sync()
    out, ncyc, st, tp, nrhs = plan.fetch()
    i = int(np.argmax(tp))
    res[f'lone_us_per_rhs_n{n}'] = float(tp[i] / nrhs[i] * 1e6)
    res[f'lone_hash_n{n}'] = hashlib.sha256(out.tobytes() + nrhs.tobytes()).hexdigest()[:12]
    plan.destroy()
# the heaviest C2 chain alone (16 nm, 20 kHz)
bls16 = [ps.NeuronalBilayerSonophore(16e-9, pn).abi_params()]
plan = _lib.Plan(0, bls16, pn.neuron_id, 8, np.zeros(4, np.int32), np.full(4, 20e3), np.array([3e3, 5e3, 8e3, 1e5]),
                 np.full(4, -106e-5), np.array([1.0]))
plan.launch(); plan.sync()
out, ncyc, st, tp, nrhs = plan.fetch()
res['heavy_chain'] = {'nrhs': nrhs.tolist(), 'seconds': tp.tolist(), 'us_per_rhs': (tp / nrhs * 1e6).tolist()}
plan.destroy()
for wl in ('c1', 'c2'):
    w = bench.workload(wl)
    bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
    ia, f, A, Q = bench.flatten(w)
    plan = _lib.Plan(0, bls, pn.neuron_id, 8, ia, f, A, Q, w['fs'])
    ms = []
    for _ in range(3):
        plan.launch(); plan.sync(); ms.append(plan.stats()['ms_integrate'])
    out, ncyc, st, tp, nrhs = plan.fetch()
    res[f'{wl}_ms'] = ms
    res[f'{wl}_hash'] = hashlib.sha256(out.tobytes() + ncyc.tobytes() + nrhs.tobytes()).hexdigest()[:12]
    plan.destroy()
# one rank's share of the C2 table under 2 / 8-way strong scaling (trajectory groups dealt round-robin)
from pysonic_b200.parallel import predicted_log_cost, shard_indices, trajectory_groups
w = bench.workload('c2')
bls = [ps.NeuronalBilayerSonophore(float(a), pn).abi_params() for a in w['a']]
ia, f, A, Q = bench.flatten(w)
cost = predicted_log_cost(w['a'][ia], f, A, Q); grp = trajectory_groups(ia, f, A, Q)
for world in (2, 8):
    idx = shard_indices(cost, 0, world, grp)
    plan = _lib.Plan(0, bls, pn.neuron_id, 8, ia[idx], f[idx], A[idx], Q[idx], w['fs'])
    ms = []
    for _ in range(2):
        plan.launch(); plan.sync(); ms.append(plan.stats()['ms_integrate'])
    res[f'c2_shard_1_of_{world}_ms'] = ms
    plan.destroy()
print(json.dumps(res), flush=True)
