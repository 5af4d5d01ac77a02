# lone-lane tick latency vs warps per SM: C1 rows with the longest chains, n points -> n warps
import os, sys, time
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench, pysonic_b200 as ps
from pysonic_b200 import _lib
w = bench.workload('c1'); pn = ps.getPointNeuron('RS')
bls = [ps.NeuronalBilayerSonophore(32e-9, pn).abi_params()]
for n in (148, 296, 592, 1000):
    A = np.full(n, 600e3); f = np.full(n, 500e3); Q = np.linspace(-107e-5, 50e-5, n) + 1.2345e-7
    plan = _lib.Plan(0, bls, pn.neuron_id, 8, np.zeros(n, np.int32), f, A, Q, np.array([1.0]))
    plan.launch(); plan.sync()
    t0 = time.perf_counter(); plan.launch(); plan.sync(); dt = time.perf_counter() - t0
    out, ncyc, st, tp, nrhs = plan.fetch()
    print(f'n={n:5d} warps/SM={n/148:4.1f}  kernel {dt*1e3:7.1f} ms  max nrhs {nrhs.max()}  us per lsoda-rhs (slowest point) {tp.max()/nrhs[np.argmax(tp)]*1e6:.3f}  mean {np.mean(tp/nrhs)*1e6:.3f}')
    plan.destroy()
