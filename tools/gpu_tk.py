# tick time of a staged warp with k busy lanes: every warp of the device is given k long chains (SONIC_SCHED_FORCE_CAP=k,
# no lone SMs), all of the same kind (32 nm, 500 kHz, 600 kPa, spread charges); us per LSODA right-hand side of a lane
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
if len(sys.argv) > 1:
    k = int(sys.argv[1])
    sys.path.insert(0, ROOT)
    import pysonic_b200 as ps
    from pysonic_b200 import _lib
    pn = ps.getPointNeuron('RS')
    bls32 = [ps.NeuronalBilayerSonophore(32e-9, pn).abi_params()]
    n = 1184 * k
    A = np.full(n, 600e3); f = np.full(n, 500e3); Q = np.linspace(-107e-5, 50e-5, n) + 1.2345e-7
    plan = _lib.Plan(0, bls32, pn.neuron_id, len(pn.rates), np.zeros(n, np.int32), f, A, Q, np.array([1.0]))
    plan.launch(); plan.sync()
    out, ncyc, st, tp, nrhs = plan.fetch()
    r = tp / nrhs * 1e6
    print(json.dumps({'k': k, 'median_us_per_rhs': float(np.median(r)), 'p90': float(np.percentile(r, 90)), 'ms': plan.stats()['ms_integrate']}))
else:
    for k in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
        env = dict(os.environ, SONIC_SCHED_FORCE_CAP=str(k), SONIC_SCHED_TIER1_SMS='0', SONIC_SCHED_TIER2_SMS='0', SONIC_NESTED='0')
        print(subprocess.run([sys.executable, __file__, str(k)], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1], flush=True)
