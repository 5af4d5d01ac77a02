# -*- coding: utf-8 -*-
''' SURVEY 7, hard part 2: the reference's own integration error is about the parity tolerance (odeint at
    rtol = atol = 1.49e-8, where atol exceeds |Z| and |ng|: only U is error-controlled).  This diagnostic puts
    three solutions of the same points side by side:

        reference   the reference-generated fixture (tests/golden/c1_RS_32nm_500kHz.npz: BASELINE config 1)
        engine      the integrator of the engine (CPU build of the lane state machine, tests/hostsim: the same
                    arithmetic as the CUDA kernel up to the elementary functions) + the oracle's averaging
        tight       the oracle with rtol = 1e-12 and per-component atol = 1e-12 x (1 m/s, 1 nm, ng0): a converged
                    solution of the same cycle-by-cycle problem (same sampling, same convergence test)

    and reports, over the effective variables (V and every rate) of every sampled point, how far the reference and the
    engine sit from the converged solution and from each other.  CPU only (test infrastructure: imports oracle/ and
    tests/).

        python tools/tight_oracle_diag.py [stride] [out.json]
'''
import json
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, ROOT)
import sonic_oracle as so  # noqa: E402


def rel(x, r):
    d = abs(x - r)
    return 0.0 if d < 1e-9 else d / max(abs(r), 1e-300)


def tight_point(job):
    a, f, A, Q = job
    b = so.get_bls('RS', a)
    atol = np.array([1e-12, 1e-12 * 1e-9, 1e-12 * b.ng0])
    return _tight(b, f, A, Q, atol)


def _tight(b, f, A, Q, atol):
    t, y, ncyc = so.sim_cycles(b, f, A, float(Q), rtol=1e-12, atol=atol)
    z = y[-so.NPC_DENSE:, 1]
    Cm = so.v_capacitance(b, z)
    Vm = Q / Cm * 1e3
    ev = {'V': float(np.mean(Vm))}
    ev.update({k: float(v) for k, v in so.eff_rates('RS', Vm).items()})
    return ev, int(ncyc)


def main():
    stride = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    out = sys.argv[2] if len(sys.argv) > 2 else None
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'c1_RS_32nm_500kHz.npz'))
    keys = [str(k) for k in g['keys']]
    a, f = float(g['a'][0]), float(g['f'][0])
    pts = [(iA, iQ) for iA in range(len(g['A'])) for iQ in range(len(g['Q']))][::stride]
    jobs = [(a, f, float(g['A'][iA]), float(g['Q'][iQ])) for iA, iQ in pts]
    with mp.Pool(min(mp.cpu_count(), 16)) as pool:
        tight = pool.map(tight_point, jobs, chunksize=4)
    # engine integrator on the CPU
    from conftest import HostSim
    import __graft_entry__ as ge
    ge.build()
    hs = HostSim(ge.HOSTSIM_LIB)
    hs.set_driver(2)
    b = so.get_bls('RS', a)
    e_ref, e_eng, e_re, cyc = [], [], [], []
    for (iA, iQ), job, (tev, tcyc) in zip(pts, jobs, tight):
        h = hs.point(b, job[1], job[2], job[3])
        Vm = job[3] / so.v_capacitance(b, h['z']) * 1e3
        eng = {'V': float(np.mean(Vm))}
        eng.update({k: float(v) for k, v in so.eff_rates('RS', Vm).items()})
        ref = {k: float(g['tab_' + k][0, 0, iA, iQ, 0]) for k in keys}
        e_ref.append(max(rel(ref[k], tev[k]) for k in keys))
        e_eng.append(max(rel(eng[k], tev[k]) for k in keys))
        e_re.append(max(rel(eng[k], ref[k]) for k in keys))
        cyc.append((int(g['ncycles'][0, 0, iA, iQ]), int(h['ncycles']), tcyc, job[2]))
    e_ref, e_eng, e_re = np.array(e_ref), np.array(e_eng), np.array(e_re)
    cyc = np.array(cyc, float)

    def stats(e):
        return {'median': float(np.median(e)), 'p90': float(np.percentile(e, 90)), 'p99': float(np.percentile(e, 99)),
                'max': float(e.max()), 'frac_within_1e-4': float(np.mean(e <= 1e-4)), 'frac_within_1e-5': float(np.mean(e <= 1e-5))}
    hi = cyc[:, 3] >= 10e3
    res = {'fixture': 'c1_RS_32nm_500kHz.npz', 'points': len(pts), 'stride': stride,
           'tight_tolerances': 'rtol 1e-12, atol 1e-12 x (1 m/s, 1 nm, ng0)',
           'worst_relative_deviation_per_point_over_V_and_rates': {
               'reference_vs_tight': stats(e_ref), 'engine_vs_tight': stats(e_eng), 'engine_vs_reference': stats(e_re)},
           'ncycles': {'reference_eq_tight': float(np.mean(cyc[:, 0] == cyc[:, 2])), 'engine_eq_tight': float(np.mean(cyc[:, 1] == cyc[:, 2])),
                       'engine_eq_reference': float(np.mean(cyc[:, 0] == cyc[:, 1])),
                       'A_ge_10kPa': {'reference_eq_tight': float(np.mean(cyc[hi, 0] == cyc[hi, 2])),
                                      'engine_eq_tight': float(np.mean(cyc[hi, 1] == cyc[hi, 2])),
                                      'engine_eq_reference': float(np.mean(cyc[hi, 0] == cyc[hi, 1]))}}}
    print(json.dumps(res, indent=1))
    if out:
        with open(out, 'w') as fh:
            json.dump(res, fh, indent=1)


if __name__ == '__main__':
    main()
