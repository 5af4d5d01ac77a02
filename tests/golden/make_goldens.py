# -*- coding: utf-8 -*-
''' Golden-vector generator: runs the UNMODIFIED reference (/root/reference, through
    `_refshim`) and stores its outputs as small fixtures next to this script.

    Build container only -- /root/reference does not exist on the GPU box.

        python tests/golden/make_goldens.py points      # known-answer points, all neurons
        python tests/golden/make_goldens.py rates       # rate functions on a Vm sweep
        python tests/golden/make_goldens.py c1          # BASELINE config 1 (1000 points)
        python tests/golden/make_goldens.py neurons     # small grids for the C3-C5 neurons
        python tests/golden/make_goldens.py c2sub       # stratified subsample of RS-4D (C2)
        python tests/golden/make_goldens.py noise       # reference re-run with +-2 ulp amplitude
        python tests/golden/make_goldens.py noise4      # small grids re-run with +-4 ulp amplitude
        python tests/golden/make_goldens.py c2big       # dense RS 4-D fixture (+-2 ulp amplitude re-runs)
        python tests/golden/make_goldens.py c2big_q     # ... re-run with the charge changed by +-2 ulp (_ulp_up2/_dn2)
        python tests/golden/make_goldens.py neurons_big # >= 1000-point fixtures for the C3-C5 neurons
        python tests/golden/make_goldens.py all

    Every record is produced by `NeuronalBilayerSonophore.computeEffVars`
    (PySONIC/core/nbls.py:153-222); cycle counts are read off the solver by wrapping
    `PeriodicSolver.integrateCycle` (solvers.py:332-334), RHS-call counts by wrapping
    `BilayerSonophore.derivatives` (bls.py:681).
'''

import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refshim import load_reference  # noqa: E402

load_reference()
import PySONIC.core.solvers as _solvers  # noqa: E402
import PySONIC.core.bls as _bls  # noqa: E402
from PySONIC.core import NeuronalBilayerSonophore, AcousticDrive  # noqa: E402
from PySONIC.neurons import getPointNeuron  # noqa: E402

_counters = {'ncycles': 0, 'nfe': 0}
_orig_cycle = _solvers.PeriodicSolver.integrateCycle
_orig_der = _bls.BilayerSonophore.derivatives


def _cycle(self):
    _counters['ncycles'] += 1
    return _orig_cycle(self)


_rhs_scale = [1.0]     # rounding-level scaling of the right-hand side (see gen_c2big_rhs)


def _der(self, *a, **k):
    _counters['nfe'] += 1
    out = _orig_der(self, *a, **k)
    if _rhs_scale[0] != 1.0:
        out = [x * _rhs_scale[0] for x in out]
    return out


_solvers.PeriodicSolver.integrateCycle = _cycle
_bls.BilayerSonophore.derivatives = _der

_nbls_cache = {}


def _nbls(name, a):
    key = (name, a)
    if key not in _nbls_cache:
        _nbls_cache[key] = NeuronalBilayerSonophore(a, getPointNeuron(name))
    return _nbls_cache[key]


def run_point(args):
    ''' :return: dict with effvars (one dict per fs), ncycles, nfe, tcomp '''
    name, a, f, A, Q, fs = args
    nbls = _nbls(name, a)
    _counters['ncycles'] = 0
    _counters['nfe'] = 0
    t0 = time.perf_counter()
    effvars, _ = nbls.computeEffVars(AcousticDrive(float(f), float(A)), np.asarray(fs, float),
                                     float(Q))
    tcomp = time.perf_counter() - t0
    return {
        'neuron': name, 'a': a, 'f': f, 'A': A, 'Q': Q, 'fs': list(map(float, fs)),
        'ncycles': _counters['ncycles'], 'nfe': _counters['nfe'], 'tcomp': tcomp,
        'effvars': [{k: float(v) for k, v in ev.items()} for ev in effvars],
    }


def pmap(jobs, nproc=None):
    nproc = nproc or int(os.environ.get('GOLDEN_NPROC', 0)) or mp.cpu_count()
    with mp.get_context('fork').Pool(nproc) as pool:
        return pool.map(run_point, jobs, chunksize=1)


def default_charges(name):
    pn = getPointNeuron(name)
    Qmin, Qmax = pn.Qbounds
    return np.arange(Qmin, Qmax + 1e-5, 1e-5)   # run_lookups.py:196-199


def c2_amps():
    return np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3


# ---------------------------------------------------------------------------------------------

def gen_points():
    ''' Known-answer points: SURVEY Appendix A set + corners + fs vectors. '''
    jobs = [
        ('RS', 32e-9, 500e3, 100e3, -71.9e-5, [1.0]),
        ('RS', 32e-9, 500e3, 600e3, 50e-5, [1.0]),
        ('RS', 16e-9, 20e3, 50e3, -107e-5, [1.0]),
        ('RS', 64e-9, 4e6, 300e3, 0.0, [1.0]),
        ('RS', 32e-9, 500e3, 100e3, -71.9e-5, [0.5]),
        ('RS', 32e-9, 500e3, 0.0, -71.9e-5, [1.0]),
        ('RS', 32e-9, 500e3, 1e3, -50e-5, [1.0]),
        ('RS', 32e-9, 1e6, 300e3, -30e-5, [1.0]),
        ('RS', 16e-9, 4e6, 600e3, -107e-5, [1.0]),
        ('RS', 64e-9, 20e3, 100e3, 50e-5, [1.0]),
        ('RS', 64e-9, 4e6, 600e3, 50e-5, [1.0]),
        ('RS', 32e-9, 100e3, 20e3, 10e-5, [1.0]),
        ('RS', 32e-9, 2e6, 5e3, -71.9e-5, [1.0]),
        ('RS', 32e-9, 3e6, 600e3, -107e-5, [1.0]),
        ('STN', 32e-9, 500e3, 100e3, -58e-5, [0.75]),
        ('STN', 32e-9, 500e3, 300e3, -93e-5, [0.01, 0.25, 0.5, 0.75, 1.0]),
        ('STN', 32e-9, 500e3, 50e3, 20e-5, list(np.arange(1, 101) * 1e-2)),
        ('FHnode', 32e-9, 500e3, 100e3, -140e-5, [1.0]),
        ('FHnode', 32e-9, 4e6, 600e3, 100e-5, [1.0]),
        ('SWnode', 32e-9, 500e3, 100e3, -200e-5, [1.0]),
        ('SWnode', 32e-9, 100e3, 300e3, -287.5e-5, [1.0]),
        ('MRGnode', 32e-9, 500e3, 100e3, -160e-5, [1.0]),
        ('MRGnode', 32e-9, 2e6, 400e3, 100e-5, [1.0]),
        ('SUseg', 32e-9, 500e3, 100e3, -60e-5, [1.0]),
        ('SUseg', 32e-9, 1e6, 600e3, -95e-5, [1.0]),
        ('RE', 32e-9, 500e3, 300e3, -89.5e-5, [1.0]),
        ('RE', 32e-9, 4e6, 600e3, -124e-5, [1.0]),
        ('RE', 32e-9, 20e3, 50e3, 50e-5, [1.0]),
        ('TC', 32e-9, 500e3, 300e3, -61.93e-5, [1.0]),
        ('TC', 32e-9, 500e3, 600e3, -97e-5, [1.0]),
        ('TC', 32e-9, 4e6, 50e3, 50e-5, [1.0]),
        ('FS', 32e-9, 500e3, 100e3, -71.4e-5, [1.0]),
        ('LTS', 32e-9, 500e3, 100e3, -54e-5, [1.0]),
        ('IB', 32e-9, 500e3, 100e3, -71.4e-5, [1.0]),
    ]
    recs = pmap(jobs)
    # the reference's own reproducibility on each point: re-run with A * (1 +- 2 ulp)
    for scale in (1.0 + 4.440892098500626e-16, 1.0 - 4.440892098500626e-16):
        pert = pmap([(j[0], j[1], j[2], j[3] * scale, j[4], j[5]) for j in jobs])
        for r, q in zip(recs, pert):
            dev = 0.0
            for ev, evq in zip(r['effvars'], q['effvars']):
                for k in ev:
                    d = abs(ev[k] - evq[k])
                    if d >= 1e-9:
                        dev = max(dev, d / abs(ev[k]))
            r['self_noise'] = max(r.get('self_noise', 0.0), dev)
            r.setdefault('ncycles_ulp', []).append(q['ncycles'])
    # constants of the sonophore instances, for the host-side parameter tests
    consts = {}
    for name, a in sorted({(j[0], j[1]) for j in jobs}):
        nb = _nbls(name, a)
        consts[f'{name}@{a * 1e9:.0f}nm'] = {
            'Delta': nb.Delta, 'ng0': nb.ng0, 'V0': nb.V0, 'Zmin': nb.Zmin, 'S0': nb.S0,
            'Cm0': nb.Cm0, 'Qm0': nb.Qm0, 'LJ': nb.LJ_approx,
            'Qbounds': list(map(float, nb.pneuron.Qbounds)), 'rates': list(nb.pneuron.rates)}
    # a few initial deflections (bls.py:720-725)
    z0 = []
    for name, a, f, A, Q in [('RS', 32e-9, 500e3, 100e3, -71.9e-5), ('RS', 32e-9, 20e3, 100e3, -71.9e-5),
                             ('RS', 16e-9, 500e3, 600e3, 50e-5), ('RS', 64e-9, 4e6, 0., 0.),
                             ('SWnode', 32e-9, 500e3, 600e3, -287.5e-5)]:
        nb = _nbls(name, a)
        drive = AcousticDrive(f, A)
        z0.append({'neuron': name, 'a': a, 'f': f, 'A': A, 'Q': Q,
                   'Z0': float(nb.computeInitialDeflection(drive, Q, drive.dt))})
    with open(os.path.join(HERE, 'points.json'), 'w') as fh:
        json.dump({'points': recs, 'consts': consts, 'Z0': z0}, fh, indent=1)
    print('points.json:', len(recs), 'points')


def gen_points2():
    ''' Known-answer points for the remaining @addSonicFeatures neurons (hh.py, leech.py, template.py) and
        the passive membrane (pas.py; run_lookups.py:141-145 simulates it on the default passive neuron). '''
    from PySONIC.neurons import getDefaultPassiveNeuron
    jobs = []
    for name, Q in [('HHseg', -65e-5), ('LeechT', -53.58e-5), ('LeechP', -48.865e-5), ('template', -71.9e-5)]:
        jobs += [(name, 32e-9, 500e3, 100e3, Q, [1.0]), (name, 32e-9, 2e6, 400e3, 30e-5, [0.5, 1.0]),
                 (name, 32e-9, 100e3, 20e3, -80e-5, [1.0])]
    recs = pmap(jobs)
    for scale in (1.0 + 4.440892098500626e-16, 1.0 - 4.440892098500626e-16):
        pert = pmap([(j[0], j[1], j[2], j[3] * scale, j[4], j[5]) for j in jobs])
        for r, q in zip(recs, pert):
            dev = 0.0
            for ev, evq in zip(r['effvars'], q['effvars']):
                for k in ev:
                    d = abs(ev[k] - evq[k])
                    if d >= 1e-9:
                        dev = max(dev, d / abs(ev[k]))
            r['self_noise'] = max(r.get('self_noise', 0.0), dev)
    # passive membrane: only V
    pas = getDefaultPassiveNeuron()
    nb = NeuronalBilayerSonophore(32e-9, pas)
    prec = []
    for f, A, Q, fs in [(500e3, 100e3, -70e-5, [1.0]), (1e6, 300e3, 20e-5, [0.3, 1.0]), (20e3, 50e3, -100e-5, [1.0])]:
        _counters['ncycles'] = 0
        ev, _ = nb.computeEffVars(AcousticDrive(f, A), np.array(fs), Q)
        prec.append({'neuron': pas.name, 'lookup_name': pas.lookup_name, 'a': 32e-9, 'f': f, 'A': A, 'Q': Q, 'fs': fs,
                     'ncycles': _counters['ncycles'], 'effvars': [{k: float(v) for k, v in e.items()} for e in ev],
                     'self_noise': 0.0, 'Qbounds': list(map(float, pas.Qbounds)), 'Qm0': pas.Qm0,
                     'fname': nb.getLookupFileName()})
    consts = {}
    for name in ['HHseg', 'LeechT', 'LeechP', 'template']:
        nb = _nbls(name, 32e-9)
        consts[f'{name}@32nm'] = {'Delta': nb.Delta, 'Cm0': nb.Cm0, 'Qm0': nb.Qm0, 'LJ': nb.LJ_approx,
                                  'Qbounds': list(map(float, nb.pneuron.Qbounds)), 'rates': list(nb.pneuron.rates)}
    with open(os.path.join(HERE, 'points_r02.json'), 'w') as fh:
        json.dump({'points': recs, 'passive': prec, 'consts': consts}, fh, indent=1)
    print('points_r02.json:', len(recs), '+', len(prec), 'points')


def gen_rates():
    ''' Every tabulated rate function of every supported neuron on a Vm sweep that includes
        the singular / branch points (vtrap 0/0, tauu branch). '''
    names = ['RS', 'FS', 'LTS', 'IB', 'RE', 'TC', 'STN', 'FHnode', 'SWnode', 'MRGnode', 'SUseg',
             'HHseg', 'LeechT', 'LeechP', 'template']
    Vm = np.concatenate([np.linspace(-450., 350., 401),
                         np.array([-43.2, -16.2, -41.2, -48.0, -57.0, -35.0, -61.0, -27.0,
                                   -21.4, -25.7, -114.0, -80.0, -73.0, -87.0 + 7.0, -83.0,
                                   -46.9, -18.9, -58.9, -10.9])])
    out = {'Vm': Vm.tolist(), 'neurons': {}}
    with np.errstate(all='ignore'):
        for n in names:
            pn = getPointNeuron(n)
            tab = {}
            for k, fn in pn.effRates().items():
                tab[k] = [float(fn(np.float64(v))) for v in Vm]
            out['neurons'][n] = {'rates': list(pn.rates), 'Cm0': pn.Cm0, 'Vm0': pn.Vm0, 'values': tab}
    with open(os.path.join(HERE, 'rates_sweep.json'), 'w') as fh:
        json.dump(out, fh)
    print('rates_sweep.json:', len(names), 'neurons x', Vm.size, 'potentials')


def _grid_to_npz(fname, name, aref, fref, Aref, Qref, fsref, amp_scale=1.0, q_scale=1.0):
    # amp_scale != 1 (a 1-2 ulp relative change of the drive amplitude) is used to measure the
    # reference's own reproducibility floor: how far its outputs move under a rounding-level
    # perturbation of its inputs.
    # (q_scale != 1: the same for the imposed charge -- a rounding-level change of the drive amplitude is
    #  invisible below ~1 kPa, where the drive is a 1e-3 ... 1e-2 fraction of the static pressures, so the
    #  low-amplitude rows need a perturbation that reaches the dynamics)
    jobs = [(name, a, f, A * amp_scale, Q * q_scale, list(fsref))
            for a in aref for f in fref for A in Aref for Q in Qref]
    t0 = time.perf_counter()
    recs = pmap(jobs)
    wall = time.perf_counter() - t0
    dims = (len(aref), len(fref), len(Aref), len(Qref), len(fsref))
    keys = list(recs[0]['effvars'][0].keys())
    tables = {k: np.array([ev[k] for r in recs for ev in r['effvars']]).reshape(dims) for k in keys}
    np.savez_compressed(
        os.path.join(HERE, fname), neuron=name, a=np.array(aref), f=np.array(fref),
        A=np.array(Aref), Q=np.array(Qref), fs=np.array(fsref), keys=np.array(keys),
        ncycles=np.array([r['ncycles'] for r in recs]).reshape(dims[:-1]),
        nfe=np.array([r['nfe'] for r in recs]).reshape(dims[:-1]),
        tcomp=np.array([r['tcomp'] for r in recs]).reshape(dims[:-1]),
        wall_s=wall, nproc=mp.cpu_count(), **{f'tab_{k}': v for k, v in tables.items()})
    print(f'{fname}: {len(jobs)} points in {wall:.1f} s on {mp.cpu_count()} processes')


def gen_c1(amp_scale=1.0, tag=''):
    ''' BASELINE config 1 (SURVEY 8d "C1"). '''
    A = np.insert(np.logspace(np.log10(0.1), np.log10(600), 19), 0, 0.) * 1e3
    Q = np.linspace(-107e-5, 50e-5, 50)
    _grid_to_npz(f'c1_RS_32nm_500kHz{tag}.npz', 'RS', [32e-9], [500e3], A, Q, [1.0], amp_scale)


def gen_neurons(amp_scale=1.0, tag=''):
    ''' Small grids for the neurons of BASELINE configs 3-5. '''
    Aall = c2_amps()
    for name in ['FHnode', 'SWnode', 'MRGnode', 'SUseg']:
        Q = default_charges(name)
        _grid_to_npz(f'c4_{name}_sub{tag}.npz', name, [32e-9], [100e3, 500e3, 4e6],
                     Aall[[0, 20, 36, 44, 50]], Q[::max(1, Q.size // 6)], [1.0], amp_scale)
    for name in ['RE', 'TC']:
        pn = getPointNeuron(name)
        Qmin, Qmax = pn.Qbounds
        Q = np.arange(Qmin, Qmax + 5e-6, 5e-6)
        A = np.logspace(np.log10(50), np.log10(600), 26) * 1e3
        _grid_to_npz(f'c5_{name}_sub{tag}.npz', name, [32e-9], [20e3, 500e3, 4e6],
                     A[[0, 12, 25]], Q[::max(1, Q.size // 6)], [1.0], amp_scale)
    Q = default_charges('STN')
    _grid_to_npz(f'c3_STN_sub{tag}.npz', 'STN', [32e-9], [500e3], Aall[[0, 16, 30, 40, 50]],
                 Q[::24], np.arange(1, 101)[::11] * 1e-2, amp_scale)


def gen_c2sub(amp_scale=1.0, tag=''):
    ''' Stratified subsample of the RS 4-D default grid (SURVEY 8d "C2"). '''
    A = c2_amps()[::7]            # 8 amplitudes incl. 0 and ~502 kPa
    A = np.append(A, c2_amps()[-1])
    Q = default_charges('RS')[::22]
    _grid_to_npz(f'c2_RS_sub{tag}.npz', 'RS', [16e-9, 32e-9, 64e-9],
                 [20e3, 100e3, 500e3, 1e6, 2e6, 3e6, 4e6], A, Q, [1.0], amp_scale)


def gen_c2big(amp_scale=1.0, tag='', q_scale=1.0):
    ''' Dense subsample of the RS 4-D default grid: every radius, frequency and amplitude (the
        16 nm / 20 kHz heavy rows included), every 16th charge -> 3 x 7 x 51 x 10 = 10 710 points. '''
    _grid_to_npz(f'c2_RS_big{tag}.npz', 'RS', [16e-9, 32e-9, 64e-9],
                 [20e3, 100e3, 500e3, 1e6, 2e6, 3e6, 4e6], c2_amps(), default_charges('RS')[::16], [1.0],
                 amp_scale, q_scale)


def gen_c2big_rhs():
    ''' The dense RS fixture re-run with the right-hand side itself changed at rounding level (every
        derivative multiplied by 1 +- 2.2e-16): below ~8 kPa neither a 2-ulp change of the amplitude nor of
        the charge reaches the dynamics at every point, a rounding inside the integration does -- which is
        what any re-implementation of the arithmetic amounts to. '''
    for sc, tag in ((1.0 + 2.220446049250313e-16, '_ulp_up3'), (1.0 - 1.1102230246251565e-16, '_ulp_dn3')):
        _rhs_scale[0] = sc
        gen_c2big(1.0, tag)
    _rhs_scale[0] = 1.0


def gen_neurons_big(amp_scale=1.0, tag=''):
    ''' >= 1000-point grids for each neuron of BASELINE configs 3-5 (32 nm). '''
    Aall = c2_amps()
    for name in ['FHnode', 'SWnode', 'MRGnode', 'SUseg']:
        Q = default_charges(name)
        Q = Q[::max(1, Q.size // 12)][:13]
        _grid_to_npz(f'c4_{name}_big{tag}.npz', name, [32e-9], [20e3, 100e3, 500e3, 1e6, 2e6, 3e6, 4e6],
                     Aall[[0, 8, 16, 22, 28, 32, 36, 40, 43, 46, 48, 50]], Q, [1.0], amp_scale)
    for name in ['RE', 'TC']:
        pn = getPointNeuron(name)
        Qmin, Qmax = pn.Qbounds
        Q = np.arange(Qmin, Qmax + 5e-6, 5e-6)
        A = np.logspace(np.log10(50), np.log10(600), 26) * 1e3
        _grid_to_npz(f'c5_{name}_big{tag}.npz', name, [32e-9], [20e3, 500e3, 4e6],
                     A[::2], Q[::max(1, Q.size // 26)][:27], [1.0], amp_scale)
    Q = default_charges('STN')
    _grid_to_npz(f'c3_STN_big{tag}.npz', 'STN', [32e-9], [500e3], Aall[::2],
                 Q[::4], np.arange(1, 101)[::11] * 1e-2, amp_scale)


def gen_cortical(amp_scale=1.0, tag=''):
    ''' Small grids for the other cortical neurons that share the RS kinetics family (FS, LTS, IB):
        two radii, three frequencies, five amplitudes, six charges. '''
    Aall = c2_amps()
    for name in ['FS', 'LTS', 'IB']:
        Q = default_charges(name)
        _grid_to_npz(f'c7_{name}_sub{tag}.npz', name, [16e-9, 64e-9], [100e3, 1e6, 3e6],
                     Aall[[0, 24, 38, 46, 50]], Q[::max(1, Q.size // 5)], [1.0], amp_scale)


def _ov_point(args):
    name, a, f, A, Q, fs, ov = args
    _counters['ncycles'] = 0
    ev, tcomp = _nbls(name, a).computeEffVars(AcousticDrive(f, A), np.array(fs), Q, ov)
    return {'effvars': [{k: float(v) for k, v in e.items()} for e in ev], 'ncycles': _counters['ncycles'], 'tcomp': tcomp}


def gen_overtones():
    ''' Charge-overtone variant of computeEffVars (nbls.py:169-201, run_lookups.py:105-128):
        RS, 32 nm, 500 kHz, sub-grid of (A, Q, AQ1, phiQ1) plus one two-overtone point. '''
    A = [50e3, 300e3]
    Q = [-80e-5, -20e-5, 30e-5]
    AQ = [0., 50e-5, 100e-5]
    phi = [0., 2 * np.pi / 5, 6 * np.pi / 5]
    jobs = [('RS', 32e-9, 500e3, a_, q_, [1.0], [(aq, ph)]) for a_ in A for q_ in Q for aq in AQ for ph in phi]
    jobs.append(('RS', 32e-9, 500e3, 100e3, -71.9e-5, [0.5, 1.0], [(50e-5, 1.0), (25e-5, 4.0)]))
    with mp.get_context('fork').Pool(mp.cpu_count()) as pool:
        recs = pool.map(_ov_point, jobs, chunksize=1)
    out = {'grid': {'neuron': 'RS', 'a': 32e-9, 'f': 500e3, 'A': A, 'Q': Q, 'AQ1': AQ, 'phiQ1': phi, 'fs': [1.0]},
           'points': [dict(neuron=j[0], a=j[1], f=j[2], A=j[3], Q=j[4], fs=j[5], overtones=[list(x) for x in j[6]], **r)
                      for j, r in zip(jobs, recs)]}
    with open(os.path.join(HERE, 'overtones.json'), 'w') as fh:
        json.dump(out, fh, indent=1)
    print('overtones.json:', len(jobs), 'points')


def _cm_point(args):
    a, f, A, Qm = args
    from PySONIC.core import BilayerSonophore
    bls = BilayerSonophore(a, 1e-2, 0.0)
    return bls.getRelCmCycle(AcousticDrive(f, A), Qm)


def gen_cm():
    ''' Relative-capacitance profiles of scripts/run_Cm_lookups.py (BilayerSonophore(32 nm, 1e-2, 0),
        getRelCmCycle(drive, 0.), bls.py:806-813) on a sub-grid of its default (f, A) grid. '''
    f = np.array([100e3, 500e3, 4e6])
    A = c2_amps()[[0, 30, 41, 50]]
    jobs = [(32e-9, ff, AA, 0.) for ff in f for AA in A]
    with mp.get_context('fork').Pool(min(mp.cpu_count(), len(jobs))) as pool:
        out = pool.map(_cm_point, jobs)
    np.savez_compressed(os.path.join(HERE, 'cm_lkp_32nm_sub.npz'), a=32e-9, f=f, A=A,
                        t=np.linspace(0., 1., out[0].size), Cm_rel=np.array(out).reshape(f.size, A.size, -1))
    print('cm_lkp_32nm_sub.npz:', len(jobs), 'profiles of', out[0].size, 'samples')


def gen_noise():
    ''' The reference re-run with the drive amplitude changed by +-2 ulp: its own
        reproducibility floor on the C1 grid and the C2 subsample. '''
    up, dn = 1.0 + 4.440892098500626e-16, 1.0 - 4.440892098500626e-16
    gen_c1(up, '_ulp_up')
    gen_c1(dn, '_ulp_dn')
    gen_c2sub(up, '_ulp_up')
    gen_c2sub(dn, '_ulp_dn')
    gen_neurons(up, '_ulp_up')
    gen_neurons(dn, '_ulp_dn')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    todo = {'points': gen_points, 'points2': gen_points2, 'rates': gen_rates, 'c1': gen_c1, 'neurons': gen_neurons,
            'c2sub': gen_c2sub, 'cm': gen_cm,
            'noise4': lambda: [fn(sc, tg) for fn in (gen_neurons, gen_cortical)
                               for sc, tg in ((1.0 + 8.881784197001252e-16, '_ulp_up2'),
                                              (1.0 - 8.881784197001252e-16, '_ulp_dn2'))],
            'cortical': lambda: (gen_cortical(), gen_cortical(1.0 + 4.440892098500626e-16, '_ulp_up'),
                                 gen_cortical(1.0 - 4.440892098500626e-16, '_ulp_dn')), 'overtones': gen_overtones, 'noise': gen_noise,
            'c2big': lambda: (gen_c2big(), gen_c2big(1.0 + 4.440892098500626e-16, '_ulp_up'),
                              gen_c2big(1.0 - 4.440892098500626e-16, '_ulp_dn')),
            'c2big_q': lambda: (gen_c2big(1.0, '_ulp_up2', 1.0 + 4.440892098500626e-16),
                                gen_c2big(1.0, '_ulp_dn2', 1.0 - 4.440892098500626e-16)),
            'c2big_rhs': gen_c2big_rhs,
            'neurons_big': lambda: (gen_neurons_big(), gen_neurons_big(1.0 + 4.440892098500626e-16, '_ulp_up'),
                                    gen_neurons_big(1.0 - 4.440892098500626e-16, '_ulp_dn')),
            'noise_neurons': lambda: (gen_neurons(1.0 + 4.440892098500626e-16, '_ulp_up'),
                                      gen_neurons(1.0 - 4.440892098500626e-16, '_ulp_dn'))}
    for k, fn in todo.items():
        if what == k or (what == 'all' and k not in ('noise_neurons', 'overtones', 'cm', 'cortical', 'noise4', 'c2big', 'c2big_q', 'c2big_rhs', 'neurons_big', 'points2')):
            fn()
