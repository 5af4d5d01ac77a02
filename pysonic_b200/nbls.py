# -*- coding: utf-8 -*-
''' Neuronal bilayer sonophore: effective-variable computation on the GPU
    (mirror of PySONIC/core/nbls.py:24-39,153-244 -- lookup-relevant part). '''

import logging
import os
import time

import numpy as np

from . import _lib
from .bls import BilayerSonophore
from .constants import CHARGE_RANGE, NPC_DENSE
from .lookups import Lookup
from .drives import AcousticDrive
from .neurons import PointNeuron, check_foreign_neuron, getPointNeuron

logger = logging.getLogger('pysonic_b200')

LOOKUP_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lookups')


def as_point_neuron(pneuron):
    ''' Accept this package's descriptor, a neuron name, or any object with a `.name`
        (e.g. a reference PySONIC PointNeuron instance). '''
    if isinstance(pneuron, PointNeuron):
        return pneuron
    if isinstance(pneuron, str):
        return getPointNeuron(pneuron)
    if hasattr(pneuron, 'name'):
        return check_foreign_neuron(pneuron, getPointNeuron(pneuron.name))
    raise ValueError(f'{pneuron} is not a valid PointNeuron instance')


def check_drive_phase(drive):
    ''' The kernels integrate the reference's default drive phase (phi = pi, drives.py:200): any other
        phase is refused instead of being silently replaced. '''
    phi = getattr(drive, 'phi', np.pi)
    if phi != np.pi:
        raise ValueError(f'unsupported acoustic drive phase {phi} rad (the lookup path uses phi = pi)')


def check_charges(Q, overtones=None):
    ''' Imposed charges (and the extrema of Fourier-series charge cycles) must lie in the physiological
        range, with the reference's message (bls.py:674-677 `checkInputs`).  The reference applies this
        check to `simulate` calls only (model.py:169): `simCycles` / `computeEffVars` -- the lookup path
        -- accept any charge (its own overtone grid reaches -107 - 2 x 100 nC/cm2), so the lookup entry
        points run it on request only (`check_charge=True`). '''
    Qmin, Qmax = CHARGE_RANGE
    Q = np.atleast_1d(np.asarray(Q, dtype=float))
    lo, hi = Q, Q
    if overtones is not None:
        ov = np.asarray(overtones, dtype=float).reshape(Q.size, -1, 2)
        swing = 2 * np.abs(ov[:, :, 0]).sum(axis=1)
        suspect = np.nonzero((Q - swing < Qmin) | (Q + swing > Qmax))[0]
        lo, hi = Q.copy(), Q.copy()
        j = np.arange(NPC_DENSE)
        for i in suspect:      # exact extrema of the sampled cycle, only where the bound is exceeded
            k = np.arange(1, ov.shape[1] + 1)[:, None]
            cyc = Q[i] + 2 * (ov[i, :, 0:1] * np.cos(2 * np.pi * j * k / NPC_DENSE + ov[i, :, 1:2])).sum(axis=0)
            lo[i], hi[i] = cyc.min(), cyc.max()
    bad = np.nonzero((lo < Qmin) | (hi > Qmax))[0]
    if bad.size:
        q = Q[bad[0]]
        raise ValueError(
            f'Invalid applied charge: {q * 1e5} nC/cm2 (must be within [{Qmin * 1e5}, {Qmax * 1e5}] interval')


class NeuronalBilayerSonophore(BilayerSonophore):

    def __init__(self, a, pneuron, embedding_depth=0.0):
        self.pneuron = as_point_neuron(pneuron)
        super().__init__(a, self.pneuron.Cm0, self.pneuron.Qm0, embedding_depth=embedding_depth)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.a * 1e9:.1f} nm, {self.pneuron})'

    def effvars_batch(self, f, A, Q, fs, device=0, overtones=None, check_charge=False):
        ''' Effective variables for arrays of points (same radius).
            :param overtones: None, or charge overtones [n, novertones, 2] (amplitude C/m2, phase rad)
            :return: (tables[1+2*novertones+nrates, n, nfs], ncycles, status, tpoint, nrhs, stats) '''
        f, A, Q = np.broadcast_arrays(np.asarray(f, float), np.asarray(A, float), np.asarray(Q, float))
        fs = np.atleast_1d(np.asarray(fs, float))
        if check_charge:
            check_charges(Q.ravel(), overtones)
        ia = np.zeros(f.size, dtype=np.int32)
        return _lib.points_run(device, [self.abi_params()], self.pneuron.neuron_id,
                               len(self.pneuron.rates), ia, f.ravel(), A.ravel(), Q.ravel(), fs,
                               overtones=overtones)

    def effvars_keys(self, novertones=0):
        ''' Key order of the effective-variable dictionaries (nbls.py:191-204). '''
        keys = ['V']
        for i in range(1, novertones + 1):
            keys += [f'A_V{i}', f'phi_V{i}']
        return keys + self.pneuron.rates

    def computeEffVars(self, drive, fs, Qm0, Qm_overtones=None):
        ''' Effective coefficients for one acoustic drive and charge density
            (nbls.py:153-222; returns `(effvars_list, tcomp)` like the `@timer`-decorated
            reference method, utils.py:408-417).

            :param drive: acoustic drive object
            :param fs: sonophore membrane coverage fraction(s)
            :param Qm0: imposed charge density (C/m2)
            :param Qm_overtones: optional list of (amplitude C/m2, phase rad) pairs: the imposed
                charge is then the Fourier-series cycle of nbls.py:173-178
            :return: (list of one {V, [A_Vk, phi_Vk...], rates...} dict per fs, computation time in s)
        '''
        ov = None
        if Qm_overtones is not None:
            ov = np.asarray(Qm_overtones, dtype=float).reshape(1, -1, 2)
        if not isinstance(drive, AcousticDrive) and not (hasattr(drive, 'f') and hasattr(drive, 'A')):
            raise TypeError('Invalid "drive" parameter (must be an "AcousticDrive" object)')
        check_drive_phase(drive)
        t0 = time.perf_counter()
        fs = np.atleast_1d(np.asarray(fs, dtype=float))
        out, ncyc, status, _, _, _ = self.effvars_batch(drive.f, drive.A, float(Qm0), fs, overtones=ov)
        keys = self.effvars_keys(0 if ov is None else ov.shape[1])
        effvars_list = [{k: out[i, 0, j] for i, k in enumerate(keys)} for j in range(fs.size)]
        if status[0] & 1:
            logger.warning('%s: periodic criterion not met -> stopped after %d cycles', self, ncyc[0])
        return effvars_list, time.perf_counter() - t0

    def getLookupFileName(self, a=None, f=None, A=None, fs=None, novertones=0):
        ''' nbls.py:224-241 '''
        if all(x is None for x in [a, f, A, fs]):
            fs = 1.
        fname = f"{getattr(self.pneuron, 'lookup_name', self.pneuron.name)}_lookups"
        if a is not None:
            fname += f'_{a * 1e9:.0f}nm'
        if f is not None:
            fname += f'_{f * 1e-3:.0f}kHz'
        if A is not None:
            fname += f'_{A * 1e-3:.0f}kPa'
        if fs is not None:
            fname += f'_fs{fs:.2f}'
        if novertones > 0:
            fname += f'_{novertones}overtones'
        return f'{fname}.pkl'

    def getLookupFilePath(self, *args, **kwargs):
        return os.path.join(LOOKUP_DIR, self.getLookupFileName(*args, **kwargs))

    # ---- table consumption: SONIC simulation (nbls.py:246-263,389-437,513-536) --------------------
    def getLookup(self, *args, keep_tcomp=False, **kwargs):
        ''' Load the lookup file of this neuron (nbls.py:246-251). '''
        lkp = Lookup.fromPickle(self.getLookupFilePath(*args, **kwargs))
        if not keep_tcomp and 'tcomp' in lkp.tables:
            del lkp.tables['tcomp']
        return lkp

    def getLookup2D(self, f, fs, lookup=None):
        ''' (A, Q) tables at this radius, frequency f and coverage fs (nbls.py:253-263); `lookup`: a
            Lookup object to project instead of the file of the lookup directory. '''
        if lookup is None:
            lookup = self.getLookup(**({'a': self.a, 'f': f, 'fs': None} if fs < 1. else {'fs': fs}))
        lkp = lookup.copy()
        if 'tcomp' in lkp.tables:
            del lkp.tables['tcomp']
        return lkp.projectN({k: v for k, v in (('a', self.a), ('f', f), ('fs', fs)) if k in lkp.refs})

    @staticmethod
    def _sample_plan(events, tstop, dt):
        ''' Sample times and stimulus states of EventDrivenSolver.solve (solvers.py:445-478): from the
            last time to every event a linspace of max(round(span / dt), 2) samples is appended with the
            stimulus state in force, then the event fires. '''
        events = sorted(events, key=lambda e: e[0])
        if events[-1][0] > tstop:
            raise ValueError('all events must occur before stopping time')
        t, x, xref = [0.], [0.], 0.
        for tev, xev in events + [(tstop, None)]:
            if tev < t[-1]:
                raise ValueError(f'target time ({tev} s) precedes current time {t[-1]} s')
            n = max(int(np.round((tev - t[-1]) / dt)), 2)
            seg = np.linspace(t[-1], tev, n)
            t += seg.tolist()
            x += [xref] * n
            if xev is not None:
                xref = xev
        return np.array(t), np.array(x)

    def simulate(self, drive, pp, fs=1., method='sonic', qss_vars=None, lookup=None, nsub=64, device=0):
        ''' SONIC simulation of the electro-mechanical model on the lookup tables (nbls.py:513-536 ->
            __simSonic :389-437), on the GPU.

            :param drive: acoustic drive object
            :param pp: pulsed protocol object
            :param fs: sonophore membrane coverage fraction (-)
            :param lookup: optional Lookup object (default: the neuron's lookup file)
            :return: (pandas DataFrame with the columns t, stimstate, Qm, Vm, states..., Z, ng -- the
                reference's TimeSeries --, metadata dict)
        '''
        if method != 'sonic':
            raise ValueError(f'Invalid integration method: "{method}" (only "sonic" runs on the tables)')
        if qss_vars is not None:
            raise NotImplementedError('quasi-steady-state variables are not supported')
        check_drive_phase(drive)
        datas = self.simulate_batch(drive.f, [drive.A], pp, fs=fs, lookup=lookup, nsub=nsub, device=device)
        meta = {'simkey': 'ASTIM', 'neuron': self.pneuron.name, 'a': self.a, 'fs': fs, 'method': method,
                'drive': drive, 'pp': pp, 'qss_vars': qss_vars}
        return datas[0], meta

    def simulate_batch(self, f, amps, pp, fs=1., lookup=None, nsub=64, device=0):
        ''' The same simulation for several drive amplitudes (Pa) in one launch; one DataFrame each. '''
        import pandas as pd
        from . import _lib
        from .constants import DT_EFFECTIVE, MAX_NSAMPLES_EFFECTIVE
        pn = self.pneuron
        if not pn.states:
            raise NotImplementedError(f'the SONIC simulation of the {pn.name} neuron is not supported')
        if not isinstance(fs, float):
            raise TypeError('Invalid "fs" parameter (must be float typed)')
        lkp2d = self.getLookup2D(float(f), fs, lookup)
        if lkp2d.inputs != ['A', 'Q']:
            raise ValueError(f'expected an (A, Q) lookup after projection, got {lkp2d.inputs}')
        keys = ['V'] + pn.rates
        amps = np.atleast_1d(np.asarray(amps, dtype=float))
        on = lkp2d.project('A', amps)                     # (nA, nQ) tables
        off = lkp2d.project('A', 0.)
        tab_on = np.stack([on[k] for k in keys], axis=1)  # [nsim, nvar, nQ]
        tab_off = np.stack([off[k] for k in keys], axis=0)
        # initial conditions: resting charge, steady states at the resting potential (nbls.py:410-414)
        r0 = _lib.eval_rates(pn, np.array([pn.Vm0]), device=device)
        y0 = [pn.Qm0] + [float(r0[f'alpha{k}'][0] / (r0[f'alpha{k}'][0] + r0[f'beta{k}'][0])) for k in pn.states]
        dt = DT_EFFECTIVE * pn.dt_factor               # pneuron.py:481-483 and the neuron-specific overrides
        t, x = self._sample_plan(pp.stimEvents(), pp.tstop, dt)
        nsub = int(np.clip(np.ceil(nsub * pn.dt_factor), 4, nsub))    # sub-steps of ~1 us whatever the output step
        out, status = _lib.simulate(pn.neuron_id, lkp2d.refs['Q'], tab_on, tab_off, t, x > 0, y0, nsub=nsub,
                                    device=device)
        datas = []
        Qref = lkp2d.refs['Q']
        for i, A in enumerate(amps):
            if status[i]:
                bad = out[i, :, 0]
                k = int(np.argmax(np.isnan(bad)))
                raise ValueError(f'Q value left the [{Qref.min()}, {Qref.max()}] interval of the lookup '
                                 f'(A = {A * 1e-3:g} kPa, t = {t[k] * 1e3:.3f} ms)')
            ti, xi, yi = t, x, out[i]
            if ti.size > MAX_NSAMPLES_EFFECTIVE:          # solvers.py:219-222
                tn = np.linspace(ti[0], ti[-1], max(int(np.round((ti[-1] - ti[0]) / (np.ptp(ti) / MAX_NSAMPLES_EFFECTIVE))), 2))
                yi = np.array([np.interp(tn, ti, c) for c in yi.T]).T
                xi = xi[np.clip(np.searchsorted(ti, tn), 0, ti.size - 1)]
                ti = tn
            Qm = yi[:, 0]
            Vm = np.where(xi > 0, np.interp(Qm, Qref, on['V'][i]), np.interp(Qm, Qref, off['V']))   # nbls.py:425-427
            cols = {'t': ti, 'stimstate': xi, 'Qm': Qm}
            for k, st in enumerate(pn.states):
                cols[st] = yi[:, 1 + k]
            cols['Vm'] = Vm                               # (the reference's addColumn leaves it after the states)
            cols['Z'] = np.full(ti.size, np.nan)          # nbls.py:432-434
            cols['ng'] = np.full(ti.size, np.nan)
            datas.append(pd.DataFrame(cols))
        return datas

    def getNSpikes(self, data):
        ''' Number of spikes in the charge profile (pneuron.py:545-551 -> postpro.detectSpikes / find_tpeaks):
            peaks of Qm at least 0.5 ms apart, above 3 nC/cm2, with a prominence of 20 nC/cm2 measured in a
            window of five times the narrowest peak. '''
        from scipy.signal import find_peaks, peak_prominences      # third-party, as in the reference
        from .constants import DT_MAX_REL_TOL, SPIKE_MIN_DT, SPIKE_MIN_QAMP, SPIKE_MIN_QPROM
        t, y = np.asarray(data['t'], dtype=float), np.asarray(data['Qm'], dtype=float)
        ipad = 0                                          # redundant initial samples (postpro.py:183-194)
        while t[ipad + 1] == t[ipad]:
            ipad += 1
        t, y = t[ipad:], y[ipad:]
        steps = np.diff(t)
        nz = steps[steps != 0]                            # repeated times at the stimulus transitions
        if (nz.max() - nz.min()) / nz.min() > DT_MAX_REL_TOL:     # irregular: resample (postpro.py:196-207)
            dt = max(steps.min(), 1e-7)
            tn = np.linspace(t.min(), t.max(), int(np.ptp(t) / dt) + 1)
            t, y = tn, np.interp(tn, t, y)
            nz = np.diff(t)
        dt = float(np.mean(nz))
        mph = SPIKE_MIN_QAMP if self.pneuron.spike_mph is None else self.pneuron.spike_mph
        ipeaks, pps = find_peaks(y, height=mph, distance=int(np.ceil(SPIKE_MIN_DT / dt)),
                                 prominence=SPIKE_MIN_QPROM, width=1)
        return int(ipeaks.size)
