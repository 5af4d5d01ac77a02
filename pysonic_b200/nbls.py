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
