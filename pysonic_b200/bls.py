# -*- coding: utf-8 -*-
''' Bilayer-sonophore constants needed by the lookup path (mirror of the constructor side of
    PySONIC/core/bls.py:80-137; the dynamics themselves run in the CUDA kernels). '''

import json
import os

import numpy as np

from .constants import Rg

_TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'bls_params.json')
_rows = None


def _table():
    global _rows
    if _rows is None:
        with open(_TABLE) as fh:
            _rows = {(r[0], r[1]): r[2:] for r in json.load(fh)['rows']}
    return _rows


class BilayerSonophore:
    ''' Geometry and resting constants of a bilayer sonophore. '''

    T = 309.15       # Temperature (K)                               bls.py:88
    P0 = 1.0e5       # Static pressure in the surrounding fluid (Pa) bls.py:101
    rel_Zmin = -0.49  # bls.py:110

    def __init__(self, a, Cm0, Qm0, embedding_depth=0.0):
        if a <= 0.:
            raise ValueError('Sonophore radius must be positive')
        if Cm0 <= 0.:
            raise ValueError('Resting membrane capacitance must be positive')
        if embedding_depth < 0.:
            raise ValueError('Embedding depth cannot be negative')
        self.a = a
        self.Cm0 = Cm0
        self.Qm0 = Qm0
        self.d = embedding_depth
        self.S0 = np.pi * self.a**2
        self.computePMparams()
        self.V0 = np.pi * self.Delta * self.a**2
        self.ng0 = self.P0 * self.V0 / (Rg * self.T)

    def computePMparams(self):
        ''' Equilibrium gap and Lennard-Jones fit of the intermolecular pressure, read from the
            table of precomputed values (same keys as bls.py:49-75). '''
        akey = f'{self.a * 1e9:.1f}'
        Qkey = f'{self.Qm0 * 1e5:.2f}'
        try:
            Delta, x0, C, nrep, nattr = _table()[(akey, Qkey)]
        except KeyError:
            raise ValueError(
                f'no precomputed intermolecular-pressure fit for a = {akey} nm, '
                f'Qm0 = {Qkey} nC/cm2 (available in {_TABLE})')
        self.Delta = Delta
        self.LJ_approx = {'x0': x0, 'C': C, 'nrep': nrep, 'nattr': nattr}

    @property
    def Zmin(self):
        return self.rel_Zmin * self.Delta

    def _zprofiles(self, f, A, Qm, relcm, device=0):
        ''' Last-cycle profiles for arrays of (f, A, Qm): Z (m) or Cm / Cm0, [n, 1000]. '''
        from . import _lib
        f, A, Qm = np.broadcast_arrays(np.asarray(f, float), np.asarray(A, float), np.asarray(Qm, float))
        n = f.size
        # the averaging stage of the plan needs a neuron; its tables are simply not fetched here
        plan = _lib.Plan(device, [self.abi_params()], 0, 8, np.zeros(n, np.int32), f.ravel(), A.ravel(),
                         Qm.ravel(), np.array([1.0]))
        try:
            plan.launch()
            return plan.fetch_relcm() if relcm else plan.fetch_zprofiles()
        finally:
            plan.destroy()

    def getZlast(self, drive, Qm):
        ''' Deflection vector (m) of the last acoustic cycle (bls.py:801-803). '''
        return self._zprofiles(drive.f, drive.A, Qm, relcm=False)[0]

    def getRelCmCycle(self, drive, Qm):
        ''' Relative capacitance vector of the last acoustic cycle (bls.py:806-808). '''
        return self._zprofiles(drive.f, drive.A, Qm, relcm=True)[0]

    @property
    def Cm_lkp_filename(self):
        return f'Cm_lkp_{self.a * 1e9:.0f}nm.pkl'        # bls.py:810-812

    def __repr__(self):
        return f'{self.__class__.__name__}({self.a * 1e9:.1f} nm)'

    def abi_params(self):
        ''' dict matching the SonicBlsParams struct of the C ABI. '''
        return {'a': self.a, 'Delta': self.Delta, 'Cm0': self.Cm0, 'depth': self.d, **self.LJ_approx}
