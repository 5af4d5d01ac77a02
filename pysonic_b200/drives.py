# -*- coding: utf-8 -*-
''' Acoustic drive object (mirror of PySONIC/core/drives.py:191-304, lookup-relevant part). '''

import numpy as np

from .constants import NPC_DENSE


class AcousticDrive:
    ''' Acoustic drive with carrier frequency f (Hz), peak pressure amplitude A (Pa), phase. '''

    def __init__(self, f, A=None, phi=np.pi):
        self.f = f
        self.A = A
        self.phi = phi

    @staticmethod
    def _check_float(key, value):
        # stimobj.py:99-104: ints are cast, anything else must be float
        if isinstance(value, (int, np.integer)) and not isinstance(value, bool):
            value = float(value)
        if not isinstance(value, float):
            raise TypeError(f'Invalid {key} (must be float typed)')
        return float(value)

    @property
    def f(self):
        return self._f

    @f.setter
    def f(self, value):
        value = self._check_float('f', value)
        if value <= 0:
            raise ValueError('Invalid f (must be strictly positive)')
        self._f = value

    @property
    def A(self):
        return self._A

    @A.setter
    def A(self, value):
        if value is not None:
            value = self._check_float('A', value)
            if value < 0:
                raise ValueError('Invalid A (must be positive or null)')
        self._A = value

    @property
    def phi(self):
        return self._phi

    @phi.setter
    def phi(self, value):
        self._phi = self._check_float('phi', value)

    def __repr__(self):
        return f'AcousticDrive({self.f * 1e-3:.1f}kHz, {self.A * 1e-3:.2f}kPa)'

    @property
    def desc(self):
        return f'f = {self.f * 1e-3:g} kHz, A = {self.A * 1e-3:g} kPa'

    def copy(self):
        return self.__class__(self.f, self.A, phi=self.phi)

    @property
    def dt(self):
        return 1 / (NPC_DENSE * self.f)

    @property
    def periodicity(self):
        return 1. / self.f

    @property
    def nPerCycle(self):
        return NPC_DENSE

    def compute(self, t):
        return self.A * np.sin(2 * np.pi * self.f * t - self.phi)

    @classmethod
    def createQueue(cls, freqs, amps):
        ''' Drives for all (f, A) combinations, f outer and A inner
            (drives.py:28-34 -> batches.py:155-171). '''
        return [cls(float(f), float(A)) for f in freqs for A in amps]
