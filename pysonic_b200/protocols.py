# -*- coding: utf-8 -*-
''' Pulsing protocol of a SONIC simulation (mirror of PySONIC/core/protocols.py:224-391, the part
    `NeuronalBilayerSonophore.simulate` uses: transition events and stopping time). '''

import numpy as np


class PulsedProtocol:
    ''' Stimulus of duration `tstim` (s) followed by `toffset` (s) of rest, optionally pulsed at `PRF` (Hz)
        with duty cycle `DC`, starting at `tstart` (s). '''

    def __init__(self, tstim, toffset, PRF=100., DC=1., tstart=0.):
        for key, v in (('tstim', tstim), ('toffset', toffset), ('PRF', PRF), ('tstart', tstart)):
            if isinstance(v, (int, np.integer)) and not isinstance(v, bool):
                v = float(v)
            if not isinstance(v, float):
                raise TypeError(f'Invalid {key} (must be float typed)')
            if v < 0:
                raise ValueError(f'Invalid {key} (must be positive or null)')
        DC = float(DC)
        if not 0. <= DC <= 1.:
            raise ValueError(f'Invalid DC: {DC} (must be within [0.0, 1.0] interval)')
        if DC < 1. and PRF < 1 / tstim:
            raise ValueError(f'Invalid PRF: {PRF} (must be within [{1 / tstim}, inf] interval)')
        self.tstim, self.toffset, self.PRF, self.DC, self.tstart = float(tstim), float(toffset), float(PRF), DC, float(tstart)

    def __repr__(self):
        s = f'{type(self).__name__}(tstim={self.tstim * 1e3:g}ms, toffset={self.toffset * 1e3:g}ms'
        if not self.isCW:
            s += f', PRF={self.PRF:g}Hz, DC={self.DC * 1e2:.1f}%'
        return s + ')'

    @property
    def tstop(self):
        return self.tstim + self.toffset + self.tstart

    @property
    def isCW(self):
        return self.DC == 1.

    @property
    def npulses(self):
        return int(np.round(self.tstim * self.PRF))

    def tOFFON(self):
        ''' Times of the OFF-ON transitions (s), protocols.py:372-377. '''
        if self.isCW:
            return np.array([self.tstart])
        return np.arange(self.npulses) / self.PRF + self.tstart

    def tONOFF(self):
        ''' Times of the ON-OFF transitions (s), protocols.py:379-384. '''
        if self.isCW:
            return np.array([self.tstart + self.tstim])
        return (np.arange(self.npulses) + self.DC) / self.PRF + self.tstart

    def stimEvents(self):
        ''' (time, stimulus state) pairs of every transition, in time order (protocols.py:386-391). '''
        on = [(float(t), 1.) for t in self.tOFFON()]
        off = [(float(t), 0.) for t in self.tONOFF()]
        return sorted(on + off, key=lambda e: e[0])
