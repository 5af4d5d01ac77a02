# -*- coding: utf-8 -*-
import ctypes
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def build():
    ''' Make sure every native artefact exists (no-op when already built). '''
    import __graft_entry__ as ge
    ge.build()
    return ge


@pytest.fixture(scope='session')
def points_golden():
    with open(os.path.join(GOLDEN, 'points.json')) as fh:
        return json.load(fh)


@pytest.fixture(scope='session')
def rates_golden():
    with open(os.path.join(GOLDEN, 'rates_sweep.json')) as fh:
        return json.load(fh)


def load_grid(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope='session')
def gpu(build):
    ''' The native library with a usable device; GPU tests fail loudly otherwise. '''
    from pysonic_b200 import _lib
    lib = _lib.load()
    if lib.sonic_device_count() < 1:
        pytest.fail('no CUDA device: GPU tests must run on the GPU box (pytest -m gpu)')
    return _lib


class HostSim:
    ''' ctypes view of the CPU build of the lane state machine (tests/hostsim). '''

    def __init__(self, path):
        self.lib = ctypes.CDLL(path)
        self.dp = ctypes.POINTER(ctypes.c_double)
        self.lib.hostsim_point.restype = ctypes.c_long
        self.lib.hostsim_z0.restype = ctypes.c_double

    def _bls(self, b):
        return np.array([b.a, b.Delta, b.x0, b.C, b.nrep, b.nattr, b.Cm0, 0.0])

    def set_driver(self, d):
        ''' 0 = staged tick (wide warps), 1 = nested tick of a lone lane, 2 = lone lane with register-resident
            BDF runs (sonic_lone_advance) '''
        self.lib.hostsim_set_driver(ctypes.c_int(d))

    def point(self, b, f, A, Q, trace=False, overtones=None):
        ''' overtones: list of (amplitude C/m2, phase rad) pairs (nbls.py:169-178) '''
        if overtones:
            return self._point_ov(b, f, A, Q, overtones)
        bls = self._bls(b)
        z, ng = np.zeros(1000), np.zeros(1000)
        ncyc, st = ctypes.c_int(), ctypes.c_uint()
        stats = (ctypes.c_uint * 3)()
        tr = np.zeros((11 * 999, 8)) if trace else None
        nrows = ctypes.c_long()
        self.lib.hostsim_point(
            bls.ctypes.data_as(self.dp), ctypes.c_double(f), ctypes.c_double(A), ctypes.c_double(Q),
            z.ctypes.data_as(self.dp), ng.ctypes.data_as(self.dp), ctypes.byref(ncyc), ctypes.byref(st),
            stats, tr.ctypes.data_as(self.dp) if trace else None,
            ctypes.c_long(tr.shape[0] if trace else 0), ctypes.byref(nrows))
        return {'z': z, 'ng': ng, 'ncycles': ncyc.value, 'status': st.value, 'nfe': stats[0],
                'nje': stats[1], 'nsteps': stats[2], 'trace': tr[:nrows.value] if trace else None}

    def _point_ov(self, b, f, A, Q, overtones):
        bls = self._bls(b)
        ov = np.ascontiguousarray(np.asarray(overtones, float).reshape(-1, 2))
        z, ng = np.zeros(1000), np.zeros(1000)
        ncyc, st = ctypes.c_int(), ctypes.c_uint()
        stats = (ctypes.c_uint * 3)()
        nrows = ctypes.c_long()
        self.lib.hostsim_point_ov.restype = ctypes.c_long
        self.lib.hostsim_point_ov(
            bls.ctypes.data_as(self.dp), ctypes.c_double(f), ctypes.c_double(A), ctypes.c_double(Q),
            ctypes.c_int(ov.shape[0]), ov.ctypes.data_as(self.dp), z.ctypes.data_as(self.dp),
            ng.ctypes.data_as(self.dp), ctypes.byref(ncyc), ctypes.byref(st), stats, None, ctypes.c_long(0),
            ctypes.byref(nrows))
        return {'z': z, 'ng': ng, 'ncycles': ncyc.value, 'status': st.value, 'nfe': stats[0],
                'nje': stats[1], 'nsteps': stats[2], 'trace': None}

    def charge_cycle(self, q0, overtones):
        ov = np.ascontiguousarray(np.asarray(overtones, float).reshape(-1, 2))
        self.lib.hostsim_charge_sample.restype = ctypes.c_double
        return np.array([self.lib.hostsim_charge_sample(ctypes.c_double(q0), ctypes.c_int(ov.shape[0]),
                                                        ov.ctypes.data_as(self.dp), ctypes.c_int(j))
                         for j in range(1000)])

    def rhs(self, b, f, A, Q, t, y):
        bls = self._bls(b)
        y = np.ascontiguousarray(y, dtype=float)
        out = np.zeros(3)
        self.lib.hostsim_rhs(bls.ctypes.data_as(self.dp), ctypes.c_double(f), ctypes.c_double(A),
                             ctypes.c_double(Q), ctypes.c_double(t), y.ctypes.data_as(self.dp),
                             out.ctypes.data_as(self.dp))
        return out

    def z0(self, b, f, A, Q):
        bls = self._bls(b)
        return self.lib.hostsim_z0(bls.ctypes.data_as(self.dp), ctypes.c_double(f),
                                   ctypes.c_double(A), ctypes.c_double(Q))

    def tables(self):
        self.lib.hostsim_tables_size.restype = ctypes.c_long
        buf = np.zeros(self.lib.hostsim_tables_size())
        self.lib.hostsim_tables(buf.ctypes.data_as(self.dp))
        return {'elco': buf[:312].reshape(2, 12, 13), 'tesco': buf[312:384].reshape(2, 12, 3),
                'cm1': buf[384:396], 'cm2': buf[396:401], 'sm1': buf[401:413]}


@pytest.fixture(scope='session')
def hostsim(build):
    return HostSim(build.HOSTSIM_LIB)
