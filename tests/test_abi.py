# -*- coding: utf-8 -*-
''' The C-ABI shared library: loads, exports every symbol that include/sonic_b200.h declares,
    and refuses to compute without a device (no CPU fallback). '''

import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def declared_symbols():
    with open(os.path.join(ROOT, 'include', 'sonic_b200.h')) as fh:
        text = fh.read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(sonic_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported(build):
    from pysonic_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in include/sonic_b200.h but not exported'
    # and the binding declares a prototype for each of them
    assert sorted(_lib.EXPORTS.keys()) == syms


def test_struct_layouts(build):
    from pysonic_b200 import _lib
    assert ctypes.sizeof(_lib.SonicBlsParams) == 8 * 8
    assert ctypes.sizeof(_lib.SonicStats) == 10 * 8


def test_neuron_registry(build):
    from pysonic_b200 import _lib
    from pysonic_b200.neurons import NEURON_ORDER, spec_rate_names
    lib = _lib.load()
    assert lib.sonic_version() == 2
    assert lib.sonic_neuron_count() == len(NEURON_ORDER)
    for i, name in enumerate(NEURON_ORDER):
        assert lib.sonic_neuron_id(name.encode()) == i
        assert _lib.neuron_rate_names(i) == spec_rate_names(name)
    assert lib.sonic_neuron_id(b'nope') == -1
    with pytest.raises(_lib.SonicError):
        _lib.check(lib.sonic_neuron_nrates(99))


def test_no_cpu_fallback(build):
    ''' Without a CUDA device every compute entry point fails loudly. '''
    from pysonic_b200 import _lib
    import pysonic_b200 as ps
    if _lib.device_count() > 0:
        pytest.skip('a CUDA device is present')
    nbls = ps.NeuronalBilayerSonophore(32e-9, ps.getPointNeuron('RS'))
    with pytest.raises(_lib.SonicError, match='no CUDA device'):
        nbls.computeEffVars(ps.AcousticDrive(500e3, 100e3), 1.0, -71.9e-5)
    with pytest.raises(_lib.SonicError, match='no CUDA device'):
        ps.computeAStimLookup(ps.getPointNeuron('RS'), np.array([32e-9]), np.array([500e3]),
                              np.array([1e5]), np.array([1.0]), np.array([0.0]))
    with pytest.raises(_lib.SonicError):
        _lib.fp64_peak(0)


def test_product_does_not_import_oracle():
    ''' The package must never route through oracle/ or the CPU harness. '''
    pkg = os.path.join(ROOT, 'pysonic_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith('.py'):
                with open(os.path.join(dirpath, fn)) as fh:
                    text = fh.read()
                for banned in ('sonic_oracle', 'hostsim', 'import scipy', 'from scipy.integrate', 'odeint', 'oracle/'):
                    assert banned not in text, (fn, banned)
