# -*- coding: utf-8 -*-
''' pysonic_b200 -- B200-native engine for PySONIC's lookup-table generation path.

    Host-side mirror of the reference interfaces on that path (same names, argument meaning and
    error behaviour), backed by hand-written sm_100a CUDA kernels behind a C ABI
    (include/sonic_b200.h, libsonic_b200.so).  No CPU fallback. '''

from .constants import *  # noqa: F401,F403
from .drives import AcousticDrive  # noqa: F401
from .neurons import PointNeuron, getDefaultPassiveNeuron, getPointNeuron, passiveNeuron  # noqa: F401
from .bls import BilayerSonophore  # noqa: F401
from .nbls import NeuronalBilayerSonophore  # noqa: F401
from .batches import Batch  # noqa: F401
from .lookups import Lookup  # noqa: F401
from .protocols import PulsedProtocol  # noqa: F401
from .run_lookups import computeAStimLookup, computeAStimLookups  # noqa: F401
from .run_cm_lookups import computeCmLookup  # noqa: F401

__version__ = '0.1.0'
