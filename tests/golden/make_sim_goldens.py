# -*- coding: utf-8 -*-
''' Golden vectors for the SONIC simulation that consumes the lookup tables: the UNMODIFIED
    reference's NeuronalBilayerSonophore.simulate(drive, pp, fs=1., method='sonic') (nbls.py:389-437,513-536;
    effDerivatives :280-315; Lookup.project / interpolate1D, lookups.py:259-333) run on tables built by the
    reference itself (the grid fixtures of this directory written back as lookup pickles).  Build container only.

        python tests/golden/make_sim_goldens.py [--tables-only | --ulp NEURON ...]
'''
import json
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _refshim import load_reference  # noqa: E402

load_reference()
import PySONIC.core.nbls as ref_nbls  # noqa: E402
from PySONIC.core import NeuronalBilayerSonophore, AcousticDrive, PulsedProtocol  # noqa: E402
from PySONIC.neurons import getPointNeuron  # noqa: E402

import make_goldens as mg  # noqa: E402  (reference-built tables through the same per-point runner)

RS_RUNS = [(20e3, 100e-3, 50e-3, 100., 1.), (50e3, 100e-3, 50e-3, 100., 1.), (100e3, 100e-3, 50e-3, 100., 1.),
           (300e3, 100e-3, 50e-3, 100., 1.), (600e3, 100e-3, 50e-3, 100., 1.), (100e3, 100e-3, 50e-3, 100., 0.5),
           (250e3, 60e-3, 20e-3, 1000., 0.3), (37.5e3, 150e-3, 30e-3, 10., 0.8)]
STD_RUNS = [(50e3, 100e-3, 50e-3, 100., 1.), (100e3, 100e-3, 50e-3, 100., 1.), (300e3, 100e-3, 50e-3, 100., 0.5)]
FIBRE_RUNS = [(50e3, 3e-3, 5e-3, 100., 1.), (100e3, 3e-3, 5e-3, 100., 1.)]
CASES = [('c1_RS_32nm_500kHz.npz', 32e-9, 500e3, RS_RUNS)]
# dense-charge tables at one (radius, frequency) for the other neurons whose states are all gates
SIM_AMPS = np.array([0., 50e3, 100e3, 300e3])
for name, runs in [('FS', STD_RUNS), ('LTS', STD_RUNS), ('IB', STD_RUNS), ('RE', STD_RUNS), ('HHseg', STD_RUNS),
                   ('template', STD_RUNS), ('SWnode', FIBRE_RUNS), ('MRGnode', FIBRE_RUNS), ('SUseg', FIBRE_RUNS)]:
    fname = f'sim_tab_{name}.npz'
    if not os.path.isfile(os.path.join(HERE, fname)):
        mg._grid_to_npz(fname, name, [32e-9], [500e3], SIM_AMPS, mg.default_charges(name), [1.0])
    CASES.append((fname, 32e-9, 500e3, runs))
# reference-built tables only (states that are not gates: simulated by the reference alone, for the spike-count
# parity of engine-built tables, tools/spike_parity.py)
for name in ('STN', 'TC'):
    fname = f'sim_tab_{name}.npz'
    if not os.path.isfile(os.path.join(HERE, fname)):
        mg._grid_to_npz(fname, name, [32e-9], [500e3], SIM_AMPS, mg.default_charges(name), [1.0])

# the reference's own reproducibility for the spike-count check: a table re-built with the drive amplitude changed by
# +-2 ulp (python tests/golden/make_sim_goldens.py --ulp RE [TC ...])
if '--ulp' in sys.argv:
    for name in sys.argv[sys.argv.index('--ulp') + 1:]:
        for sc, tag in ((1.0 + 4.440892098500626e-16, '_ulp_up'), (1.0 - 4.440892098500626e-16, '_ulp_dn')):
            fname = f'sim_tab_{name}{tag}.npz'
            if not os.path.isfile(os.path.join(HERE, fname)):
                mg._grid_to_npz(fname, name, [32e-9], [500e3], SIM_AMPS, mg.default_charges(name), [1.0], sc)
    sys.exit(0)
if '--tables-only' in sys.argv:
    sys.exit(0)
out = {'cases': []}
for fixture, a, f, runs in CASES:
    g = np.load(os.path.join(HERE, fixture))
    name = str(g['neuron'])
    pn = getPointNeuron(name)
    d = tempfile.mkdtemp()
    refs = {k: g[k] for k in ('a', 'f', 'A', 'Q', 'fs')}
    tables = {str(k): g['tab_' + str(k)] for k in g['keys']}
    tables['tcomp'] = np.moveaxis(np.array([g['tcomp']]), 0, -1)
    nbls = NeuronalBilayerSonophore(a, pn)
    with open(os.path.join(d, nbls.getLookupFileName(fs=1.0)), 'wb') as fh:
        pickle.dump({'refs': refs, 'tables': tables}, fh)
    ref_nbls.LOOKUP_DIR = d
    for A, tstim, toffset, PRF, DC in runs:
        pp = PulsedProtocol(tstim, toffset, PRF=PRF, DC=DC)
        try:
            data, meta = nbls.simulate(AcousticDrive(f, A), pp, fs=1.0, method='sonic')
        except ValueError as e:          # the charge left the tabulated range
            out['cases'].append({'fixture': fixture, 'neuron': name, 'a': a, 'f': f, 'A': A, 'tstim': tstim, 'toffset': toffset,
                                 'PRF': PRF, 'DC': DC, 'error': str(e)})
            print(name, A, 'ValueError', e, flush=True)
            continue
        nspikes = int(nbls.pneuron.getNSpikes(data))
        step = max(1, len(data) // 400)
        cols = [c for c in data.columns if c not in ('Z', 'ng')]
        rec = {'fixture': fixture, 'neuron': name, 'a': a, 'f': f, 'A': A, 'tstim': tstim, 'toffset': toffset, 'PRF': PRF,
               'DC': DC, 'nsamples': int(len(data)), 'nspikes': nspikes, 'columns': cols, 'step': step,
               't_all_head': data['t'].values[:8].tolist(), 't_last': float(data['t'].values[-1]),
               'stim_sum': float(data['stimstate'].values.sum()),
               'samples': {c: data[c].values[::step].tolist() for c in cols},
               'final': {c: float(data[c].values[-1]) for c in cols}}
        out['cases'].append(rec)
        print(name, A, pp, 'samples', len(data), 'spikes', nspikes, flush=True)
    shutil.rmtree(d, ignore_errors=True)
with open(os.path.join(HERE, 'sonic_sims.json'), 'w') as fh:
    json.dump(out, fh)
