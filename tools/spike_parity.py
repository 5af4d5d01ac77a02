# -*- coding: utf-8 -*-
''' Third parity criterion of BASELINE.json: identical spike counts from the reference's own
    `NeuronalBilayerSonophore.simulate(drive, pp, method='sonic')` (nbls.py:513-536 -> :389-437)
    when it runs on an engine-built table instead of a reference-built one.

    Build container only (imports the unmodified reference through tests/golden/_refshim.py).

        python tools/spike_parity.py <engine_table.pkl> [out.json]

    <engine_table.pkl>: BASELINE config 1 (RS, 32 nm, 500 kHz, 20 A x 50 Q, fs = 1) written by
    `pysonic_b200` on the GPU box (tools/gpu_make_c1_table.py), brought back in gpurun_out/.
    The reference-built table of the same grid is tests/golden/c1_RS_32nm_500kHz.npz.
'''
import json
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
from _refshim import load_reference  # noqa: E402

load_reference()
import PySONIC.core.nbls as ref_nbls  # noqa: E402
from PySONIC.core import NeuronalBilayerSonophore, AcousticDrive, PulsedProtocol  # noqa: E402
from PySONIC.neurons import getPointNeuron  # noqa: E402


def spikes_with_table(path, amps):
    d = tempfile.mkdtemp()
    try:
        shutil.copy(path, os.path.join(d, 'RS_lookups_fs1.00.pkl'))
        ref_nbls.LOOKUP_DIR = d
        nbls = NeuronalBilayerSonophore(32e-9, getPointNeuron('RS'))
        pp = PulsedProtocol(100e-3, 50e-3)
        out = []
        for A in amps:
            data, meta = nbls.simulate(AcousticDrive(500e3, A), pp, method='sonic')
            out.append(int(nbls.pneuron.getNSpikes(data)))
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


def golden_table_pickle(path):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'c1_RS_32nm_500kHz.npz'))
    refs = {k: g[k] for k in ('a', 'f', 'A', 'Q', 'fs')}
    tables = {str(k): g['tab_' + str(k)] for k in g['keys']}
    tables['tcomp'] = np.moveaxis(np.array([g['tcomp']]), 0, -1)
    with open(path, 'wb') as fh:
        pickle.dump({'refs': refs, 'tables': tables}, fh)


def main():
    engine = sys.argv[1]
    amps = [20e3, 50e3, 100e3, 200e3, 400e3, 600e3]
    tmp = tempfile.mkdtemp()
    refp = os.path.join(tmp, 'ref.pkl')
    golden_table_pickle(refp)
    n_ref = spikes_with_table(refp, amps)
    n_eng = spikes_with_table(engine, amps)
    res = {'neuron': 'RS', 'a': 32e-9, 'f': 500e3, 'protocol': 'PulsedProtocol(100 ms, 50 ms)',
           'amplitudes_Pa': amps, 'spikes_reference_table': n_ref, 'spikes_engine_table': n_eng,
           'identical': n_ref == n_eng, 'engine_table': os.path.basename(engine)}
    print(json.dumps(res))
    if len(sys.argv) > 2:
        with open(sys.argv[2], 'w') as fh:
            json.dump(res, fh, indent=1)
    shutil.rmtree(tmp, ignore_errors=True)
    sys.exit(0 if res['identical'] else 1)


if __name__ == '__main__':
    main()
