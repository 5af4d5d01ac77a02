# SONIC simulations of the golden cases on the GPU, full traces saved for offline comparison with the reference
import json, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps
gold = os.path.join(ROOT, 'tests', 'golden')
cases = json.load(open(os.path.join(gold, 'sonic_sims.json')))['cases']
out = {}
for i, c in enumerate(cases):
    if 'error' in c:
        continue
    g = np.load(os.path.join(gold, c['fixture']))
    lkp = ps.Lookup({k: g[k] for k in ('a', 'f', 'A', 'Q', 'fs')}, {str(k): g['tab_' + str(k)] for k in g['keys']})
    nbls = ps.NeuronalBilayerSonophore(c['a'], ps.getPointNeuron(c['neuron']))
    pp = ps.PulsedProtocol(c['tstim'], c['toffset'], PRF=c['PRF'], DC=c['DC'])
    data, _ = nbls.simulate(ps.AcousticDrive(c['f'], c['A']), pp, lookup=lkp)
    out[f'case{i}_t'] = data['t'].values; out[f'case{i}_Qm'] = data['Qm'].values; out[f'case{i}_stim'] = data['stimstate'].values
    print(i, c['neuron'], c['A'], 'spikes', nbls.getNSpikes(data), 'ref', c['nspikes'])
np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'sim_traces.npz'), **out)
