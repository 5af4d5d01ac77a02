# -*- coding: utf-8 -*-
''' GPU box: build the BASELINE config-1 table with the engine and write it to
    gpurun_out/RS_lookups_c1_engine.pkl (input of tools/spike_parity.py). '''
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pysonic_b200 as ps  # noqa: E402

w = bench.workload('c1')
lkp = ps.computeAStimLookup(ps.getPointNeuron('RS'), w['a'], w['f'], w['A'], w['fs'], w['Q'])
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
lkp.toPickle(os.path.join(ROOT, 'gpurun_out', 'RS_lookups_c1_engine.pkl'))
print(lkp)
