# -*- coding: utf-8 -*-
''' Host-side logic of bench.py that needs no GPU: workload definitions, the bounded CPU sample
    and the reference arm's JSON line. '''

import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workloads_match_baseline_configs():
    c1, c2 = bench.workload('c1'), bench.workload('c2')
    assert (c1['a'].size, c1['f'].size, c1['A'].size, c1['Q'].size) == (1, 1, 20, 50)
    assert (c2['a'].size, c2['f'].size, c2['A'].size, c2['Q'].size) == (3, 7, 51, 158)
    assert c2['A'][0] == 0. and np.isclose(c2['A'][-1], 600e3) and np.isclose(c2['A'][1], 100.)
    ia, f, A, Q = bench.flatten(c2)
    assert ia.size == 169218 and ia.dtype == np.int32
    # reference queue order a > f > A > Q (run_lookups.py:98-103)
    assert Q[1] != Q[0] and A[0] == A[157] and A[158] != A[0] and f[51 * 158] != f[0] and ia[7 * 51 * 158] == 1


def test_cpu_sample_is_stratified_and_bounded():
    w = bench.workload('c2')
    jobs, desc = bench.cpu_sample(w, 8, 20.0)
    assert 100 <= len(jobs) <= 400 and 'systematic' in desc
    fs = {j[2] for j in jobs}
    assert fs == set(w['f'].tolist())                   # every frequency stratum is present
    assert {j[1] for j in jobs} == set(w['a'].tolist())
    assert min(j[3] for j in jobs) == 0. and max(j[3] for j in jobs) > 3e5


def test_rates_weight_counts_transcendentals():
    assert 400 < bench.rates_weight('RS') < 900         # ~560 for RS
    assert bench.rates_weight('STN') > bench.rates_weight('RS')


def test_reference_arm_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--workload', 'c1', '--cpu-budget', '1'], capture_output=True,
                         text=True, timeout=300, check=True).stdout.strip().splitlines()[-1]
    line = json.loads(out)
    assert line['impl'] == 'reference' and line['unit'] == 'points/s' and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['e2e']['h2d_bytes_per_step'] == 0 and line['e2e']['value'] == line['value']
    # the other ranks of a torchrun launch exit without work
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                         capture_output=True, text=True, timeout=60, check=True, env=env).stdout
    assert out.strip() == ''
