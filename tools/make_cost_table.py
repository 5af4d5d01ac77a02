# -*- coding: utf-8 -*-
''' Build the work-queue cost table from a measured run of the RS 4-D grid.

        python tools/make_cost_table.py gpurun_out/diag_<tag>.npz [more diag files ...]

    Input: per-point right-hand-side counts of BASELINE config 2 (tools/gpu_diag.py on the GPU
    box).  Output: log(cost) on a coarse (radius, frequency, amplitude bin, |charge| bin) grid,
    each node holding the maximum over its bin (an envelope: a point predicted cheap that turns
    out to be expensive costs wall time, the opposite costs nothing), written to
      pysonic_b200/data/cost_table.json            (read by pysonic_b200/parallel.py)
      pysonic_b200/csrc/generated/cost_table.h     (compiled into libsonic_b200.so)
    The mechanics depend on the charge through Q^2 only (bls.py:482-491), hence the |Q| axis.
    The table only orders the work queue and sizes the lane budgets; it never touches results. '''
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')

A_EDGES_KPA = [0., 0.05, 0.3, 0.7, 1.5, 3., 5., 8., 12., 20., 35., 60., 100., 140., 190., 250., 320., 400., 500.,
               1e9]
Q_EDGES = [0., 10., 20., 30., 40., 50., 60., 70., 80., 85., 89., 92., 95., 98., 101., 104., 110., 125., 140.,
           160., 180., 200., 225., 250., 275., 1e9]   # |Q| in nC/cm2


def main():
    a_nodes = np.array([16e-9, 32e-9, 64e-9])
    f_nodes = None
    nA, nQ = len(A_EDGES_KPA) - 1, len(Q_EDGES) - 1
    tab = None
    parts = []
    for path in sys.argv[1:]:
        d = np.load(path)
        if f_nodes is None:
            f_nodes = np.unique(d['f'])
            tab = np.zeros((a_nodes.size, f_nodes.size, nA, nQ))
        # radius index of the file's own grid -> index in a_nodes (files of one radius hold 32 nm)
        ia = d['ia'] if d['ia'].max() > 0 else np.full(d['ia'].shape, 1)
        f, A, Q, nr = d['f'], d['A'], d['Q'], d['nrhs'].astype(float)
        jA = np.digitize(A * 1e-3, A_EDGES_KPA) - 1
        jQ = np.digitize(np.abs(Q) * 1e5 + 1e-9, Q_EDGES) - 1
        jf = np.searchsorted(f_nodes, f)
        np.maximum.at(tab, (ia, jf, jA, jQ), nr)
        parts.append((ia, jf, jA, jQ, nr))
    have = tab > 0
    with np.errstate(divide='ignore'):
        tab = np.log(tab)
    # bins without data (large |Q| at 16 / 64 nm): the 32 nm value of the same (f, A, |Q|) bin shifted
    # by the radius offset seen at the largest |Q| bin that has data for that radius
    for i in (0, 2):
        for j in range(f_nodes.size):
            for k in range(nA):
                lq = np.where(have[i, j, k])[0].max()
                off = tab[i, j, k, lq] - tab[1, j, k, lq]
                for l in range(nQ):
                    if not have[i, j, k, l]:
                        tab[i, j, k, l] = tab[1, j, k, l] + off
    assert np.isfinite(tab).all(), 'empty bin'
    out = {'a': a_nodes.tolist(), 'f': f_nodes.tolist(), 'A_edges_kPa': A_EDGES_KPA, 'absQ_edges_nCcm2': Q_EDGES,
           'log_cost': np.round(tab, 3).tolist(), 'source': [os.path.basename(x) for x in sys.argv[1:]]}
    with open(os.path.join(ROOT, 'pysonic_b200', 'data', 'cost_table.json'), 'w') as fh:
        json.dump(out, fh)
    with open(os.path.join(ROOT, 'pysonic_b200', 'csrc', 'generated', 'cost_table.h'), 'w') as fh:
        fh.write('// GENERATED FILE -- do not edit.  Produced by tools/make_cost_table.py from a measured run of the\n'
                 '// RS 4-D grid: log(right-hand-side evaluations) per (radius, frequency, amplitude bin, |charge| bin),\n'
                 '// maximum over each bin.  Only used to order the work queue and to size the lane budgets.\n#pragma once\n')
        fh.write(f'#define SONIC_COST_NA {a_nodes.size}\n#define SONIC_COST_NF {f_nodes.size}\n'
                 f'#define SONIC_COST_NAMP {nA}\n#define SONIC_COST_NQ {nQ}\n')
        fh.write('static const double SONIC_COST_A[] = {' + ', '.join(repr(float(x)) for x in a_nodes) + '};\n')
        fh.write('static const double SONIC_COST_F[] = {' + ', '.join(repr(float(x)) for x in f_nodes) + '};\n')
        fh.write('static const double SONIC_COST_AMP_EDGES[] = {' + ', '.join(repr(float(x) * 1e3) for x in A_EDGES_KPA) + '};\n')
        fh.write('static const double SONIC_COST_Q_EDGES[] = {' + ', '.join(repr(float(x) * 1e-5) for x in Q_EDGES) + '};\n')
        fh.write('static const float SONIC_COST_LOG[] = {\n')
        flat = tab.ravel()
        for i in range(0, flat.size, 12):
            fh.write('    ' + ', '.join(f'{v:.3f}f' for v in flat[i:i + 12]) + ',\n')
        fh.write('};\n')
    # quality of the envelope on the source data
    for path, (ia, jf, jA, jQ, nr) in zip(sys.argv[1:], parts):
        r = np.log(nr) - tab[ia, jf, jA, jQ]
        print(os.path.basename(path), 'mean over-prediction (log)', float(-r.mean()), 'max under-prediction', float(r.max()))
    print('table', tab.shape)


if __name__ == '__main__':
    main()
