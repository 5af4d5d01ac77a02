# -*- coding: utf-8 -*-
''' Development check on a GPU box: FP64 peak, smoke, C1 parity against the golden grid, C2 timing. '''
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps  # noqa: E402
from pysonic_b200 import _lib  # noqa: E402


def parity(lkp, g, keys):
    worst = {}
    allerr = []
    for k in keys:
        r = g['tab_' + k]
        d = np.abs(lkp[k] - r)
        e = d / np.maximum(np.abs(r), 1e-300)
        e[d < 1e-9] = 0
        worst[k] = float(e.max())
        allerr.append(e)
    return worst, np.max(np.array(allerr), axis=0)


def main():
    what = sys.argv[1:] or ['peak', 'smoke', 'c1', 'c2']
    print('devices', _lib.device_count())
    if 'peak' in what:
        print('fp64 peak TFLOP/s', _lib.fp64_peak(0))
    if 'smoke' in what:
        import __graft_entry__ as ge
        ge.smoke()
    if 'c1' in what:
        g = np.load(os.path.join(ROOT, 'tests/golden/c1_RS_32nm_500kHz.npz'))
        pn = ps.getPointNeuron('RS')
        for rep in range(2):
            t0 = time.perf_counter()
            lkp, info = ps.computeAStimLookup(pn, g['a'], g['f'], g['A'], g['fs'], g['Q'], return_info=True)
            dt = time.perf_counter() - t0
            print('C1 run', rep, 'wall %.3f s' % dt, info['stats'])
        keys = [str(k) for k in g['keys']]
        worst, err = parity(lkp, g, keys)
        print('C1 worst rel err per table', worst)
        print('C1 frac points > 1e-4:', float(np.mean(err > 1e-4)), 'median', float(np.median(err)))
        same = info['ncycles'] == g['ncycles']
        print('C1 ncycles identical', float(same.mean()), 'for A>=10kPa', float(same[:, :, g['A'] >= 1e4].mean()))
        print('C1 status counts', np.unique(info['status'], return_counts=True))
        print('C1 nfe ratio gpu/ref', info['stats']['n_rhs'] / g['nfe'].sum())
    if 'c2' in what:
        pn = ps.getPointNeuron('RS')
        a = np.array([16e-9, 32e-9, 64e-9])
        f = np.array([20., 100., 500., 1e3, 2e3, 3e3, 4e3]) * 1e3
        A = np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3
        Qmin, Qmax = pn.Qbounds
        Q = np.arange(Qmin, Qmax + 1e-5, 1e-5)
        sub = os.environ.get('C2_SUB')
        if sub:
            s = int(sub)
            A, Q = A[::s], Q[::s]
        t0 = time.perf_counter()
        lkp, info = ps.computeAStimLookup(pn, a, f, A, np.array([1.0]), Q, return_info=True)
        dt = time.perf_counter() - t0
        n = a.size * f.size * A.size * Q.size
        print('C2 points', n, 'wall %.3f s' % dt, '-> %.1f points/s' % (n / dt), info['stats'])
        print('C2 status counts', np.unique(info['status'], return_counts=True))
        print('C2 ncycles hist', np.bincount(info['ncycles'].ravel()))
        print('C2 finite', all(np.isfinite(v).all() for v in lkp.tables.values()))
        tp = lkp['tcomp'][..., 0]
        print('C2 tpoint max %.3f s, sum %.1f s' % (tp.max(), tp.sum()))
        gs = os.path.join(ROOT, 'tests/golden/c2_RS_sub.npz')
        if os.path.isfile(gs) and not sub:
            g = np.load(gs)
            iA = [int(np.argmin(np.abs(A - x))) for x in g['A']]
            iQ = [int(np.argmin(np.abs(Q - x))) for x in g['Q']]
            keys = [str(k) for k in g['keys']]
            errs = []
            for k in keys:
                mine = lkp[k][:, :, iA][:, :, :, iQ]
                r = g['tab_' + k]
                d = np.abs(mine - r)
                e = d / np.maximum(np.abs(r), 1e-300)
                e[d < 1e-9] = 0
                errs.append(e)
            err = np.max(np.array(errs), axis=0)
            print('C2-sub parity: frac > 1e-4 %.4f, median %.2e, p99 %.2e, max %.2e' % (
                np.mean(err > 1e-4), np.median(err), np.percentile(err, 99), err.max()))
            nc = info['ncycles'][:, :, iA][:, :, :, iQ]
            same = nc == g['ncycles']
            print('C2-sub ncycles identical %.4f; for A>=10kPa %.4f' % (same.mean(), same[:, :, g['A'] >= 1e4].mean()))
            np.savez_compressed(os.path.join(ROOT, 'gpurun_out', 'c2_err.npz'), err=err, nc=nc, ncref=g['ncycles'])


if __name__ == '__main__':
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    main()
