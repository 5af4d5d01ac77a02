# -*- coding: utf-8 -*-
''' The CPU oracle (oracle/sonic_oracle.py) against golden vectors produced by the unmodified
    reference (tests/golden/make_goldens.py).  With the same numpy / scipy the restatement is
    bit-compatible, so the tolerance is 1e-12 relative. '''

import numpy as np
import pytest

import sonic_oracle as so
from conftest import load_grid

TOL = 1e-12

# a subset of tests/golden/points.json that runs in a few seconds (the 20 kHz points cost 5-10 s)
FAST = lambda p: p['tcomp'] < 1.5   # noqa: E731


def test_neuron_tables_match_reference(points_golden):
    for key, c in points_golden['consts'].items():
        name, a = key.split('@')
        a = float(a[:-2]) * 1e-9
        b = so.get_bls(name, a)
        assert so.rate_names(name) == c['rates']
        assert b.Delta == c['Delta'] and b.ng0 == c['ng0'] and b.V0 == c['V0']
        assert b.Zmin == c['Zmin'] and b.S0 == c['S0']
        assert np.allclose(so.neuron_Qbounds(name), c['Qbounds'], rtol=0, atol=0)
        assert so.neuron_Qm0(name) == c['Qm0']


def test_initial_deflection(points_golden):
    for r in points_golden['Z0']:
        b = so.get_bls(r['neuron'], r['a'])
        z0 = so.balancedef_qs(b, b.ng0, r['Q'], so.pac(r['f'], r['A'], 1 / (1000 * r['f'])))
        assert z0 == pytest.approx(r['Z0'], rel=TOL)


def test_rate_functions(rates_golden):
    Vm = np.array(rates_golden['Vm'])
    for name, rec in rates_golden['neurons'].items():
        assert so.rate_names(name) == rec['rates']
        with np.errstate(all='ignore'):
            for k, fn in so.NEURONS[name][2]:
                mine = np.asarray(fn(Vm), float)
                ref = np.array(rec['values'][k], float)
                both_nan = np.isnan(mine) & np.isnan(ref)
                ok = both_nan | (np.abs(mine - ref) <= 1e-13 * np.abs(ref)) | (mine == ref)
                assert ok.all(), (name, k, Vm[~ok][:5], mine[~ok][:5], ref[~ok][:5])


def test_effective_variables_points(points_golden):
    pts = [p for p in points_golden['points'] if FAST(p)]
    assert len(pts) >= 20
    for p in pts:
        b = so.get_bls(p['neuron'], p['a'])
        ev, ncyc = so.compute_effvars(p['neuron'], b, p['f'], p['A'], p['fs'], p['Q'])
        assert ncyc == p['ncycles'], p
        assert len(ev) == len(p['effvars'])
        for mine, ref in zip(ev, p['effvars']):
            assert list(mine.keys()) == list(ref.keys())
            for k in ref:
                assert mine[k] == pytest.approx(ref[k], rel=TOL, abs=1e-300), (p['neuron'], p['f'], p['A'], k)


def test_lookup_grid_layout_and_values():
    ''' compute_astim_lookup against rows of the reference-built config-1 table. '''
    g = load_grid('c1_RS_32nm_500kHz.npz')
    iA, iQ = [0, 7, 15], [0, 24, 49]
    refs, tables, ncyc = so.compute_astim_lookup('RS', g['a'], g['f'], g['A'][iA], g['fs'], g['Q'][iQ])
    assert list(refs.keys()) == ['a', 'f', 'A', 'Q', 'fs']
    keys = [str(k) for k in g['keys']]
    assert list(tables.keys()) == keys + ['tcomp']
    for k in keys:
        ref = g['tab_' + k][:, :, iA][:, :, :, iQ]
        assert tables[k].shape == (1, 1, 3, 3, 1)
        np.testing.assert_allclose(tables[k], ref, rtol=TOL, atol=0)
    np.testing.assert_array_equal(ncyc, g['ncycles'][:, :, iA][:, :, :, iQ])


def test_relative_capacitance_profiles():
    ''' rel_cm_cycle (bls.py:801-808) against profiles produced by the reference's
        BilayerSonophore(32 nm, 1e-2, 0).getRelCmCycle, the per-point function of run_Cm_lookups.py. '''
    g = load_grid('cm_lkp_32nm_sub.npz')
    b = so.BlsConsts.from_table(32e-9, 1e-2, 0.0)
    for i in (1, 2):                      # 500 kHz and 4 MHz (the 100 kHz profiles take seconds each)
        for j in range(g['A'].size):
            prof = so.rel_cm_cycle(b, float(g['f'][i]), float(g['A'][j]), 0.)
            np.testing.assert_allclose(prof, g['Cm_rel'][i, j], rtol=TOL, atol=0)
    refs, tables = so.compute_cm_lookup(b, g['f'][2:], g['A'][:2])
    assert list(refs) == ['f', 'A', 't'] and tables['Cm_rel'].shape == (1, 2, 1000)
    np.testing.assert_array_equal(refs['t'], g['t'])


def test_charge_overtones_points():
    ''' compute_effvars(..., Qm_overtones) (nbls.py:169-201, bls.py:766-768) against the reference:
        two of the golden points (each costs seconds: the staircase charge makes LSODA restart
        a thousand times per cycle). '''
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'overtones.json')) as fh:
        pts = json.load(fh)['points']
    b = so.get_bls('RS', 32e-9)
    for p in (pts[4], pts[-1]):
        ev, ncyc = so.compute_effvars('RS', b, p['f'], p['A'], p['fs'], p['Q'],
                                      Qm_overtones=[tuple(x) for x in p['overtones']])
        assert ncyc == p['ncycles']
        for mine, ref in zip(ev, p['effvars']):
            assert list(mine.keys()) == list(ref.keys())
            for k in ref:
                assert mine[k] == pytest.approx(ref[k], rel=TOL, abs=1e-300), k
