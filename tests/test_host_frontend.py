# -*- coding: utf-8 -*-
''' Host-side mirror of the reference interfaces on the lookup path: descriptors, grid
    validation, queue order, file naming, pickle format.  No GPU needed. '''

import logging
import os
import pickle

import numpy as np
import pytest

import pysonic_b200 as ps
from pysonic_b200 import codegen
from pysonic_b200.batches import Batch
from pysonic_b200.neurons import NEURON_ORDER, NEURON_SPECS, spec_rate_names
from pysonic_b200.parallel import predicted_log_cost, shard_indices
from pysonic_b200.run_lookups import _parser, _validate


def test_neuron_descriptors_match_reference(points_golden):
    for key, c in points_golden['consts'].items():
        name, a = key.split('@')
        a = float(a[:-2]) * 1e-9
        pn = ps.getPointNeuron(name)
        assert pn.rates == c['rates']
        assert pn.Cm0 == c['Cm0'] and pn.Qm0 == c['Qm0']
        np.testing.assert_array_equal(pn.Qbounds, c['Qbounds'])
        nbls = ps.NeuronalBilayerSonophore(a, pn)
        assert nbls.Delta == c['Delta'] and nbls.ng0 == c['ng0'] and nbls.V0 == c['V0']
        assert nbls.Zmin == c['Zmin'] and nbls.S0 == c['S0']
        assert nbls.LJ_approx == c['LJ']
        p = nbls.abi_params()
        assert set(p) == {'a', 'Delta', 'x0', 'C', 'nrep', 'nattr', 'Cm0', 'depth'}


def test_rates_golden_neurons_all_declared(rates_golden):
    assert set(rates_golden['neurons']) == set(NEURON_ORDER) - {'pas'}      # the passive membrane has no gate
    for name, rec in rates_golden['neurons'].items():
        assert spec_rate_names(name) == rec['rates']
        assert NEURON_SPECS[name]['Cm0'] == rec['Cm0'] and NEURON_SPECS[name]['Vm0'] == rec['Vm0']


def test_unknown_neuron_and_radius():
    with pytest.raises(ValueError):
        ps.getPointNeuron('XYZ')
    # a radius absent from the parameter table is computed (findDeltaEq + LJfitPMavg with the
    # quadratures on the GPU): without a device that fails loudly, there is no CPU path
    from pysonic_b200 import _lib
    if _lib.device_count() == 0:
        with pytest.raises(_lib.SonicError, match='no CUDA device'):
            ps.NeuronalBilayerSonophore(33.3e-9, ps.getPointNeuron('RS'))
    with pytest.raises(ValueError):
        ps.BilayerSonophore(-1e-9, 1e-2, -7e-4)


def test_acoustic_drive_checks():
    d = ps.AcousticDrive(500e3, 100e3)
    assert d.dt == 1 / (1000 * 500e3) and d.periodicity == 1 / 500e3 and d.nPerCycle == 1000
    assert d.compute(0.) == pytest.approx(100e3 * np.sin(-np.pi))
    with pytest.raises(ValueError):
        ps.AcousticDrive(0., 1e3)
    with pytest.raises(ValueError):
        ps.AcousticDrive(1e3, -1.)
    with pytest.raises(TypeError):
        ps.AcousticDrive('500', 1e3)
    q = ps.AcousticDrive.createQueue([1e5, 2e5], [0., 1e3, 2e3])
    assert [(x.f, x.A) for x in q] == [(1e5, 0.), (1e5, 1e3), (1e5, 2e3), (2e5, 0.), (2e5, 1e3), (2e5, 2e3)]


def test_grid_validation_mirrors_reference():
    ''' run_lookups.py:85-96 and :58-60: same exception types. '''
    pn = ps.getPointNeuron('RS')
    a, f, A, fs, Q = (np.array([32e-9]), np.array([500e3]), np.array([0., 1e5]), np.array([1.]),
                      np.array([-7e-4, 0.]))
    with pytest.raises(TypeError):
        ps.computeAStimLookup(pn, a, f, np.array([0, 100000]), fs, Q)        # int typed
    with pytest.raises(TypeError):
        ps.computeAStimLookup(pn, 32e-9, f, A, fs, Q)                        # not iterable
    with pytest.raises(ValueError):
        ps.computeAStimLookup(pn, a, np.array([]), A, fs, Q)                 # empty
    with pytest.raises(ValueError):
        ps.computeAStimLookup(pn, np.array([-1e-9]), f, A, fs, Q)            # non-positive radius
    with pytest.raises(ValueError):
        ps.computeAStimLookup(pn, a, f, np.array([-1.]), fs, Q)              # negative amplitude
    with pytest.raises(AssertionError):
        ps.computeAStimLookup(pn, a, np.array([1e5, 5e5]), A, np.array([0.5, 1.]), Q)   # fs sweep, 2 f
    with pytest.raises(ValueError):
        ps.computeAStimLookup(pn, a, f, A, fs, Q, novertones=9)                # more overtones than supported
    _validate({'a': list(a), 'f': list(f), 'A': list(A), 'Q': list(Q), 'fs': list(fs)})


def test_lookup_file_names():
    nbls = ps.NeuronalBilayerSonophore(32e-9, ps.getPointNeuron('RS'))
    assert nbls.getLookupFileName() == 'RS_lookups_fs1.00.pkl'
    assert nbls.getLookupFileName(a=32e-9, f=500e3, fs=1.) == 'RS_lookups_32nm_500kHz_fs1.00.pkl'
    assert nbls.getLookupFileName(a=32e-9, f=500e3, A=1e5) == 'RS_lookups_32nm_500kHz_100kPa.pkl'
    assert nbls.getLookupFileName(fs=0.75, novertones=2) == 'RS_lookups_fs0.75_2overtones.pkl'


def test_lookup_pickle_format(tmp_path):
    ''' {'refs': dict(a, f, A, Q, fs), 'tables': dict(V, rates..., tcomp)} of plain ndarrays
        (lookups.py:381-392); readable without this package. '''
    refs = {'a': np.array([32e-9]), 'f': np.array([5e5]), 'A': np.array([0., 1e5]),
            'Q': np.array([-1e-3, 0., 1e-3]), 'fs': np.array([1.])}
    dims = (1, 1, 2, 3, 1)
    tables = {k: np.random.default_rng(0).random(dims) for k in ['V', 'alpham', 'betam', 'tcomp']}
    lkp = ps.Lookup(refs, tables)
    assert repr(lkp) == 'Lookup5D(a: 1, f: 1, A: 2, Q: 3, fs: 1)[V, alpham, betam, tcomp]'
    fpath = os.path.join(tmp_path, 'x.pkl')
    lkp.toPickle(fpath)
    with open(fpath, 'rb') as fh:
        d = pickle.load(fh)
    assert list(d.keys()) == ['refs', 'tables']
    assert type(d['refs']) is dict and type(d['tables']) is dict
    assert list(d['refs'].keys()) == ['a', 'f', 'A', 'Q', 'fs']
    assert list(d['tables'].keys()) == ['V', 'alpham', 'betam', 'tcomp']
    for v in d['tables'].values():
        assert type(v) is np.ndarray and v.dtype == np.float64 and v.shape == dims
    back = ps.Lookup.fromPickle(fpath)
    np.testing.assert_array_equal(back['V'], tables['V'])
    with pytest.raises(ValueError):
        ps.Lookup(refs, {'V': np.zeros((1, 1, 2, 3))})
    with pytest.raises(FileNotFoundError):
        ps.Lookup.fromPickle(os.path.join(tmp_path, 'missing.pkl'))


def test_batch_generic_function_and_queue_order():
    out = Batch(lambda x, y=0: x + y, [[1], ([2], {'y': 5}), [3]])(mpi=False, loglevel=logging.ERROR)
    assert out == [1, 7, 3]
    q = Batch.createQueue([1., 2.], [10., 20., 30.])
    assert q == [[1., 10.], [1., 20.], [1., 30.], [2., 10.], [2., 20.], [2., 30.]]
    q3 = Batch.createQueue([1., 2.], [10., 20.], [5., 6.])
    assert q3[0] == [1., 10., 5.] and q3[1] == [1., 10., 6.] and q3[-1] == [2., 20., 6.]


def test_cli_defaults_match_reference():
    ''' run_lookups.py:183-188 and parsers.py:439-448 '''
    args = _parser().parse_args([])
    assert args.neuron == ['RS'] and args.radius == [16.0, 32.0, 64.0]
    assert args.freq == [20., 100., 500., 1e3, 2e3, 3e3, 4e3]
    assert args.fs == [100.] and not args.spanFs and not args.mpi and not args.test
    args = _parser().parse_args('-n STN -a 32 -f 500 --spanFs --mpi --test'.split())
    assert args.neuron == ['STN'] and args.spanFs and args.mpi and args.test


def test_codegen_is_committed_and_deterministic():
    path = os.path.join(os.path.dirname(codegen.__file__), 'csrc', 'generated', 'neuron_rates.cuh')
    with open(path) as fh:
        assert fh.read() == codegen.generate()
    text = codegen.generate()
    for i, n in enumerate(NEURON_ORDER):
        assert f'template <> struct SonicRates<{i}>' in text and f'// ---- {n} ----' in text


def test_cost_sharding_round_robin():
    a = np.full(8, 32e-9)
    f = np.array([2e4, 5e5, 4e6, 2e4, 5e5, 4e6, 1e5, 1e6])
    A = np.array([6e5, 6e5, 6e5, 0., 1e3, 1e4, 3e5, 1e5])
    cost = predicted_log_cost(a, f, A)
    assert f[np.argmax(cost)] == 2e4 and f[np.argmin(cost)] == 4e6    # cost falls with frequency
    Q = np.array([-1e-3, 0., 0., -1e-3, 0., 0., 0., 0.])
    costq = predicted_log_cost(a, f, A, Q)
    assert np.all(costq <= cost + 1e-12)  # without charges: the envelope over all charges
    assert costq[0] > costq[1] > costq[2]  # 600 kPa: 20 kHz dearer than 500 kHz dearer than 4 MHz
    parts = [shard_indices(cost, r, 3) for r in range(3)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(8))
    order = np.argsort(-cost, kind='stable')
    assert [p[0] for p in parts] == order[:3].tolist()


def test_cm_lookup_validation_and_names():
    ''' run_Cm_lookups.py:19-41 (same exception types) and bls.py:810-812 (file name). '''
    bls = ps.BilayerSonophore(32e-9, 1e-2, 0.0)
    assert bls.Cm_lkp_filename == 'Cm_lkp_32nm.pkl'
    with pytest.raises(TypeError):
        ps.computeCmLookup(bls, 500e3, np.array([0., 1e5]))
    with pytest.raises(TypeError):
        ps.computeCmLookup(bls, np.array([500e3]), np.array([0, 100000]))
    with pytest.raises(ValueError):
        ps.computeCmLookup(bls, np.array([]), np.array([0., 1e5]))
    with pytest.raises(ValueError):
        ps.computeCmLookup(bls, np.array([-1.]), np.array([0., 1e5]))
    with pytest.raises(ValueError):
        ps.computeCmLookup(bls, np.array([5e5]), np.array([-1.]))


# ---------------------------------------------------------------------------------------------
# round 2: grid-shaping rules of the overtone lookup, input validation, trajectory-aware sharding
# ---------------------------------------------------------------------------------------------
def _lookup_refs_golden():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'lookup_refs.json')) as fh:
        return json.load(fh)


@pytest.mark.parametrize('case', ['overtones1_default', 'overtones1_test', 'overtones2_small', 'plain_test', 'fs_span'])
def test_refs_match_reference_front_end(case):
    ''' Keys, order, sizes and values of `refs` against the reference's computeAStimLookup
        (scripts/run_lookups.py:48-128; fixture made by tests/golden/make_lookup_refs.py). '''
    from pysonic_b200.run_lookups import build_refs, overtone_refs
    c = _lookup_refs_golden()[case]
    a = c['args']
    refs = build_refs(np.array(a['a']), np.array(a['f']), np.array(a['A']), np.array(a['fs']), np.array(a['Q']),
                      novertones=a['novertones'], test=a['test'])
    if a['novertones']:
        refs = overtone_refs(refs, a['novertones'], a['test'])
    assert list(refs.keys()) == list(c['refs'].keys())
    for k, v in c['refs'].items():
        np.testing.assert_array_equal(refs[k], np.array(v), err_msg=k)
    assert [x.size for x in refs.values()] == c['shape']


def test_overtone_lookup_refuses_several_radii():
    from pysonic_b200.run_lookups import build_refs
    msg = _lookup_refs_golden()['overtones_multi_a_error']
    A = np.array([0., 1e4])
    with pytest.raises(AssertionError) as e:
        build_refs(np.array([16e-9, 32e-9]), np.array([500e3]), A, np.array([1.0]), np.array([0.]), novertones=1)
    assert str(e.value) == msg
    with pytest.raises(AssertionError):
        build_refs(np.array([32e-9]), np.array([500e3, 1e6]), A, np.array([1.0]), np.array([0.]), novertones=1)
    with pytest.raises(AssertionError):
        build_refs(np.array([32e-9]), np.array([500e3]), A, np.array([0.5, 1.0]), np.array([0.]), novertones=1)


def test_charge_range_and_phase_checks():
    ''' bls.py:674-677: charges (and the extrema of Fourier-series charge cycles) outside
        CHARGE_RANGE raise ValueError; a non-default drive phase is refused, not ignored. '''
    from pysonic_b200.nbls import check_charges, check_drive_phase
    check_charges(np.array([-300e-5, 0., 150e-5]))
    with pytest.raises(ValueError, match='Invalid applied charge'):
        check_charges(np.array([0., -301e-5]))
    with pytest.raises(ValueError, match='Invalid applied charge'):
        check_charges(151e-5)
    # cycle Q0 + 2 A cos(.): -107 - 2 x 100 nC/cm2 leaves the range, -107 - 2 x 50 does not
    check_charges(np.array([-107e-5]), np.array([[[50e-5, 0.]]]))
    with pytest.raises(ValueError, match='Invalid applied charge'):
        check_charges(np.array([-107e-5]), np.array([[[100e-5, 0.]]]))
    check_drive_phase(ps.AcousticDrive(500e3, 1e5))
    with pytest.raises(ValueError, match='phase'):
        check_drive_phase(ps.AcousticDrive(500e3, 1e5, phi=0.5))
    nbls = ps.NeuronalBilayerSonophore(32e-9, ps.getPointNeuron('RS'))
    with pytest.raises(ValueError, match='phase'):
        nbls.computeEffVars(ps.AcousticDrive(500e3, 1e5, phi=0.), 1.0, -71.9e-5)


def test_sharding_keeps_both_signs_of_a_charge_together():
    from pysonic_b200.parallel import trajectory_groups
    a = np.array([16e-9, 32e-9, 64e-9])
    Qv = np.arange(-107e-5, 50e-5 + 1e-5, 1e-5)
    ia, f, A, Q = np.meshgrid(np.arange(3), [2e4, 5e5, 4e6], [0., 1e4, 3e5], Qv, indexing='ij')
    ia, f, A, Q = [x.ravel() for x in (ia, f, A, Q)]
    groups = trajectory_groups(ia, f, A, Q)
    # 158 charges of np.arange: 50 of them have a mirror image -> 108 trajectories per (a, f, A)
    assert groups.max() + 1 == 27 * 108
    cost = predicted_log_cost(a[ia], f, A, Q)
    for world in (2, 3, 8):
        owner = np.full(ia.size, -1)
        for r in range(world):
            idx = shard_indices(cost, r, world, groups)
            assert np.all(owner[idx] == -1)
            owner[idx] = r
            assert np.all(np.diff(cost[idx]) <= 0)           # most expensive first
        assert np.all(owner >= 0)
        for g in range(groups.max() + 1):
            assert np.unique(owner[groups == g]).size == 1
        # balanced: the predicted work of the ranks differs by a few percent at most
        work = np.array([np.exp(cost[owner == r]).sum() for r in range(world)])
        assert work.max() / work.min() < 1.1
    # without groups: plain cost-sorted round-robin
    np.testing.assert_array_equal(shard_indices(cost, 1, 2), np.argsort(-cost, kind='stable')[1::2])


# ---------------------------------------------------------------------------------------------
# intermolecular-pressure parameters: host logic of findDeltaEq / LJfitPMavg
# ---------------------------------------------------------------------------------------------
def _ljfit_golden():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ljfit.json')) as fh:
        return json.load(fh)


def test_brentq_matches_scipy():
    from scipy.optimize import brentq as sbrentq
    from pysonic_b200.bls import brentq
    for f, a, b in [(lambda x: x**3 - 2 * x - 5, 2., 3.), (lambda x: np.cos(x) - x, 0., 1.),
                    (lambda x: np.exp(-x) - 1e-3, 0., 20.), (lambda x: (x - 1e-9) * 1e9, -1e-8, 1e-8)]:
        assert brentq(f, a, b, xtol=1e-16) == sbrentq(f, a, b, xtol=1e-16)
    with pytest.raises(ValueError):
        brentq(lambda x: x * x + 1, -1., 1.)


def test_equilibrium_gap_is_bit_identical():
    ''' findDeltaEq (bls.py:493-506) for charges absent from the parameter table. '''
    from pysonic_b200.bls import BilayerSonophore
    for rec in _ljfit_golden()['fits']:
        if rec['Qm0'] == 0.0:
            continue
        b = BilayerSonophore.__new__(BilayerSonophore)
        b.a, b.Qm0, b.S0 = rec['a'], rec['Qm0'], np.pi * rec['a']**2
        D, P = b.findDeltaEq(rec['Qm0'])
        assert D == rec['Delta_eq'] and P == rec['Pnet_eq']


def test_lj_fit_host_logic_with_cpu_quadrature():
    ''' LJfitPMavg with the quadrature replaced by scipy.integrate.quad (the GPU quadrature is tested
        with -m gpu): bracket search, deflection grid and fit reproduce the reference parameters to
        the reference's own reproducibility. '''
    from scipy import integrate
    from pysonic_b200.bls import BilayerSonophore

    def make_pmavg(b):
        def one(Z):
            R, S = b.curvrad(Z), b.surface(Z)

            def g(r):
                z = 0.0 if Z == 0 else np.sign(Z) * (np.sqrt(R**2 - r**2) - abs(R) + abs(Z))
                rel = (2 * z + b.Delta) / b.Delta_
                return 2 * np.pi * r * b.pDelta * ((1 / rel)**b.m - (1 / rel)**b.n)
            return integrate.quad(g, 0, b.a)[0] / S
        return lambda Z: np.array([one(z) for z in np.atleast_1d(Z)])

    for rec in _ljfit_golden()['fits'][:1] + _ljfit_golden()['fits'][5:]:
        b = BilayerSonophore.__new__(BilayerSonophore)
        b.a, b.Qm0, b.S0 = rec['a'], rec['Qm0'], np.pi * rec['a']**2
        b.Delta = rec['Delta']
        LJ, std_err, max_err = b.LJfitPMavg(pmavg=make_pmavg(b))
        tol = max(1e-6, 5 * rec['self_noise'])
        for k in ('x0', 'C', 'nrep', 'nattr'):
            assert abs(LJ[k] - rec[k]) <= tol * abs(rec[k]), (rec['a'], k, LJ[k], rec[k])
        assert std_err < 5e3


# ---------------------------------------------------------------------------------------------
# remaining @addSonicFeatures neurons, passive membrane factory, foreign PointNeuron objects
# ---------------------------------------------------------------------------------------------
def _points_r02():
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'points_r02.json')) as fh:
        return json.load(fh)


def test_new_neuron_descriptors_match_reference():
    g = _points_r02()
    for key, c in g['consts'].items():
        name = key.split('@')[0]
        pn = ps.getPointNeuron(name)
        assert pn.rates == c['rates'] and pn.Cm0 == c['Cm0'] and pn.Qm0 == c['Qm0']
        np.testing.assert_array_equal(pn.Qbounds, c['Qbounds'])
        nbls = ps.NeuronalBilayerSonophore(32e-9, pn)
        assert nbls.Delta == c['Delta'] and nbls.LJ_approx == c['LJ']
    with pytest.raises(ValueError, match='not found'):
        ps.getPointNeuron('LeechR')          # not a SONIC-enabled neuron in the reference either


def test_passive_neuron_factory_mirrors_reference():
    ''' pas.py:16-107: name and lookup name carry the parameters, the resting potential is ELeak,
        there is no rate constant; run_lookups.py:141-145 files its lookup under `lookup_name`. '''
    rec = _points_r02()['passive'][0]
    pas = ps.getDefaultPassiveNeuron()
    assert pas.is_passive and pas.rates == [] and not ps.getPointNeuron('RS').is_passive
    assert pas.name == rec['neuron'] and pas.lookup_name == rec['lookup_name']
    assert pas.Qm0 == pytest.approx(rec['Qm0'], rel=1e-15)
    np.testing.assert_allclose(pas.Qbounds, rec['Qbounds'], rtol=1e-15)
    assert ps.NeuronalBilayerSonophore(32e-9, pas).getLookupFileName() == rec['fname']
    # by name (getPointNeuron on a passive-neuron code) and with other parameters
    again = ps.getPointNeuron(pas.name)
    assert again == pas and again.Cm0 == 1e-2 and again.gLeak == 100.0 and again.Vm0 == -70.0
    p2 = ps.passiveNeuron(2e-2, 50., -65.)
    assert p2.name == 'pas_Cm0_2.0uF_cm2_gLeak_50.0S_m2_ELeak_-65.0mV' and p2.Qm0 == pytest.approx(2e-2 * -65e-3)


def test_foreign_point_neuron_is_verified_not_trusted_by_name():
    ''' An object that merely carries a known name (a reference PointNeuron instance, say) must be
        that neuron: modified parameters raise instead of being ignored. '''
    from pysonic_b200.nbls import as_point_neuron

    class Foreign:
        name, Cm0, Vm0 = 'RS', 1e-2, -71.9
        rates = ['alpham', 'betam', 'alphah', 'betah', 'alphan', 'betan', 'alphap', 'betap']

    assert as_point_neuron(Foreign()) == ps.getPointNeuron('RS')
    f = Foreign()
    f.Vm0 = -65.0
    with pytest.raises(ValueError, match='Vm0'):
        as_point_neuron(f)
    f = Foreign()
    f.Cm0 = 2e-2
    with pytest.raises(ValueError, match='Cm0'):
        as_point_neuron(f)
    f = Foreign()
    f.rates = f.rates[:6]
    with pytest.raises(ValueError, match='rate constants'):
        as_point_neuron(f)


# ---------------------------------------------------------------------------------------------
# table consumption: lookup projections, pulsing protocol, sample plan of the solver
# ---------------------------------------------------------------------------------------------
def test_lookup_projection_is_linear_interpolation():
    from scipy.interpolate import interp1d
    rng = np.random.default_rng(4)
    refs = {'a': np.array([16e-9, 32e-9]), 'A': np.array([0., 1e4, 5e4, 3e5]), 'Q': np.linspace(-1e-3, 5e-4, 7)}
    tabs = {'V': rng.normal(size=(2, 4, 7)), 'alpham': rng.uniform(size=(2, 4, 7))}
    lkp = ps.Lookup(refs, tabs)
    p = lkp.project('A', 2.3e4)
    assert p.inputs == ['a', 'Q'] and p['V'].shape == (2, 7)
    np.testing.assert_allclose(p['V'], interp1d(refs['A'], tabs['V'], axis=1)(2.3e4), rtol=1e-14)
    p2 = lkp.project('A', np.array([0., 2e5]))
    assert p2.inputs == ['a', 'A', 'Q'] and p2['alpham'].shape == (2, 2, 7)
    np.testing.assert_allclose(p2['alpham'], interp1d(refs['A'], tabs['alpham'], axis=1)([0., 2e5]), rtol=1e-14)
    one = lkp.projectN({'a': 32e-9, 'A': 5e4})
    assert one.inputs == ['Q'] and one.ndims == 1
    np.testing.assert_allclose(one['V'], tabs['V'][1, 2], rtol=1e-14)
    np.testing.assert_allclose(one.interpolate1D(-2e-4)['V'], np.interp(-2e-4, refs['Q'], tabs['V'][1, 2]))
    with pytest.raises(ValueError, match='out of'):
        lkp.project('A', 4e5)
    with pytest.raises(ValueError, match='out of'):
        one.interpolate1D(6e-4)
    assert ps.Lookup({'f': np.array([5e5]), 'Q': refs['Q']}, {'V': tabs['V'][:1, 0]}).project('f', 5e5)['V'].shape == (7,)


def test_pulsed_protocol_and_sample_plan_match_reference():
    ''' Transition events (protocols.py:372-391) and the sample times / stimulus states of
        EventDrivenSolver.solve (solvers.py:445-478), against the reference's simulation outputs. '''
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'sonic_sims.json')) as fh:
        cases = json.load(fh)['cases']
    seen = set()
    for c in cases:
        dtf = ps.getPointNeuron(c['neuron']).dt_factor       # neuron-specific output step (chooseTimeStep)
        key = (c['tstim'], c['toffset'], c['PRF'], c['DC'], dtf)
        if key in seen or 'error' in c:
            continue
        seen.add(key)
        pp = ps.PulsedProtocol(c['tstim'], c['toffset'], PRF=c['PRF'], DC=c['DC'])
        assert pp.tstop == c['tstim'] + c['toffset']
        t, x = ps.NeuronalBilayerSonophore._sample_plan(pp.stimEvents(), pp.tstop, ps.DT_EFFECTIVE * dtf)
        assert t.size == c['nsamples'] and t[-1] == c['t_last'] and x.sum() == c['stim_sum']
        np.testing.assert_array_equal(t[:8], c['t_all_head'])
        np.testing.assert_array_equal(t[::c['step']], c['samples']['t'])
        np.testing.assert_array_equal(x[::c['step']], c['samples']['stimstate'])
    assert len(seen) >= 5
    ev = ps.PulsedProtocol(0.1, 0.05, PRF=100., DC=0.5).stimEvents()
    assert len(ev) == 20 and ev[0] == (0.0, 1.0) and ev[1] == (0.005, 0.0)
    with pytest.raises(ValueError):
        ps.PulsedProtocol(0.1, 0.05, PRF=5., DC=0.5)
