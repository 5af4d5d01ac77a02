# -*- coding: utf-8 -*-
''' Lookup-table generation front end: `computeAStimLookup` and the `run_lookups.py` command
    line, with the reference's signatures, grid order, table layout and file naming
    (scripts/run_lookups.py:22-238, PySONIC/parsers.py:422-529). '''

import argparse
import logging
import os
import time

import numpy as np

from . import _lib
from .constants import DQ_LOOKUP
from .lookups import Lookup
from .nbls import NeuronalBilayerSonophore, as_point_neuron, check_charges
from .neurons import getDefaultPassiveNeuron, getPointNeuron
from .parallel import (dist_info, env_world, finalize, init_from_env, predicted_log_cost, run_sharded,
                       trajectory_groups)

logger = logging.getLogger('pysonic_b200')

_DESCS = {
    'a': 'sonophore radii',
    'f': 'US frequencies',
    'A': 'US amplitudes',
    'fs': 'sonophore membrane coverage fractions',
    'Q': 'membrane charge densities',
    'overtones': 'charge Fourier overtones',
}


def _is_iterable(x):
    return isinstance(x, (list, tuple, np.ndarray))


def _validate(refs):
    ''' Same checks, messages and exception types as run_lookups.py:85-96. '''
    for key, values in refs.items():
        desc = _DESCS[key]
        if not _is_iterable(values):
            raise TypeError(f'Invalid {desc} (must be provided as list or numpy array)')
        if not all(isinstance(x, float) for x in values):
            raise TypeError(f'Invalid {desc} (must all be float typed)')
        if len(values) == 0:
            raise ValueError(f'Empty {key} array')
        if key in ('a', 'f') and min(values) <= 0:
            raise ValueError(f'Invalid {desc} (must all be strictly positive)')
        if key in ('A', 'fs') and min(values) < 0:
            raise ValueError(f'Invalid {desc} (must all be positive or null)')


def build_refs(aref, fref, Aref, fsref, Qref, novertones=0, test=False):
    ''' Reference vectors of a lookup, shaped and validated as run_lookups.py:48-96 does (without the
        overtone axes, which `_overtones_lookup` appends as run_lookups.py:105-128 does). '''
    refs = {'a': aref, 'f': fref, 'A': Aref, 'Q': Qref}
    fsref = np.asarray(fsref) if _is_iterable(fsref) else fsref
    err_span = 'cannot span {} for more than 1 {}'
    if _is_iterable(fsref) and len(fsref) > 0 and (len(fsref) > 1 or fsref[0] != 1.):
        for x in ['a', 'f']:
            assert len(refs[x]) == 1, err_span.format(_DESCS['fs'], _DESCS[x])
    refs['fs'] = fsref
    if novertones > 0:
        # single radius, frequency and coverage fraction (run_lookups.py:64-67) ...
        for x in ['a', 'f', 'fs']:
            assert np.size(refs[x]) == 1, err_span.format(_DESCS['overtones'], _DESCS[x])
        # ... and down-sampled charge and amplitude vectors (run_lookups.py:69-79)
        nQmax, nAmax = 50, 15
        if len(refs['Q']) > nQmax:
            refs['Q'] = np.linspace(refs['Q'][0], refs['Q'][-1], nQmax)
        if len(refs['A']) > nAmax:
            refs['A'] = np.insert(
                np.logspace(np.log10(refs['A'][1]), np.log10(refs['A'][-1]), num=nAmax - 1), 0, 0.0)
    if novertones > _lib_max_overtones():
        raise ValueError(f'at most {_lib_max_overtones()} charge overtones are supported')
    if test:
        refs = {k: np.array([v.min(), v.max()]) if v.size > 1 else v for k, v in refs.items()}
    _validate(refs)
    return {k: np.asarray(v, dtype=np.float64) for k, v in refs.items()}


def overtone_refs(refs, novertones, test=False):
    ''' Append the overtone axes AQ1, phiQ1, ... before `fs` (run_lookups.py:105-128). '''
    nAQ, nphiQ = 5, 5
    AQ_ref = np.linspace(0, 100e-5, nAQ)                          # C/m2
    phiQ_ref = np.linspace(0, 2 * np.pi, nphiQ, endpoint=False)   # rad
    if test:
        AQ_ref = np.array([AQ_ref.min(), AQ_ref.max()])
        phiQ_ref = np.array([phiQ_ref.min(), phiQ_ref.max()])
    refs = dict(refs)
    fsref = refs.pop('fs')
    for i in range(novertones):
        refs[f'AQ{i + 1}'] = AQ_ref
        refs[f'phiQ{i + 1}'] = phiQ_ref
    refs['fs'] = fsref
    return refs


def computeAStimLookup(pneuron, aref, fref, Aref, fsref, Qref, novertones=0,
                       test=False, mpi=False, loglevel=logging.INFO, return_info=False,
                       device=None, shard=True, check_charge=False):
    ''' Effective-variable lookup tables over (a, f, A, Q, fs).

        Drop-in for `computeAStimLookup` of scripts/run_lookups.py:22: same arguments (SI units:
        m, Hz, Pa, -, C/m2), same validation, same table layout `(na, nf, nA, nQ, nfs)`, same
        key order (`V`, rates in `pneuron.rates` order, `tcomp`).  `mpi=True` means "use every
        available GPU": under `torchrun` the grid is sharded over the ranks (one process per
        GPU), otherwise over the visible devices of this process.

        Extra keyword arguments (not in the reference): `return_info` also returns the per-point
        cycle counts / status words / run statistics; `device` pins the run to one CUDA device;
        `shard=False` makes a rank under `torchrun` compute the whole grid on its own device;
        `check_charge=True` applies the physiological-range check of `checkInputs` (bls.py:674-677) to the
        imposed charges; the reference runs that check for `simulate` only (model.py:169), not on this path.

        :return: Lookup (and, if return_info, a dict with ncycles/status/stats)
    '''
    pneuron = as_point_neuron(pneuron)
    refs = build_refs(aref, fref, Aref, fsref, Qref, novertones=novertones, test=test)
    if novertones > 0:
        return _overtones_lookup(pneuron, refs, novertones, test, loglevel, return_info, device, check_charge)
    if check_charge:
        check_charges(refs['Q'])
    dims = tuple(x.size for x in refs.values())
    na, nf, nA, nQ, nfs = dims
    keys = ['V'] + pneuron.rates
    nrates = len(pneuron.rates)
    logger.log(loglevel, 'Starting lookup batch for %s neuron: %d points x %d fs', pneuron.name,
               na * nf * nA * nQ, nfs)

    # a passive neuron is simulated on the sonophore of the default passive neuron (run_lookups.py:141-145)
    bls_neuron = getDefaultPassiveNeuron() if pneuron.is_passive else pneuron
    bls_params = [NeuronalBilayerSonophore(float(a), bls_neuron).abi_params() for a in refs['a']]
    if mpi:
        init_from_env()          # under torchrun: bind to the rank's GPU, join the process group
    rank, world, local_rank = dist_info()
    t0 = time.perf_counter()
    if world > 1 and shard:
        # one process per GPU: shard the flattened (a > f > A > Q) list over the ranks
        ia, fi, Ai, Qi = np.meshgrid(np.arange(na), refs['f'], refs['A'], refs['Q'], indexing='ij')
        ia, fi, Ai, Qi = [x.ravel() for x in (ia, fi, Ai, Qi)]
        n = ia.size
        cost = predicted_log_cost(refs['a'][ia], fi, Ai, Qi)

        def compute(idx):
            out, ncyc, status, tpoint, nrhs, st = _lib.points_run(
                local_rank, bls_params, pneuron.neuron_id, nrates, ia[idx].astype(np.int32),
                fi[idx], Ai[idx], Qi[idx], refs['fs'])
            return [(out, 1), (ncyc, 0), (status, 0), (tpoint, 0), (nrhs, 0)]

        out, ncyc, status, tpoint, nrhs = run_sharded(compute, n, cost, groups=trajectory_groups(ia, fi, Ai, Qi))
        out = out.reshape((1 + nrates,) + dims)
        ncyc, status, tpoint = [x.reshape(dims[:-1]) for x in (ncyc, status, tpoint)]
        stats = {'n_points': n, 'n_rhs': int(nrhs.sum()), 'world_size': world}
    else:
        ndev = _lib.device_count()
        if device is not None:
            mask = 1 << int(device)
        elif world > 1:
            mask = 1 << local_rank
        else:
            mask = (1 << ndev) - 1 if (mpi and ndev > 1) else 1
        out, ncyc, status, tpoint, stats = _lib.lookup_run(
            bls_params, pneuron.neuron_id, nrates, refs['f'], refs['A'], refs['Q'], refs['fs'], mask)
    wall = time.perf_counter() - t0
    logger.log(loglevel, 'Lookup batch completed in %.3f s', wall)

    tables = {k: np.ascontiguousarray(out[i]) for i, k in enumerate(keys)}
    # per-point computation time, tiled over the fs dimension (run_lookups.py:169-172)
    tables['tcomp'] = np.ascontiguousarray(
        np.moveaxis(np.array([tpoint for _ in range(nfs)]), 0, -1))
    lkp = Lookup(refs, tables)
    if return_info:
        return lkp, {'ncycles': ncyc, 'status': status, 'stats': stats, 'wall_s': wall}
    return lkp


def computeAStimLookups(pneurons, aref, fref, Aref, fsref, Qrefs, test=False, mpi=False,
                        loglevel=logging.INFO, return_info=False, device=None, check_charge=False):
    ''' Lookups of SEVERAL neurons over the same (a, f, A, fs) vectors in one batch: what the
        `for name in args['neuron']` loop of scripts/run_lookups.py:193-238 computes with one
        computeAStimLookup call per neuron.  All grids go through a single integrator launch
        (trajectories depend on the sonophore constants and |Q| only, so neurons with the same resting
        charge share them, and the long serial chains of every grid overlap), followed by one averaging
        launch per neuron.  Every table is bit-identical to the single-neuron call.

        :param pneurons: list of point-neuron models (or names)
        :param Qrefs: one charge vector per neuron (C/m2)
        :return: list of Lookup objects (and, if return_info, a dict with per-neuron ncycles/status + stats)
    '''
    pneurons = [as_point_neuron(pn) for pn in pneurons]
    all_refs = [build_refs(aref, fref, Aref, fsref, Q, novertones=0, test=test) for Q in Qrefs]
    if check_charge:
        for refs in all_refs:
            check_charges(refs['Q'])
    r0 = all_refs[0]
    bls = [[NeuronalBilayerSonophore(float(a), pn).abi_params() for a in r0['a']] for pn in pneurons]
    ndev = _lib.device_count()
    if device is not None:
        mask = 1 << int(device)
    else:
        mask = (1 << ndev) - 1 if (mpi and ndev > 1) else 1
    logger.log(loglevel, 'Starting lookup batch for %s neurons: %d points x %d fs', [pn.name for pn in pneurons],
               sum(int(np.prod([x.size for x in list(r.values())[:-1]])) for r in all_refs), r0['fs'].size)
    t0 = time.perf_counter()
    res, stats = _lib.lookup_run_multi(bls, [pn.neuron_id for pn in pneurons], [len(pn.rates) for pn in pneurons],
                                       r0['f'], r0['A'], [r['Q'] for r in all_refs], r0['fs'], mask)
    wall = time.perf_counter() - t0
    logger.log(loglevel, 'Lookup batch completed in %.3f s', wall)
    lkps, infos = [], []
    nfs = r0['fs'].size
    for pn, refs, (out, ncyc, status, tpoint) in zip(pneurons, all_refs, res):
        tables = {k: np.ascontiguousarray(out[i]) for i, k in enumerate(['V'] + pn.rates)}
        tables['tcomp'] = np.ascontiguousarray(np.moveaxis(np.array([tpoint for _ in range(nfs)]), 0, -1))
        lkps.append(Lookup(refs, tables))
        infos.append({'ncycles': ncyc, 'status': status})
    if return_info:
        return lkps, {'neurons': infos, 'stats': stats, 'wall_s': wall}
    return lkps


def _lib_max_overtones():
    return 4     # SONIC_MAX_OVERTONES of the native library


def _overtones_lookup(pneuron, refs, novertones, test, loglevel, return_info, device, check_charge=False):
    ''' Lookup with charge overtones (run_lookups.py:105-128): every (a, f, A, Q) point is
        combined with every (AQ1, phiQ1, ..., AQn, phiQn) combination; the overtone dimensions come
        after Q and before fs, and every overtone adds the tables A_Vk, phi_Vk after V. '''
    refs = overtone_refs(refs, novertones, test)
    dims = tuple(x.size for x in refs.values())
    na = dims[0]
    grids = np.meshgrid(np.arange(na), *[refs[k] for k in list(refs)[1:-1]], indexing='ij')
    ia = grids[0].ravel().astype(np.int32)
    f, A, Q = (g.ravel() for g in grids[1:4])
    ov = np.stack([g.ravel() for g in grids[4:]], axis=1).reshape(ia.size, novertones, 2)
    nrates = len(pneuron.rates)
    if check_charge:
        check_charges(Q, ov)
    bls_params = [NeuronalBilayerSonophore(float(a), pneuron).abi_params() for a in refs['a']]
    logger.log(loglevel, 'Starting lookup batch for %s neuron: %d points (%d charge overtones) x %d fs',
               pneuron.name, ia.size, novertones, dims[-1])
    t0 = time.perf_counter()
    dev = dist_info()[2] if device is None else int(device)
    out, ncyc, status, tpoint, nrhs, stats = _lib.points_run(
        dev, bls_params, pneuron.neuron_id, nrates, ia, f, A, Q, refs['fs'], overtones=ov)
    wall = time.perf_counter() - t0
    logger.log(loglevel, 'Lookup batch completed in %.3f s', wall)
    keys = ['V']
    for i in range(1, novertones + 1):
        keys += [f'A_V{i}', f'phi_V{i}']
    keys += pneuron.rates
    tables = {k: np.ascontiguousarray(out[i].reshape(dims)) for i, k in enumerate(keys)}
    tp = tpoint.reshape(dims[:-1])
    tables['tcomp'] = np.ascontiguousarray(np.moveaxis(np.array([tp for _ in range(dims[-1])]), 0, -1))
    lkp = Lookup(refs, tables)
    if return_info:
        return lkp, {'ncycles': ncyc.reshape(dims[:-1]), 'status': status.reshape(dims[:-1]), 'stats': stats,
                     'wall_s': wall}
    return lkp


def _parser():
    ''' Same flags, units and defaults as MechSimParser + run_lookups.main
        (parsers.py:422-529, run_lookups.py:180-189). '''
    p = argparse.ArgumentParser(description='Create SONIC lookup tables on the GPU')
    p.add_argument('-n', '--neuron', type=str, nargs='+', default=['RS'], help='Neuron name (string)')
    p.add_argument('-a', '--radius', nargs='+', type=float, default=[16.0, 32.0, 64.0],
                   help='Sonophore radius (nm)')
    p.add_argument('-f', '--freq', nargs='+', type=float,
                   default=[20., 100., 500., 1e3, 2e3, 3e3, 4e3], help='US frequency (kHz)')
    p.add_argument('-A', '--amp', nargs='+', type=float, default=None,
                   help='Acoustic pressure amplitude (kPa)')
    p.add_argument('-Q', '--charge', nargs='+', type=float, default=None,
                   help='Membrane charge density (nC/cm2)')
    p.add_argument('--fs', nargs='+', type=float, default=[100.], help='Sonophore coverage fraction (%%)')
    p.add_argument('--spanFs', default=False, action='store_true', help='Span Fs from 1 to 100%%')
    p.add_argument('--mpi', default=False, action='store_true', help='Use all available GPUs')
    p.add_argument('--test', default=False, action='store_true', help='Run test configuration')
    p.add_argument('--novertones', type=int, default=0, help='Number of Fourier overtones')
    p.add_argument('-v', '--verbose', default=False, action='store_true', help='Increase verbosity')
    p.add_argument('-o', '--outputdir', type=str, default=None, help='Output directory')
    p.add_argument('-y', '--yes', default=False, action='store_true',
                   help='Overwrite existing lookup files without asking')
    return p


def main(argv=None):
    args = _parser().parse_args(argv)
    loglevel = logging.DEBUG if args.verbose else logging.INFO
    logging.basicConfig(level=loglevel, format='%(asctime)s %(message)s')
    # one process per GPU under torchrun: every rank computes its shard, rank 0 alone talks to the
    # user and writes the files
    rank, world, _ = init_from_env()
    try:
        _run(args, loglevel, rank)
    finally:
        finalize()


def _rank0_says(flag):
    ''' Rank 0's decision, known to every rank of the process group. '''
    if dist_info()[1] == 1:
        return flag
    import torch.distributed as dist
    box = [bool(flag)]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def _run(args, loglevel, rank):
    radii = np.array(args.radius) * 1e-9
    freqs = np.array(args.freq) * 1e3
    if args.amp is None:
        amps = np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) * 1e3
    else:
        amps = np.array(args.amp) * 1e3
    fs = np.arange(1, 101) * 1e-2 if args.spanFs else np.array(args.fs) * 1e-2
    jobs = []
    for name in args.neuron:
        pneuron = getPointNeuron(name)
        if args.charge is None:
            Qmin, Qmax = pneuron.Qbounds
            charges = np.arange(Qmin, Qmax + DQ_LOOKUP, DQ_LOOKUP)
        else:
            charges = np.array(args.charge) * 1e-5
        input_args = {'a': radii, 'f': freqs, 'A': amps, 'fs': fs}
        fname_args = {k: v[0] if v.size == 1 else None for k, v in input_args.items()}
        fname_args['novertones'] = args.novertones
        nbls = NeuronalBilayerSonophore(32e-9, pneuron)
        if args.outputdir is not None:
            lookup_fpath = os.path.join(args.outputdir, nbls.getLookupFileName(**fname_args))
        else:
            lookup_fpath = nbls.getLookupFilePath(**fname_args)
        if args.test:
            fcode, fext = os.path.splitext(lookup_fpath)
            lookup_fpath = f'{fcode}_test{fext}'
        go = True
        if rank == 0 and os.path.isfile(lookup_fpath) and not args.yes:
            logger.warning(f'"{lookup_fpath}" file already exists and will be overwritten. '
                           'Continue? (y/n)')
            go = input() in ['y', 'Y']
        if not _rank0_says(go):
            if rank == 0:
                logger.error('%s Lookup creation canceled', pneuron.name)
            return
        jobs.append((pneuron, charges, lookup_fpath))

    def save(pneuron, lkp, lookup_fpath):
        logger.info(f'Generated lookup: {lkp}')
        if rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(lookup_fpath)), exist_ok=True)
            logger.info('Saving %s neuron lookup in file: "%s"', pneuron.name, lookup_fpath)
            lkp.toPickle(lookup_fpath)

    if len(jobs) > 1 and args.novertones == 0 and dist_info()[1] == 1:
        # several neurons, one process: one integrator launch for all their grids
        lkps = computeAStimLookups([j[0] for j in jobs], radii, freqs, amps, fs, [j[1] for j in jobs],
                                   test=args.test, mpi=args.mpi, loglevel=loglevel)
        for (pneuron, _, lookup_fpath), lkp in zip(jobs, lkps):
            save(pneuron, lkp, lookup_fpath)
        return
    for pneuron, charges, lookup_fpath in jobs:
        lkp = computeAStimLookup(pneuron, radii, freqs, amps, fs, charges, novertones=args.novertones,
                                 test=args.test, mpi=args.mpi, loglevel=loglevel)
        save(pneuron, lkp, lookup_fpath)


if __name__ == '__main__':
    main()
