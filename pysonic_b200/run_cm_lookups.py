# -*- coding: utf-8 -*-
''' Capacitance lookup generation: `computeCmLookup` and the `run_Cm_lookups.py` command line
    (mirror of scripts/run_Cm_lookups.py:19-113).  Same integrator kernels as the effective-variable
    lookups; the output is the full 1000-sample Cm/Cm0 profile of the converged cycle per (f, A). '''

import argparse
import logging
import os

import numpy as np

from .bls import BilayerSonophore
from .lookups import Lookup
from .run_lookups import _validate

logger = logging.getLogger('pysonic_b200')


def computeCmLookup(bls, fref, Aref, mpi=False, loglevel=logging.INFO):
    ''' Drop-in for `computeCmLookup` of scripts/run_Cm_lookups.py:19: refs `f, A, t`, one table
        `Cm_rel` of shape (nf, nA, 1000). '''
    refs = {'f': fref, 'A': Aref}
    _validate(refs)        # same checks and exception types as for the effective-variable lookups
    refs = {k: np.asarray(v, dtype=np.float64) for k, v in refs.items()}
    dims = [x.size for x in refs.values()]
    logger.log(loglevel, 'Starting Cm simulation batch for %s', bls)
    # queue order of AcousticDrive.createQueue(fref, Aref): f outer, A inner; Qm = 0
    f, A = np.meshgrid(refs['f'], refs['A'], indexing='ij')
    rel_Cm_cycles = bls._zprofiles(f.ravel(), A.ravel(), 0., relcm=True)
    nsamples = rel_Cm_cycles.shape[1]
    refs['t'] = np.linspace(0., 1., nsamples)
    return Lookup(refs, {'Cm_rel': np.ascontiguousarray(rel_Cm_cycles.reshape(dims + [nsamples]))})


def main(argv=None):
    p = argparse.ArgumentParser(description='Create Cm lookup table on the GPU')
    p.add_argument('-a', '--radius', nargs='+', type=float, default=[32.0], help='Sonophore radius (nm)')
    p.add_argument('-f', '--freq', nargs='+', type=float,
                   default=[20., 100., 500., 1e3, 2e3, 3e3, 4e3], help='US frequency (kHz)')
    p.add_argument('-A', '--amp', nargs='+', type=float, default=None, help='Acoustic pressure amplitude (kPa)')
    p.add_argument('--mpi', default=False, action='store_true', help='(accepted for compatibility)')
    p.add_argument('--test', default=False, action='store_true', help='Run test configuration')
    p.add_argument('-o', '--outputdir', type=str, default='.', help='Output directory')
    p.add_argument('-y', '--yes', default=False, action='store_true', help='Overwrite without asking')
    args = p.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s %(message)s')
    freqs = np.array(args.freq) * 1e3
    amps = (np.insert(np.logspace(np.log10(0.1), np.log10(600), num=50), 0, 0.0) if args.amp is None
            else np.array(args.amp)) * 1e3
    bls = BilayerSonophore(args.radius[0] * 1e-9, 1e-2, 0.0)      # run_Cm_lookups.py:79-82
    lookup_fpath = os.path.join(args.outputdir, bls.Cm_lkp_filename)
    inputs = [freqs, amps]
    if args.test:
        inputs = [np.array([x.min(), x.max()]) if x.size > 1 else x for x in inputs]
        fcode, fext = os.path.splitext(lookup_fpath)
        lookup_fpath = f'{fcode}_test{fext}'
    if os.path.isfile(lookup_fpath) and not args.yes:
        logger.warning(f'"{lookup_fpath}" file already exists and will be overwritten. Continue? (y/n)')
        if input() not in ['y', 'Y']:
            logger.error('Cm-lookup creation canceled')
            return
    lkp = computeCmLookup(bls, *inputs, mpi=args.mpi)
    logger.info(f'Generated Cm-lookup: {lkp}')
    lkp.toPickle(lookup_fpath)


if __name__ == '__main__':
    main()
