# -*- coding: utf-8 -*-
''' Batch dispatcher (mirror of PySONIC/core/batches.py:70-171).

    The reference fans `func(*args)` out over `multiprocess` workers.  Here a queue of
    `NeuronalBilayerSonophore.computeEffVars` calls is recognised and submitted to the GPU as
    ONE batched launch (every queue item becomes a lane of the integrator kernel); any other
    function is simply called item by item, as the reference does with `mpi=False`. '''

import logging
import time

import numpy as np

logger = logging.getLogger('pysonic_b200')


class Batch:

    def __init__(self, func, queue):
        self.func = func
        self.queue = queue

    def __call__(self, *args, **kwargs):
        return self.run(*args, **kwargs)

    @staticmethod
    def resolve(params):
        if isinstance(params, list):
            return params, {}
        if isinstance(params, tuple):
            return params
        raise TypeError('queue items must be lists of args or (args, kwargs) tuples')

    def _effvars_owner(self):
        from .nbls import NeuronalBilayerSonophore
        owner = getattr(self.func, '__self__', None)
        if isinstance(owner, NeuronalBilayerSonophore) and \
                getattr(self.func, '__func__', None) is NeuronalBilayerSonophore.computeEffVars:
            return owner
        return None

    def run(self, mpi=False, loglevel=logging.INFO):
        ''' Run the batch; outputs are returned in queue order. '''
        t0 = time.perf_counter()
        nbls = self._effvars_owner()
        if nbls is not None and len(self.queue) > 0 and all(
                isinstance(q, list) and len(q) == 3 for q in self.queue):
            outputs = self._run_effvars(nbls)
        else:
            outputs = []
            for params in self.queue:
                args, kwargs = self.resolve(params)
                outputs.append(self.func(*args, **kwargs))
        logger.info('Batch of %d job(s) completed in %.3f s', len(self.queue), time.perf_counter() - t0)
        return outputs

    def _run_effvars(self, nbls):
        ''' One GPU launch for the whole queue.  Items are [drive, fs, Qm]; all items must share
            the same fs vector to be batched together (run_lookups.py:100-103 builds them so). '''
        fs0 = np.atleast_1d(np.asarray(self.queue[0][1], float))
        same_fs = all(np.array_equal(np.atleast_1d(np.asarray(q[1], float)), fs0) for q in self.queue)
        if not same_fs:
            return [nbls.computeEffVars(*q) for q in self.queue]
        f = np.array([q[0].f for q in self.queue])
        A = np.array([q[0].A for q in self.queue])
        Q = np.array([float(q[2]) for q in self.queue])
        out, ncyc, status, tpoint, _, _ = nbls.effvars_batch(f, A, Q, fs0)
        keys = ['V'] + nbls.pneuron.rates
        res = []
        for n in range(len(self.queue)):
            effvars = [{k: out[i, n, j] for i, k in enumerate(keys)} for j in range(fs0.size)]
            res.append((effvars, float(tpoint[n])))
        return res

    @staticmethod
    def createQueue(*dims):
        ''' All parameter combinations, first dimension outermost (batches.py:155-171). '''
        grids = np.meshgrid(*dims, indexing='ij')
        return np.stack(grids, -1).reshape(-1, len(dims)).tolist()
