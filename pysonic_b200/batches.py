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
        if nbls is not None and len(self.queue) > 0 and self._is_effvars_queue():
            outputs = self._run_effvars(nbls)
        else:
            outputs = []
            for params in self.queue:
                args, kwargs = self.resolve(params)
                outputs.append(self.func(*args, **kwargs))
        logger.info('Batch of %d job(s) completed in %.3f s', len(self.queue), time.perf_counter() - t0)
        return outputs

    def _items(self):
        ''' Queue items as (drive, fs, Qm, overtones or None): either [drive, fs, Qm] lists or
            ([drive, fs, Qm], {'Qm_overtones': [...]}) tuples (run_lookups.py:100-128), or
            [drive, fs, Qm, overtones] lists (tests/test_Qovertones.py:48-51). '''
        out = []
        for q in self.queue:
            args, kwargs = self.resolve(q)
            ov = kwargs.get('Qm_overtones') if kwargs else None
            if len(args) == 4:
                ov = args[3]
            out.append((args[0], args[1], args[2], ov))
        return out

    def _is_effvars_queue(self):
        try:
            for q in self.queue:
                args, kwargs = self.resolve(q)
                if len(args) not in (3, 4) or not (hasattr(args[0], 'f') and hasattr(args[0], 'A')):
                    return False
                if kwargs and set(kwargs) - {'Qm_overtones'}:
                    return False
        except TypeError:
            return False
        return True

    def _run_effvars(self, nbls):
        ''' One GPU launch for the whole queue.  All items must share the same fs vector and
            number of charge overtones to be batched together (run_lookups.py builds them so). '''
        items = self._items()
        fs0 = np.atleast_1d(np.asarray(items[0][1], float))
        nov0 = 0 if items[0][3] is None else len(items[0][3])
        uniform = all(np.array_equal(np.atleast_1d(np.asarray(q[1], float)), fs0) and
                      (0 if q[3] is None else len(q[3])) == nov0 for q in items)
        if not uniform:
            return [nbls.computeEffVars(q[0], q[1], q[2], q[3]) for q in items]
        from .nbls import check_drive_phase
        for q in items:
            check_drive_phase(q[0])
        f = np.array([q[0].f for q in items])
        A = np.array([q[0].A for q in items])
        Q = np.array([float(q[2]) for q in items])
        ov = np.array([np.asarray(q[3], float).reshape(-1, 2) for q in items]) if nov0 else None
        out, ncyc, status, tpoint, _, _ = nbls.effvars_batch(f, A, Q, fs0, overtones=ov)
        keys = nbls.effvars_keys(nov0)
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
