# -*- coding: utf-8 -*-
''' N-dimensional lookup container.

    Storage side: the on-disk format of the reference (PySONIC/core/lookups.py:27-53,381-392): a pickle of
    `{'refs': dict, 'tables': dict}` -- plain dicts of numpy arrays, reference vectors in axis order,
    every table shaped like the reference vectors -- so files written here load in the reference
    (`Lookup.fromPickle`) and vice versa.

    Consumption side (SONIC simulations, nbls.py:389-437): linear projection of every table at a value
    of one input (`project`, lookups.py:234-273, scipy `interp1d(kind='linear')` there), several of them
    (`projectN` :275-290) and interpolation of a 1-D lookup (`interpolate1D` :309-333). '''

import os
import pickle

import numpy as np


def _check_within(key, value, bounds, rel_tol=1e-9):
    ''' Values outside the tabulated range are an error; a value that misses a bound by rounding only
        (relative 1e-9) is moved onto it -- the rule of the reference's utils.isWithin (utils.py:321-350).
        Returns the (possibly corrected) values. '''
    lo, hi = bounds
    v = np.array(value, dtype=float)
    near_lo = (v < lo) & np.isclose(v, lo, rtol=rel_tol, atol=0.0)
    near_hi = (v > hi) & np.isclose(v, hi, rtol=rel_tol, atol=0.0)
    v = np.where(near_lo, lo, np.where(near_hi, hi, v))
    if np.any(v < lo) or np.any(v > hi):
        bad = v if v.ndim == 0 else v[(v < lo) | (v > hi)][0]
        raise ValueError(f'{key} value ({float(bad)}) out of [{lo}, {hi}] interval')
    return v


class Lookup:

    def __init__(self, refs, tables):
        self.refs = refs
        self.tables = tables
        shape = self.dims
        for name, tab in self.tables.items():
            if tab.shape != shape:
                raise ValueError(f'table "{name}" has shape {tab.shape}, reference vectors span {shape}')

    def __repr__(self):
        axes = ', '.join(f'{k}: {n}' for k, n in zip(self.inputs, self.dims))
        return f'{type(self).__name__}{self.ndims}D({axes})[{", ".join(self.outputs)}]'

    # dict-style access to the tables
    def __getitem__(self, key):
        return self.tables[key]

    def __setitem__(self, key, value):
        self.tables[key] = value

    def __delitem__(self, key):
        del self.tables[key]

    def keys(self):
        return self.tables.keys()

    def values(self):
        return self.tables.values()

    def items(self):
        return self.tables.items()

    def refitems(self):
        return self.refs.items()

    @property
    def dims(self):
        return tuple(np.size(v) for v in self.refs.values())

    @property
    def ndims(self):
        return len(self.refs)

    @property
    def inputs(self):
        return list(self.refs)

    @property
    def outputs(self):
        return list(self.tables)

    def copy(self):
        return type(self)(dict(self.refs), dict(self.tables))

    # ---- projections -------------------------------------------------------------------------
    def project(self, key, value):
        ''' New lookup with every table interpolated linearly at `value` (scalar: the dimension
            disappears; array: it is resampled) along input `key`. '''
        if key not in self.refs:
            raise KeyError(f'unknown input dimension: {key}')
        axis = self.inputs.index(key)
        ref = np.asarray(self.refs[key], dtype=float)
        scalar = np.ndim(value) == 0
        val = _check_within(key, np.atleast_1d(np.asarray(value, dtype=float)), (ref.min(), ref.max()))
        new_tables = {}
        if ref.size == 1:
            for k, tab in self.tables.items():
                t = tab.mean(axis=axis, keepdims=True)
                new_tables[k] = np.repeat(t, val.size, axis=axis)
        else:
            # bracketing nodes and weights, shared by all tables
            hi = np.clip(np.searchsorted(ref, val, side='right'), 1, ref.size - 1)
            lo = hi - 1
            w = (val - ref[lo]) / (ref[hi] - ref[lo])
            shape = [1] * self.ndims
            shape[axis] = val.size
            w = w.reshape(shape)
            for k, tab in self.tables.items():
                a, b = np.take(tab, lo, axis=axis), np.take(tab, hi, axis=axis)
                new_tables[k] = a + w * (b - a)
        new_refs = dict(self.refs)
        if scalar:
            del new_refs[key]
            new_tables = {k: np.squeeze(t, axis=axis) for k, t in new_tables.items()}
        else:
            new_refs[key] = val
        return type(self)(new_refs, new_tables)

    def projectN(self, projections):
        lkp = self
        for k, v in projections.items():
            lkp = lkp.project(k, v)
        return lkp

    def interpVar1D(self, value, key):
        if self.ndims != 1:
            raise ValueError('only a 1-dimensional lookup can be interpolated at a point')
        (name, ref), = self.refs.items()
        value = _check_within(name, value, (ref.min(), ref.max()))
        return np.interp(value, ref, self.tables[key], left=np.nan, right=np.nan)

    def interpolate1D(self, value):
        return {k: self.interpVar1D(value, k) for k in self.outputs}

    # ---- storage -------------------------------------------------------------------------------
    def toPickle(self, fpath):
        with open(fpath, 'wb') as fh:
            pickle.dump({'refs': self.refs, 'tables': self.tables}, fh)

    @classmethod
    def fromPickle(cls, fpath):
        if not os.path.isfile(fpath):
            raise FileNotFoundError(f'Missing lookup file: "{fpath}"')
        with open(fpath, 'rb') as fh:
            content = pickle.load(fh)
        return cls(content['refs'], content['tables'])
