# -*- coding: utf-8 -*-
''' N-dimensional lookup container with the reference's on-disk format
    (mirror of PySONIC/core/lookups.py:19-108,381-398 -- storage side only). '''

import os
import pickle



class Lookup:
    ''' Reference vectors + same-shaped N-D tables.

        The pickle written by `toPickle` is `{'refs': dict, 'tables': dict}` with plain dicts
        and numpy arrays, exactly what `PySONIC.core.Lookup.fromPickle` (lookups.py:386-392)
        reads. '''

    def __init__(self, refs, tables):
        self.refs = refs
        self.tables = tables
        for k, v in self.items():
            if v.shape != self.dims:
                raise ValueError(
                    f'{k} Table dimensions {v.shape} does not match references {self.dims}')

    def __repr__(self):
        ref_str = ', '.join([f'{x[0]}: {x[1]}' for x in zip(self.inputs, self.dims)])
        tables_str = ', '.join(self.outputs)
        return f'{self.__class__.__name__}{self.ndims}D({ref_str})[{tables_str}]'

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
        return tuple([x.size for x in self.refs.values()])

    @property
    def ndims(self):
        return len(self.refs)

    @property
    def inputs(self):
        return list(self.refs.keys())

    @property
    def outputs(self):
        return list(self.tables.keys())

    def toPickle(self, fpath):
        with open(fpath, 'wb') as fh:
            pickle.dump({'refs': self.refs, 'tables': self.tables}, fh)

    @classmethod
    def fromPickle(cls, fpath):
        if not os.path.isfile(fpath):
            raise FileNotFoundError(f'Missing lookup file: "{fpath}"')
        with open(fpath, 'rb') as fh:
            d = pickle.load(fh)
        return cls(d['refs'], d['tables'])
