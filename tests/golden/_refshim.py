# -*- coding: utf-8 -*-
''' Import shim for the UNMODIFIED reference (tjjlemaire/PySONIC) at /root/reference.

    Only used by the golden-vector generators in this directory, and only inside the build
    container: /root/reference does not exist on the GPU box, so nothing in tests/, bench.py
    or the package imports this module at run time.

    The reference imports five non-numerical modules that are absent from the image
    (matplotlib, colorlog, lockfile, boltons, tkinter). They are stubbed; the package
    __init__ (which pulls plotting and argparse front ends) is bypassed. See SURVEY.md §8(c).
'''

import sys
import types
import logging
import importlib.machinery as im

REF_ROOT = '/root/reference'


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__spec__ = im.ModuleSpec(name, None)
    sys.modules[name] = m
    return m


class _Any:
    ''' Object absorbing any attribute access, call or item access. '''

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, k):
        return _Any()

    def __getitem__(self, k):
        return _Any()

    def __setitem__(self, k, v):
        pass

    def __iter__(self):
        return iter([])


class _AnyMod(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return _Any()


def _anymod(name):
    m = _AnyMod(name)
    m.__spec__ = im.ModuleSpec(name, None)
    m.__path__ = []
    sys.modules[name] = m


_loaded = False


def load_reference():
    ''' Make `PySONIC.core`, `PySONIC.neurons`, ... importable from /root/reference. '''
    global _loaded
    if _loaded:
        return
    _mod('lockfile', FileLock=_Any)

    class _Fmt(logging.Formatter):
        def __init__(self, fmt=None, datefmt=None, **kw):
            super().__init__('%(asctime)s %(message)s', datefmt)

    _mod('colorlog', ColoredFormatter=_Fmt, StreamHandler=logging.StreamHandler,
         getLogger=logging.getLogger)
    _mod('boltons').__path__ = []
    _mod('boltons.strutils', cardinalize=lambda s, n: s if n == 1 else s + 's')
    for n in ('tkinter', 'tkinter.filedialog', 'matplotlib', 'matplotlib.pyplot'):
        if n not in sys.modules:
            _anymod(n)

    pkg = types.ModuleType('PySONIC')
    pkg.__path__ = [f'{REF_ROOT}/PySONIC']
    pkg.__spec__ = im.ModuleSpec('PySONIC', None, is_package=True)
    pkg.__spec__.submodule_search_locations = pkg.__path__
    sys.modules['PySONIC'] = pkg

    import PySONIC.core  # noqa: F401  (must precede PySONIC.neurons: circular import)
    import PySONIC.neurons  # noqa: F401
    from PySONIC.utils import logger
    logger.setLevel(logging.ERROR)
    _loaded = True
