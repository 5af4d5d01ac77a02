# -*- coding: utf-8 -*-
''' ctypes binding of libsonic_b200.so (C ABI declared in include/sonic_b200.h).

    There is no CPU fallback: if the shared library is missing, or no CUDA device is present
    when a compute entry point is called, an exception is raised.
'''

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# (development: PYSONIC_B200_LIB selects another build of the same library, e.g. a kernel variant to time)
LIB_PATH = os.environ.get('PYSONIC_B200_LIB') or os.path.join(_HERE, 'libsonic_b200.so')


class SonicError(RuntimeError):
    ''' Error reported by the native library. '''


class SonicBlsParams(C.Structure):
    _fields_ = [(k, C.c_double) for k in ('a', 'Delta', 'x0', 'C', 'nrep', 'nattr', 'Cm0', 'depth')]


class SonicStats(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ('n_points', 'n_rhs', 'n_jac', 'n_steps', 'n_cycles',
                                          'n_launches')] + \
               [(k, C.c_double) for k in ('ms_z0', 'ms_integrate', 'ms_average', 'ms_total')]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_uint32)
_bp = C.POINTER(SonicBlsParams)
_sp = C.POINTER(SonicStats)

EXPORTS = {
    # name: (restype, argtypes)
    'sonic_version': (C.c_int, []),
    'sonic_device_count': (C.c_int, []),
    'sonic_last_error': (C.c_int, [C.c_char_p, C.c_int]),
    'sonic_neuron_count': (C.c_int, []),
    'sonic_neuron_id': (C.c_int, [C.c_char_p]),
    'sonic_neuron_name': (C.c_int, [C.c_int, C.c_char_p, C.c_int]),
    'sonic_neuron_nrates': (C.c_int, [C.c_int]),
    'sonic_neuron_rate_name': (C.c_int, [C.c_int, C.c_int, C.c_char_p, C.c_int]),
    'sonic_eval_rates': (C.c_int, [C.c_int, C.c_int, _dp, C.c_int64, _dp]),
    'sonic_mean_rates': (C.c_int, [C.c_int, C.c_int, _dp, C.c_int64, _dp]),
    'sonic_points_run': (C.c_int, [C.c_int, _bp, C.c_int, C.c_int, C.c_int64, _ip, _dp, _dp, _dp, _dp,
                                   C.c_int, _dp, _ip, _up, _dp, _up, _sp]),
    'sonic_points_run_ex': (C.c_int, [C.c_int, _bp, C.c_int, C.c_int, C.c_int64, _ip, _dp, _dp, _dp, C.c_int, _dp,
                                      _dp, C.c_int, _dp, _ip, _up, _dp, _up, _sp]),
    'sonic_lookup_run': (C.c_int, [_bp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int,
                                   C.c_int, C.c_uint32, _dp, _ip, _up, _dp, _sp]),
    'sonic_plan_create': (C.c_int, [C.c_int, _bp, C.c_int, C.c_int, C.c_int64, _ip, _dp, _dp, _dp, _dp,
                                    C.c_int, C.POINTER(C.c_void_p)]),
    'sonic_plan_create_ex': (C.c_int, [C.c_int, _bp, C.c_int, C.c_int, C.c_int64, _ip, _dp, _dp, _dp, C.c_int, _dp,
                                       _dp, C.c_int, C.POINTER(C.c_void_p)]),
    'sonic_plan_create_multi': (C.c_int, [C.c_int, _bp, _ip, C.c_int, _ip, C.c_int, C.c_int64, _ip, _dp, _dp, _dp,
                                          _dp, C.c_int, C.POINTER(C.c_void_p)]),
    'sonic_points_run_multi': (C.c_int, [C.c_int, _bp, _ip, C.c_int, _ip, C.c_int, C.c_int64, _ip, _dp, _dp, _dp,
                                         _dp, C.c_int, _dp, _ip, _up, _dp, _up, _sp]),
    'sonic_lookup_run_multi': (C.c_int, [_bp, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, _ip, _dp, C.c_int, _ip,
                                         C.c_int, C.c_uint32, _dp, _ip, _up, _dp, _sp]),
    'sonic_plan_set_stream': (C.c_int, [C.c_void_p, C.c_void_p]),
    'sonic_plan_launch': (C.c_int, [C.c_void_p]),
    'sonic_plan_sync': (C.c_int, [C.c_void_p]),
    'sonic_plan_fetch': (C.c_int, [C.c_void_p, _dp, _ip, _up, _dp, _up]),
    'sonic_plan_fetch_zprofiles': (C.c_int, [C.c_void_p, _dp]),
    'sonic_plan_fetch_relcm': (C.c_int, [C.c_void_p, _dp]),
    'sonic_plan_stats': (C.c_int, [C.c_void_p, _sp]),
    'sonic_plan_destroy': (C.c_int, [C.c_void_p]),
    'sonic_pmavg': (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int64, _dp, _dp, _ip]),
    'sonic_simulate': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp,
                                 C.POINTER(C.c_uint8), _dp, C.c_int, _dp, _ip]),
    'sonic_sim_nstates': (C.c_int, [C.c_int]),
    'sonic_trim': (C.c_int, []),
    'sonic_fp64_peak': (C.c_int, [C.c_int, _dp]),
}

_lib = None


def load():
    ''' Load the native library (once) and declare its prototypes. '''
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SonicError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; '
            f'g.build()"` (nvcc, sm_100a). pysonic_b200 has no CPU fallback.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc < 0:
        buf = C.create_string_buffer(1024)
        load().sonic_last_error(buf, 1024)
        raise SonicError(f'libsonic_b200 error {rc}: {buf.value.decode(errors="replace")}')
    return rc


def _d(x):
    return x.ctypes.data_as(_dp)


def as_f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def bls_array(params):
    ''' list of dicts / SonicBlsParams -> ctypes array '''
    arr = (SonicBlsParams * len(params))()
    for i, p in enumerate(params):
        for k, _ in SonicBlsParams._fields_:
            setattr(arr[i], k, float(p[k]))
    return arr


def device_count():
    return load().sonic_device_count()


def neuron_rate_names(neuron_id):
    lib = load()
    n = check(lib.sonic_neuron_nrates(neuron_id))
    out = []
    buf = C.create_string_buffer(64)
    for i in range(n):
        check(lib.sonic_neuron_rate_name(neuron_id, i, buf, 64))
        out.append(buf.value.decode())
    return out


def eval_rates(pneuron, Vm, device=0):
    ''' dict rate -> array over Vm, evaluated by the generated device functions. '''
    lib = load()
    Vm = as_f64(Vm).ravel()
    nr = len(pneuron.rates)
    out = np.empty((nr, Vm.size))
    check(lib.sonic_eval_rates(device, pneuron.neuron_id, _d(Vm), Vm.size, _d(out)))
    return {k: out[i] for i, k in enumerate(pneuron.rates)}


def eval_mean_rates(pneuron, Vm, device=0):
    lib = load()
    Vm = as_f64(Vm).ravel()
    nr = len(pneuron.rates)
    out = np.empty(nr)
    check(lib.sonic_mean_rates(device, pneuron.neuron_id, _d(Vm), Vm.size, _d(out)))
    return {k: float(out[i]) for i, k in enumerate(pneuron.rates)}


def trim():
    ''' Release the device workspace kept between one-shot calls. '''
    check(load().sonic_trim())


def fp64_peak(device=0):
    v = C.c_double()
    check(load().sonic_fp64_peak(device, C.byref(v)))
    return v.value


class Plan:
    ''' Device-resident batch of (radius, f, A, Q) points (split form of sonic_points_run). '''

    def __init__(self, device, bls_params, neuron_id, nrates, ia, f, A, Q, fs, overtones=None):
        lib = load()
        self.ia = np.ascontiguousarray(ia, dtype=np.int32)
        self.f, self.A, self.Q, self.fs = as_f64(f), as_f64(A), as_f64(Q), as_f64(fs)
        self.ov, nov = as_overtones(overtones, self.f.size)
        self.n, self.nfs, self.nvar = self.f.size, self.fs.size, 1 + 2 * nov + nrates
        self._bls = bls_array(bls_params)
        self._h = C.c_void_p()
        check(lib.sonic_plan_create_ex(device, self._bls, len(bls_params), neuron_id, self.n,
                                       self.ia.ctypes.data_as(_ip), _d(self.f), _d(self.A), _d(self.Q),
                                       nov, _d(self.ov) if nov else None, _d(self.fs), self.nfs,
                                       C.byref(self._h)))

    def set_stream(self, cuda_stream):
        ''' Launch on a caller-owned stream (integer cudaStream_t handle). '''
        check(load().sonic_plan_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def launch(self):
        check(load().sonic_plan_launch(self._h))

    def sync(self):
        check(load().sonic_plan_sync(self._h))

    def fetch(self):
        out = np.empty((self.nvar, self.n, self.nfs))
        ncyc = np.empty(self.n, dtype=np.int32)
        status = np.empty(self.n, dtype=np.uint32)
        tpoint = np.empty(self.n)
        nrhs = np.empty(self.n, dtype=np.uint32)
        check(load().sonic_plan_fetch(self._h, _d(out), ncyc.ctypes.data_as(_ip),
                                      status.ctypes.data_as(_up), _d(tpoint), nrhs.ctypes.data_as(_up)))
        return out, ncyc, status, tpoint, nrhs

    def fetch_zprofiles(self):
        z = np.empty((self.n, 1000))
        check(load().sonic_plan_fetch_zprofiles(self._h, _d(z)))
        return z

    def fetch_relcm(self):
        ''' Cm(Z(t)) / Cm0 over the last cycle, [n, 1000]. '''
        cm = np.empty((self.n, 1000))
        check(load().sonic_plan_fetch_relcm(self._h, _d(cm)))
        return cm

    def stats(self):
        st = SonicStats()
        check(load().sonic_plan_stats(self._h, C.byref(st)))
        return st.asdict()

    def destroy(self):
        if self._h:
            load().sonic_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def as_overtones(overtones, n):
    ''' None or array-like [n, nov, 2] of (amplitude C/m2, phase rad) -> (contiguous array, nov) '''
    if overtones is None:
        return None, 0
    ov = np.ascontiguousarray(np.asarray(overtones, dtype=np.float64))
    if ov.ndim != 3 or ov.shape[0] != n or ov.shape[2] != 2:
        raise ValueError(f'charge overtones must have shape (n, novertones, 2), got {ov.shape}')
    return ov, ov.shape[1]


def points_run(device, bls_params, neuron_id, nrates, ia, f, A, Q, fs, overtones=None):
    ''' One-shot call of sonic_points_run(_ex) with host buffers.
        :return: (tables[1+2*nov+nrates, n, nfs], ncycles[n], status[n], tpoint[n], nrhs[n], stats) '''
    lib = load()
    ia = np.ascontiguousarray(ia, dtype=np.int32)
    f, A, Q, fs = as_f64(f), as_f64(A), as_f64(Q), as_f64(fs)
    n, nfs = f.size, fs.size
    ov, nov = as_overtones(overtones, n)
    out = np.empty((1 + 2 * nov + nrates, n, nfs))
    ncyc = np.empty(n, dtype=np.int32)
    status = np.empty(n, dtype=np.uint32)
    tpoint = np.empty(n)
    nrhs = np.empty(n, dtype=np.uint32)
    st = SonicStats()
    arr = bls_array(bls_params)
    check(lib.sonic_points_run_ex(device, arr, len(bls_params), neuron_id, n, ia.ctypes.data_as(_ip),
                                  _d(f), _d(A), _d(Q), nov, _d(ov) if nov else None, _d(fs), nfs, _d(out),
                                  ncyc.ctypes.data_as(_ip), status.ctypes.data_as(_up), _d(tpoint),
                                  nrhs.ctypes.data_as(_up), C.byref(st)))
    return out, ncyc, status, tpoint, nrhs, st.asdict()


def lookup_run(bls_params, neuron_id, nrates, f, A, Q, fs, device_mask=1):
    ''' Full-grid call of sonic_lookup_run with host buffers.
        :return: (tables[1+nrates, na, nf, nA, nQ, nfs], ncycles, status, tpoint, stats) '''
    lib = load()
    f, A, Q, fs = as_f64(f), as_f64(A), as_f64(Q), as_f64(fs)
    na = len(bls_params)
    dims = (na, f.size, A.size, Q.size)
    out = np.empty((1 + nrates,) + dims + (fs.size,))
    ncyc = np.empty(dims, dtype=np.int32)
    status = np.empty(dims, dtype=np.uint32)
    tpoint = np.empty(dims)
    st = SonicStats()
    arr = bls_array(bls_params)
    check(lib.sonic_lookup_run(arr, na, _d(f), f.size, _d(A), A.size, _d(Q), Q.size, _d(fs), fs.size,
                               neuron_id, device_mask, _d(out), ncyc.ctypes.data_as(_ip),
                               status.ctypes.data_as(_up), _d(tpoint), C.byref(st)))
    return out, ncyc, status, tpoint, st.asdict()


def lookup_run_multi(bls_params_per_neuron, neuron_ids, nrates, f, A, Qs, fs, device_mask=1):
    ''' Grids of several neurons in one batch (sonic_lookup_run_multi).
        :param bls_params_per_neuron: one list of `na` sonophore dicts per neuron
        :param Qs: one charge vector per neuron
        :return: list of (tables[1+nrates_k, na, nf, nA, nQ_k, nfs], ncycles, status, tpoint) per neuron, stats '''
    lib = load()
    f, A, fs = as_f64(f), as_f64(A), as_f64(fs)
    Qs = [as_f64(q) for q in Qs]
    nn = len(neuron_ids)
    na = len(bls_params_per_neuron[0])
    assert all(len(b) == na for b in bls_params_per_neuron) and len(Qs) == nn and len(nrates) == nn
    arr = bls_array([p for b in bls_params_per_neuron for p in b])
    Qcat = np.concatenate(Qs)
    Qoff = np.concatenate([[0], np.cumsum([q.size for q in Qs])]).astype(np.int32)
    ids = np.ascontiguousarray(neuron_ids, dtype=np.int32)
    dims = [(na, f.size, A.size, q.size) for q in Qs]
    npts = [int(np.prod(d)) for d in dims]
    out = np.empty(sum((1 + nr) * n * fs.size for nr, n in zip(nrates, npts)))
    ncyc = np.empty(sum(npts), dtype=np.int32)
    status = np.empty(sum(npts), dtype=np.uint32)
    tpoint = np.empty(sum(npts))
    st = SonicStats()
    check(lib.sonic_lookup_run_multi(arr, na, _d(f), f.size, _d(A), A.size, _d(Qcat), Qoff.ctypes.data_as(_ip),
                                     _d(fs), fs.size, ids.ctypes.data_as(_ip), nn, device_mask, _d(out),
                                     ncyc.ctypes.data_as(_ip), status.ctypes.data_as(_up), _d(tpoint), C.byref(st)))
    res, o, q = [], 0, 0
    for nr, n, d in zip(nrates, npts, dims):
        m = (1 + nr) * n * fs.size
        res.append((out[o:o + m].reshape((1 + nr,) + d + (fs.size,)), ncyc[q:q + n].reshape(d),
                    status[q:q + n].reshape(d), tpoint[q:q + n].reshape(d)))
        o += m
        q += n
    return res, st.asdict()


def pmavg(a, Delta, Z, device=0, with_last=False):
    ''' Average intermolecular pressure (Pa) at deflections Z (sonic_pmavg: one QAGS run per deflection
        on the GPU, the sequence of rules scipy.integrate.quad applies in the reference). '''
    lib = load()
    Z = as_f64(np.atleast_1d(Z)).ravel()
    out = np.empty(Z.size)
    last = np.empty(Z.size, dtype=np.int32)
    check(lib.sonic_pmavg(device, float(a), float(Delta), Z.size, _d(Z), _d(out), last.ctypes.data_as(_ip)))
    return (out, last) if with_last else out


def simulate(neuron_id, Qref, tab_on, tab_off, t, stim_on, y0, nsub=64, device=0):
    ''' Batched SONIC simulations on 1-D lookups (sonic_simulate).
        :param tab_on: [nsim, 1 + 2 NS, nQ];  tab_off: [1 + 2 NS, nQ]
        :return: (out[nsim, nt, 1 + NS], status[nsim]) '''
    lib = load()
    Qref, tab_on, tab_off, t, y0 = as_f64(Qref), as_f64(tab_on), as_f64(tab_off), as_f64(t), as_f64(y0)
    stim_on = np.ascontiguousarray(stim_on, dtype=np.uint8)
    nsim, nv, nQ = tab_on.shape
    out = np.empty((nsim, t.size, y0.size))
    status = np.empty(nsim, dtype=np.int32)
    check(lib.sonic_simulate(device, neuron_id, nsim, nQ, _d(Qref), _d(tab_on), _d(tab_off), t.size, _d(t),
                             stim_on.ctypes.data_as(C.POINTER(C.c_uint8)), _d(y0), int(nsub), _d(out),
                             status.ctypes.data_as(_ip)))
    return out, status
