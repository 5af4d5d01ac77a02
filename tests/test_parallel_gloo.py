# -*- coding: utf-8 -*-
''' Multi-rank sharding of a lookup grid (pysonic_b200/parallel.py) with world_size = 2 on the
    gloo backend: the cost-sorted round-robin split and the final gather must reproduce the
    single-process result exactly, whatever the per-point compute is. '''

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def _fake_compute(a, f, A, Q, fs):
    ''' Deterministic stand-in for the GPU call: tables[nvar, n, nfs] + per-point outputs. '''
    n = f.size
    base = np.sin(f * 1e-6) + np.log1p(A) + Q * 1e3 + a * 1e9
    out = np.stack([np.outer(base * (v + 1), fs) for v in range(3)])       # (3, n, nfs)
    ncyc = (3 + (A == 0) * 8).astype(np.int32)
    tpoint = base * 1e-3
    status = (A == 0).astype(np.uint32)
    return out, ncyc, tpoint, status


def _worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from pysonic_b200.parallel import run_sharded, predicted_log_cost, dist_info
        assert dist_info() == (rank, world, rank)
        a, f, A, Q, fs = _inputs(n)
        cost = predicted_log_cost(a, f, A)

        def compute(idx):
            out, ncyc, tp, st = _fake_compute(a[idx], f[idx], A[idx], Q[idx], fs)
            return [(out, 1), (ncyc, 0), (tp, 0), (st, 0)]

        out, ncyc, tp, st = run_sharded(compute, n, cost)
        ret[rank] = (out, ncyc, tp, st)
    finally:
        dist.destroy_process_group()


def _inputs(n):
    rng = np.random.default_rng(7)
    a = rng.choice([16e-9, 32e-9, 64e-9], n)
    f = rng.choice([2e4, 1e5, 5e5, 1e6, 4e6], n)
    A = rng.choice([0., 1e3, 5e4, 3e5, 6e5], n)
    Q = rng.uniform(-1e-3, 5e-4, n)
    fs = np.array([0.25, 0.5, 1.0])
    return a, f, A, Q, fs


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('n', [7, 64])
def test_two_rank_gather_matches_single_process(n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    a, f, A, Q, fs = _inputs(n)
    ref_out, ref_ncyc, ref_tp, ref_st = _fake_compute(a, f, A, Q, fs)
    for r in range(world):
        out, ncyc, tp, st = ret[r]
        np.testing.assert_array_equal(st, ref_st)
        assert st.dtype == np.uint32
        np.testing.assert_array_equal(out, ref_out)
        np.testing.assert_array_equal(ncyc, ref_ncyc)
        np.testing.assert_array_equal(tp, ref_tp)
        assert ncyc.dtype == np.int32


def test_single_process_path():
    sys.path.insert(0, ROOT)
    from pysonic_b200.parallel import run_sharded, predicted_log_cost
    n = 11
    a, f, A, Q, fs = _inputs(n)

    def compute(idx):
        out, ncyc, tp, st = _fake_compute(a[idx], f[idx], A[idx], Q[idx], fs)
        return [(out, 1), (ncyc, 0), (tp, 0)]

    out, ncyc, tp = run_sharded(compute, n, predicted_log_cost(a, f, A))
    ref = _fake_compute(a, f, A, Q, fs)
    np.testing.assert_array_equal(out, ref[0])
    np.testing.assert_array_equal(ncyc, ref[1])


# ---------------------------------------------------------------------------------------------
# the command line under a launcher (torchrun sets RANK / WORLD_SIZE / LOCAL_RANK): main() joins the
# process group itself, every rank integrates its shard, rank 0 alone writes the pickle
# ---------------------------------------------------------------------------------------------
def _fake_points_run(device, bls, neuron_id, nrates, ia, f, A, Q, fs, overtones=None):
    n = f.size
    base = np.sin(f * 1e-6) + np.log1p(A) + Q * 1e3 + np.array([b['a'] for b in bls])[ia] * 1e9
    out = np.stack([np.outer(base * (v + 1), fs) for v in range(1 + nrates)])
    ncyc = (3 + (A == 0) * 8).astype(np.int32)
    status = (A == 0).astype(np.uint32)
    return out, ncyc, status, base * 1e-3, np.full(n, 7 + device, np.uint32), {'n_points': n}


def _main_worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from pysonic_b200 import _lib, run_lookups
    _lib.points_run = _fake_points_run            # no GPU here: stand-in for the native call
    run_lookups.main(['-n', 'RS', '-a', '16', '32', '-f', '500', '2000', '-A', '0', '50', '300',
                      '-Q', '-50', '0', '50', '--mpi', '-o', outdir, '-y'])
    assert not dist.is_initialized()              # main() leaves the group it created


def test_cli_under_launcher_shards_and_writes_once(tmp_path):
    sys.path.insert(0, ROOT)
    world = 2
    mp.spawn(_main_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    files = sorted(os.listdir(tmp_path))
    assert files == ['RS_lookups_fs1.00.pkl'], files
    import pickle
    with open(tmp_path / 'RS_lookups_fs1.00.pkl', 'rb') as fh:
        d = pickle.load(fh)
    assert list(d['refs']) == ['a', 'f', 'A', 'Q', 'fs']
    a, f, A, Q = d['refs']['a'], d['refs']['f'], d['refs']['A'], d['refs']['Q']
    ia, fi, Ai, Qi = [x.ravel() for x in np.meshgrid(np.arange(a.size), f, A, Q, indexing='ij')]
    import pysonic_b200 as ps
    pn = ps.getPointNeuron('RS')
    bls = [{'a': x} for x in a]
    ref = _fake_points_run(0, bls, 0, len(pn.rates), ia, fi, Ai, Qi, d['refs']['fs'])
    assert list(d['tables']) == ['V'] + pn.rates + ['tcomp']
    for v, k in enumerate(['V'] + pn.rates):
        np.testing.assert_array_equal(d['tables'][k], ref[0][v].reshape(2, 2, 3, 3, 1))
    np.testing.assert_array_equal(d['tables']['tcomp'][..., 0], ref[3].reshape(2, 2, 3, 3))
