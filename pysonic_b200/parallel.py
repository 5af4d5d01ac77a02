# -*- coding: utf-8 -*-
''' Multi-GPU sharding of a lookup grid: one process per GPU (torch.distributed), no data-path
    collective.  Grid points are independent, so the cost-sorted point list is dealt round-robin
    to the ranks (every rank gets the same cost profile), each rank integrates its slab on its
    own device, and the per-point outputs are gathered once at the end. '''

import json
import os

import numpy as np


_COST = None


def _cost_table():
    global _COST
    if _COST is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'cost_table.json')
        with open(path) as fh:
            t = json.load(fh)
        _COST = (np.log(np.array(t['a'])), np.log(np.array(t['f'])), np.array(t['A_edges_kPa']) * 1e3,
                 np.array(t['absQ_edges_nCcm2']) * 1e-5, np.array(t['log_cost']))
    return _COST


def _bracket(lx, nodes):
    ''' Bracketing node indices and weight of log-values `lx` on the ascending log-axis `nodes`
        (clamped outside). '''
    i1 = np.clip(np.searchsorted(nodes, lx, side='right'), 1, nodes.size - 1) if nodes.size > 1 else np.zeros(lx.shape, int)
    i0 = np.maximum(i1 - 1, 0)
    span = np.where(i1 > i0, nodes[i1] - nodes[i0], 1.0)
    w = np.clip((lx - nodes[i0]) / span, 0.0, 1.0)
    return i0, i1, w


def predicted_log_cost(a, f, A, Q=None):
    ''' Same ordering heuristic as the native work queue (sonic_b200.cu: predict_log_cost): the
        measured cost table (tools/make_cost_table.py), interpolated linearly in log(radius) and
        log(frequency) between its nodes.  Without charges the envelope over all charges is returned. '''
    la, lf, Ae, Qe, tab = _cost_table()
    a, f, A = np.broadcast_arrays(np.asarray(a, float), np.asarray(f, float), np.asarray(A, float))
    i0, i1, wi = _bracket(np.log(a), la)
    j0, j1, wj = _bracket(np.log(f), lf)
    k = np.clip(np.digitize(A, Ae) - 1, 0, tab.shape[2] - 1)
    if Q is None:
        t = tab.max(axis=3)
        at = lambda i, j: t[i, j, k]                        # noqa: E731
    else:
        l = np.clip(np.digitize(np.abs(np.asarray(Q, float)) + 1e-14, Qe) - 1, 0, tab.shape[3] - 1)
        at = lambda i, j: tab[i, j, k, l]                   # noqa: E731
    c0 = at(i0, j0) + wj * (at(i0, j1) - at(i0, j0))
    c1 = at(i1, j0) + wj * (at(i1, j1) - at(i1, j0))
    return c0 + wi * (c1 - c0)


_OWN_GROUP = False


def env_world():
    ''' (rank, world_size, local_rank) announced by the launcher (torchrun) in the environment. '''
    return (int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)),
            int(os.environ.get('LOCAL_RANK', 0)))


def init_from_env():
    ''' Under `torchrun` (RANK / WORLD_SIZE / LOCAL_RANK in the environment): bind this process to
        its GPU and join the process group (NCCL when a device is present, gloo otherwise), unless the
        caller has already done so.  Returns (rank, world_size, local_rank).  Outside torchrun: no-op. '''
    global _OWN_GROUP
    rank, world, local_rank = env_world()
    if world <= 1:
        return 0, 1, 0
    import torch
    import torch.distributed as dist
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if cuda:
            dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group('gloo')
        _OWN_GROUP = True
    return rank, world, local_rank


def finalize():
    ''' Leave the process group if `init_from_env` created it. '''
    global _OWN_GROUP
    if _OWN_GROUP:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
        _OWN_GROUP = False


def dist_info():
    ''' (rank, world_size, local_rank) of the current process; (0, 1, 0) outside a process group. '''
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get('LOCAL_RANK', 0))
    except ImportError:
        pass
    return 0, 1, 0


def trajectory_groups(ia, f, A, Q):
    ''' Group id of every point: points that differ by the sign of the charge only share one
        trajectory (the mechanics see Q^2, bls.py:482-491) and must stay on the same device so that
        the engine integrates it once.  |Q| is rounded to 36 significant bits exactly as
        sonic_plan_create_ex does. '''
    q = np.abs(np.ascontiguousarray(Q, dtype=np.float64))
    qb = (q.view(np.uint64) + np.uint64(0x8000)) & ~np.uint64(0xFFFF)
    # one integer id per column (few distinct values each), combined into a single key
    key = np.zeros(q.size, dtype=np.int64)
    for col in (np.asarray(ia).astype(np.int64), np.ascontiguousarray(f, dtype=np.float64).view(np.int64),
                np.ascontiguousarray(A, dtype=np.float64).view(np.int64), qb.view(np.int64)):
        vals, inv = np.unique(col, return_inverse=True)
        key = key * vals.size + inv.ravel()
    _, inv = np.unique(key, return_inverse=True)
    return inv.ravel()


def shard_indices(cost, rank, world_size, groups=None):
    ''' Indices of the points owned by `rank`: cost-sorted (descending), dealt round-robin.  With
        `groups` (one id per point, equal cost inside a group) whole groups are dealt, so that the
        +Q / -Q points of a trajectory land on the same rank. '''
    cost = np.asarray(cost)
    if groups is None:
        order = np.argsort(-cost, kind='stable')
        return order[rank::world_size]
    groups = np.asarray(groups)
    first = np.full(groups.max() + 1, -1, dtype=np.int64)
    first[groups[::-1]] = np.arange(groups.size)[::-1]          # first point of every group
    gorder = np.argsort(-cost[first], kind='stable')            # groups, most expensive first
    owner = np.empty(gorder.size, dtype=np.int64)
    owner[gorder] = np.arange(gorder.size) % world_size
    idx = np.nonzero(owner[groups] == rank)[0]
    return idx[np.argsort(-cost[idx], kind='stable')]


def gather_slabs(n, idx, arrays, rank, world_size):
    ''' Gather per-point outputs from all ranks into full arrays (on every rank): ONE all_gather of a
        byte matrix whose rows are [global index | the point's slice of every array].

        :param n: total number of points
        :param idx: indices owned by this rank
        :param arrays: list of (array, axis) with `axis` the point axis of the array
        :return: list of full arrays with `n` along the point axis
    '''
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device('cuda', int(os.environ.get('LOCAL_RANK', torch.cuda.current_device()))) if backend == 'nccl' \
        else torch.device('cpu')
    m = int(len(idx))
    # rows: raw bytes (NCCL has no unsigned 32-bit type for the status words and RHS counts)
    cols = [np.ascontiguousarray(np.asarray(idx, dtype=np.int64)).view(np.uint8).reshape(m, 8)]
    moved = []
    for arr, axis in arrays:
        a = np.ascontiguousarray(np.moveaxis(np.asarray(arr), axis, 0))
        moved.append(a)
        cols.append(a.view(np.uint8).reshape(m, -1) if m else np.zeros((0, a[0:1].nbytes if a.size else a.dtype.itemsize * int(np.prod(a.shape[1:]))), np.uint8))
    rowbytes = sum(c.shape[1] for c in cols)
    # shard sizes differ by a few points at most (whole trajectory groups are dealt): pad to the largest
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world_size)]
    dist.all_gather(sizes, torch.tensor([m], dtype=torch.int64, device=dev))
    sizes = [int(x.item()) for x in sizes]
    mmax = max(sizes)
    mat = np.zeros((mmax, rowbytes), dtype=np.uint8)
    if m:
        mat[:m] = np.concatenate(cols, axis=1)
    t = torch.from_numpy(mat.reshape(-1)).to(dev)
    parts = [torch.empty_like(t) for _ in range(world_size)]
    dist.all_gather(parts, t)
    parts = [x.cpu().numpy().reshape(mmax, rowbytes) for x in parts]
    out = []
    off = 8
    for (arr, axis), a in zip(arrays, moved):
        width = int(np.prod(a.shape[1:])) * a.dtype.itemsize
        full = np.zeros((n,) + a.shape[1:], dtype=a.dtype)
        for r in range(world_size):
            k = sizes[r]
            if k == 0:
                continue
            gidx = np.ascontiguousarray(parts[r][:k, :8]).view(np.int64).ravel()
            full[gidx] = np.ascontiguousarray(parts[r][:k, off:off + width]).view(a.dtype).reshape((k,) + a.shape[1:])
        out.append(np.moveaxis(full, 0, axis))
        off += width
    return out


def run_sharded(compute, n, cost, rank=None, world_size=None, groups=None):
    ''' Run `compute(idx) -> list of (array, point_axis)` on this rank's shard and gather.

        `compute` receives the global indices of the points this rank owns and returns its
        per-point outputs; the function returns the gathered full-size arrays. '''
    r, w, _ = dist_info()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    idx = shard_indices(cost, rank, world_size, groups)
    outs = compute(idx)
    if world_size == 1:
        full = []
        for arr, axis in outs:
            arr = np.moveaxis(np.asarray(arr), axis, 0)
            f = np.zeros((n,) + arr.shape[1:], dtype=arr.dtype)
            f[idx] = arr
            full.append(np.moveaxis(f, 0, axis))
        return full
    return gather_slabs(n, idx, outs, rank, world_size)
