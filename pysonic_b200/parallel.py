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


def predicted_log_cost(a, f, A, Q=None):
    ''' Same ordering heuristic as the native work queue (sonic_b200.cu: predict_log_cost):
        nearest node of the measured cost table (tools/make_cost_table.py).  Without charges the
        envelope over all charges is returned. '''
    la, lf, Ae, Qe, tab = _cost_table()
    a, f, A = np.broadcast_arrays(np.asarray(a, float), np.asarray(f, float), np.asarray(A, float))
    i = np.argmin(np.abs(np.log(a)[..., None] - la), axis=-1)
    j = np.argmin(np.abs(np.log(f)[..., None] - lf), axis=-1)
    k = np.clip(np.digitize(A, Ae) - 1, 0, tab.shape[2] - 1)
    if Q is None:
        return tab.max(axis=3)[i, j, k]
    l = np.clip(np.digitize(np.abs(np.asarray(Q, float)) + 1e-14, Qe) - 1, 0, tab.shape[3] - 1)
    return tab[i, j, k, l]


def dist_info():
    ''' (rank, world_size, local_rank) of the current process; (0, 1, 0) outside torchrun. '''
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get('LOCAL_RANK', 0))
    except ImportError:
        pass
    return 0, 1, 0


def shard_indices(cost, rank, world_size):
    ''' Indices of the points owned by `rank`: cost-sorted (descending), dealt round-robin. '''
    order = np.argsort(-np.asarray(cost), kind='stable')
    return order[rank::world_size]


def gather_slabs(n, idx, arrays, rank, world_size):
    ''' Gather per-point outputs from all ranks into full arrays (on every rank).

        :param n: total number of points
        :param idx: indices owned by this rank
        :param arrays: list of arrays whose LAST-BUT-`k` axis layout is (..., len(idx), ...):
            each array is given as (array, axis) with `axis` the point axis
        :return: list of full arrays with `n` along the point axis
    '''
    import torch
    import torch.distributed as dist
    backend = dist.get_backend()
    dev = torch.device('cuda', torch.cuda.current_device()) if backend == 'nccl' else torch.device('cpu')
    # sizes differ by at most one between ranks: pad to the max
    m = int(len(idx))
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world_size)]
    dist.all_gather(sizes, torch.tensor([m], dtype=torch.int64, device=dev))
    sizes = [int(s.item()) for s in sizes]
    mmax = max(sizes)
    idx_pad = np.full(mmax, -1, dtype=np.int64)
    idx_pad[:m] = idx
    all_idx = [torch.zeros(mmax, dtype=torch.int64, device=dev) for _ in range(world_size)]
    dist.all_gather(all_idx, torch.from_numpy(idx_pad).to(dev))
    all_idx = [t.cpu().numpy() for t in all_idx]
    out = []
    for arr, axis in arrays:
        arr = np.moveaxis(np.asarray(arr), axis, 0)
        pad = np.zeros((mmax,) + arr.shape[1:], dtype=arr.dtype)
        pad[:m] = arr
        # ship raw bytes: NCCL has no unsigned 32-bit type (status words, RHS counts)
        t = torch.from_numpy(np.ascontiguousarray(pad).view(np.uint8).reshape(-1)).to(dev)
        parts = [torch.zeros_like(t) for _ in range(world_size)]
        dist.all_gather(parts, t)
        full = np.zeros((n,) + arr.shape[1:], dtype=arr.dtype)
        for r in range(world_size):
            k = sizes[r]
            part = parts[r].cpu().numpy().view(arr.dtype).reshape(pad.shape)
            full[all_idx[r][:k]] = part[:k]
        out.append(np.moveaxis(full, 0, axis))
    return out


def run_sharded(compute, n, cost, rank=None, world_size=None):
    ''' Run `compute(idx) -> list of (array, point_axis)` on this rank's shard and gather.

        `compute` receives the global indices of the points this rank owns and returns its
        per-point outputs; the function returns the gathered full-size arrays. '''
    r, w, _ = dist_info()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    idx = shard_indices(cost, rank, world_size)
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
