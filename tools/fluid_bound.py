# -*- coding: utf-8 -*-
''' Fluid lower bound of a lookup's integrator time on one B200 under the measured tick table: every chain runs in a
    warp with the largest number of busy lanes k that still meets the deadline (c t(k) <= T; a lone lane in the
    register-resident run if even k = 1 does not), warps are perfectly packed and chain lengths are known exactly.
    The smallest T whose warp-time fits into 1184 warps x T bounds what any schedule can reach with these tick costs.

        python tools/fluid_bound.py [profiles/r02_c2_chain_lengths.npz]

    (chain lengths = LSODA right-hand sides per unique trajectory of C2, measured on the device: tools/gpu_c2diag.py)
'''
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
TK_US = np.array([0, 1.645, 2.250, 2.701, 3.051, 3.32, 3.549, 3.73, 3.879, 3.99, 4.09, 4.18, 4.263, 4.33, 4.39, 4.45, 4.496,
                  4.52, 4.55, 4.57, 4.60, 4.62, 4.64, 4.66, 4.677, 4.70, 4.71, 4.73, 4.75, 4.76, 4.78, 4.79, 4.809])
WARPS = 148 * 8


def warp_time(c, T, t_lone=1.08, scale=1.0):
    kbest = np.zeros(len(c), int)
    for k in range(1, 33):
        kbest[c * TK_US[k] * scale <= T * 1e6] = k
    lone = kbest == 0
    if np.any(c[lone] * t_lone > T * 1e6):
        return np.inf, 0
    wt = np.where(lone, c * t_lone, c * TK_US[np.maximum(kbest, 1)] * scale / np.maximum(kbest, 1))
    return wt.sum() / 1e6, int(lone.sum())


def bound(c, scale=1.0):
    for T in np.arange(0.5, 3.0, 0.01):
        need, lone = warp_time(c, T, scale=scale)
        if need <= WARPS * T:
            return T, need, lone
    return None


if __name__ == '__main__':
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r02_c2_chain_lengths.npz')
    c = np.load(path)['nrhs'].astype(float)
    print(f'{len(c)} chains, {c.sum():.4g} right-hand sides, longest {c.max():.0f} ({c.max() * 1.02e-6:.2f} s alone at 1.02 us)')
    for scale in (1.0, 0.85, 0.7):
        T, need, lone = bound(c, scale)
        print(f'staged tick table x {scale}: fluid bound {T:.2f} s ({need:.0f} of {WARPS * T:.0f} warp-seconds, {lone} lone chains)')
