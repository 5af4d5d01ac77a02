# -*- coding: utf-8 -*-
''' ORACLE -- CPU restatement of the reference's SONIC lookup-generation path.

    THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
    the `cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  The product path
    (`pysonic_b200`) never imports, calls or links anything under `oracle/`.

    What it restates (citations relative to /root/reference):

    ============================  =====================================================
    this file                     reference
    ============================  =====================================================
    `BlsConsts`                   `PySONIC/core/bls.py:87-137` (constants, ng0, V0)
    `derivatives`                 `bls.py:681-718` + helpers `:286-319,472-526,575-655`,
                                  `drives.py:303-304`
    `balancedef_qs`               `bls.py:538-573,720-725`
    `sim_cycles`                  `bls.py:749-789`, `solvers.py:77-170,283-365`
    `capacitance`                 `bls.py:330-349`
    `compute_effvars`             `nbls.py:148-222`, `pneuron.py:268-271`
    `NEURONS` rate functions      `neurons/*.py`, helpers `pneuron.py:351-413`,
                                  alpha/beta from xinf/tau: `translators.py:317-324`
    `compute_astim_lookup`        `scripts/run_lookups.py:22-175`, `batches.py:135-171`
    `rel_cm_cycle`, `compute_cm_lookup`  `bls.py:801-808`, `scripts/run_Cm_lookups.py:19-64`
    ============================  =====================================================

    The numerical integrator is the reference's own third-party dependency: scipy's ODEPACK
    LSODA through `scipy.integrate.odeint` (call site `solvers.py:167`; `requirements.txt:2`
    says `scipy>=0.17`, unpinned; the image used here and on the GPU box has scipy 1.18.1).
    It is called with the reference's arguments (default rtol = atol = 1.49012e-8,
    `tfirst=True`, no Jacobian, 1000-sample `linspace` per cycle).

    PARITY PIN: `tests/test_oracle_golden.py` checks this oracle against golden vectors generated
    by importing and running the unmodified reference in the build container
    (`tests/golden/make_goldens.py` -> `tests/golden/*.json`).  The arithmetic below follows the
    reference's operation order so that, with the same numpy/scipy, results agree to the last
    bits (the test asserts 1e-12 relative).
'''

import json
import os
import time

import numpy as np
from scipy.integrate import odeint
from scipy.optimize import brentq

# ----------------------------------------------------------------------------------------------
# Constants (reference: PySONIC/constants.py:13,27,31,34-38 and bls.py:87-110)
# ----------------------------------------------------------------------------------------------
Rg = 8.31342
FARADAY = 9.64853e4
CELSIUS_2_KELVIN = 273.15
NPC_DENSE = 1000
NCYCLES_MAX = 10
MAX_RMSE_PTP_RATIO = 1e-4
DQ_LOOKUP = 1e-5

T_BLS = 309.15
delta0 = 2.0e-9
rhoL = 1075.0
muL = 7.0e-4
muS = 0.035
kA = 0.24
alpha_tissue = 7.56
C0 = 0.62
kH = 1.613e5
P0 = 1.0e5
Dgl = 3.68e-9
xi = 0.5e-9
epsilon0 = 8.854e-12
epsilonR = 1.0
rel_Zmin = -0.49

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pysonic_b200', 'data',
                     'bls_params.json')


class BlsConsts:
    ''' Per-(radius, neuron) constants of the bilayer sonophore (bls.py:115-137). '''

    def __init__(self, a, Cm0, Qm0, Delta, lj, embedding_depth=0.0):
        self.a = a
        self.Cm0 = Cm0
        self.Qm0 = Qm0
        self.d = embedding_depth
        self.S0 = np.pi * self.a**2
        self.kA_tissue = 0.
        self.Delta = Delta
        self.x0, self.C, self.nrep, self.nattr = lj
        self.V0 = np.pi * self.Delta * self.a**2
        self.ng0 = P0 * self.V0 / (Rg * T_BLS)
        self.Zmin = rel_Zmin * self.Delta

    @classmethod
    def from_table(cls, a, Cm0, Qm0, path=_DATA):
        ''' Load Delta and the Lennard-Jones fit from the input table (bls.py:49-75 keys). '''
        akey = f'{a * 1e9:.1f}'
        qkey = f'{Qm0 * 1e5:.2f}'
        with open(path) as fh:
            rows = json.load(fh)['rows']
        for r in rows:
            if r[0] == akey and r[1] == qkey:
                return cls(a, Cm0, Qm0, r[2], tuple(r[3:7]))
        raise KeyError(f'no bilayer-sonophore constants for a={akey} nm, Qm0={qkey} nC/cm2')


# ----------------------------------------------------------------------------------------------
# Mechanics (bls.py)
# ----------------------------------------------------------------------------------------------

def curvrad(b, Z):
    if Z == 0.0:
        return np.inf
    return (b.a**2 + Z**2) / (2 * Z)


def surface(b, Z):
    return np.pi * (b.a**2 + Z**2)


def volume(b, Z):
    return np.pi * b.a**2 * b.Delta * (1 + (Z / (3 * b.Delta) * (3 + Z**2 / b.a**2)))


def gasmol2Pa(ng, V):
    return ng * Rg * T_BLS / V


def pm_avg_pred(b, Z):
    ''' Lennard-Jones approximation of the average intermolecular pressure
        (bls.py:29-41, 472-480). '''
    x = b.x0 / (2 * Z + b.Delta)
    return b.C * (np.power(x, b.nrep) - np.power(x, b.nattr))


def pelec(b, Z, Qm):
    relS = b.S0 / surface(b, Z)
    abs_perm = epsilon0 * epsilonR
    return - relS * Qm**2 / (2 * abs_perm)


def pac(f, A, t, phi=np.pi):
    return A * np.sin(2 * np.pi * f * t - phi)


def derivatives(t, y, b, f, A, Qm):
    ''' RHS of the 3-state mechanical system (bls.py:681-718). '''
    U, Z, ng = y
    if Z < b.Zmin:
        Z = b.Zmin
    R = curvrad(b, Z)
    Pg = gasmol2Pa(ng, volume(b, Z))
    Pm = pm_avg_pred(b, Z)
    Pac = pac(f, A, t)
    Pv = (- 12 * U * delta0 * muS / R**2) + (- 4 * U * muL / np.abs(R))
    strain = (Z / b.a)**2
    PE = - (kA * strain + b.kA_tissue * strain) / R
    Ptot = Pm + Pg - P0 - Pac + PE + Pv + pelec(b, Z, Qm)
    dUdt = Ptot / (rhoL * np.abs(R)) + (-(3 * U**2) / (2 * R))
    dZdt = U
    dngdt = 2 * surface(b, Z) * Dgl * (C0 - Pg / kH) / xi
    return [dUdt, dZdt, dngdt]


def ptot_qs(Z, b, ng, Qm, Pac):
    ''' Quasi-static net pressure (bls.py:538-553). '''
    return pm_avg_pred(b, Z) + gasmol2Pa(ng, volume(b, Z)) - P0 - Pac + pelec(b, Z, Qm)


def balancedef_qs(b, ng, Qm, Pac):
    ''' Quasi-static equilibrium deflection (bls.py:555-573). '''
    lo, hi = b.Zmin, b.a
    plo, phi_ = ptot_qs(lo, b, ng, Qm, Pac), ptot_qs(hi, b, ng, Qm, Pac)
    if not (plo > 0 > phi_):
        raise ValueError(f'P_QS not changing sign within [{lo * 1e9:.2f}, {hi * 1e9:.2f}] nm')
    return brentq(ptot_qs, lo, hi, args=(b, ng, Qm, Pac), xtol=1e-16)


def capacitance(b, Z):
    ''' Membrane capacitance (bls.py:330-345). '''
    if Z == 0.0:
        return b.Cm0
    Z2 = (b.a**2 - Z**2 - Z * b.Delta) / (2 * Z)
    return b.Cm0 * b.Delta / b.a**2 * (Z + Z2 * np.log((2 * Z + b.Delta) / b.Delta))


def v_capacitance(b, Z):
    return np.array([capacitance(b, z) for z in Z])


# ----------------------------------------------------------------------------------------------
# Periodic solver (solvers.py)
# ----------------------------------------------------------------------------------------------

def _rmse(x1, x2, axis=None):
    return np.sqrt(np.mean((x1 - x2)**2, axis=axis))  # utils.py:185-187


def sim_cycles(b, f, A, Qm, nmax=NCYCLES_MAX, nmin=2, stats=None, rtol=None, atol=None):
    ''' Cycle-by-cycle integration until periodic stabilisation
        (bls.py:749-789 driving solvers.py:336-365).

        :return: (t, y[:, (U, Z, ng)], ncycles) with 2 + 999 * ncycles rows
    '''
    b.kA_tissue = 2 * (alpha_tissue * f) * b.d           # bls.py:583-586
    T = 1. / f
    dt = 1 / (NPC_DENSE * f)
    tall = np.ones(2) * 0.
    nfe = [0]

    if isinstance(Qm, float):
        Qm0, Qm_t = Qm, lambda t: Qm                                  # bls.py:763-765
    else:
        Qm0, Qm_t = Qm[0], lambda t: Qm[int((t % T) / dt)]            # bls.py:766-768
    Z0 = balancedef_qs(b, b.ng0, Qm0, pac(f, A, dt))      # bls.py:720-725,771-772
    yall = np.array([[0., 0.], [0., Z0], [b.ng0, b.ng0]]).T   # solvers.py:111-115

    def dfunc(t, y):
        nfe[0] += 1
        return derivatives(t, y, b, f, A, Qm_t(t))

    kw = {}
    if rtol is not None:
        kw['rtol'] = rtol
    if atol is not None:
        kw['atol'] = atol
    infos = []

    def integrate_cycle(tall, yall):
        t0 = tall[-1]
        target = t0 + T
        nsamples = max(int(np.round((target - t0) / dt)), 2)   # solvers.py:77-87
        tv = np.linspace(t0, target, nsamples)                 # solvers.py:97
        if stats is not None:
            y, info = odeint(dfunc, yall[-1], tv, tfirst=True, full_output=True, **kw)
            infos.append(info)
        else:
            y = odeint(dfunc, yall[-1], tv, tfirst=True, **kw)  # solvers.py:167
        return np.concatenate((tall, tv[1:])), np.concatenate((yall, y[1:]), axis=0)

    def stable(tall, yall):
        # solvers.py:283-330 (cycle extraction reduces to the last two 999-sample blocks)
        i_diff_dt = np.where(np.invert(np.isclose(np.diff(tall)[::-1], dt)))[0]
        nsamples = i_diff_dt[0] if i_diff_dt.size > 0 else tall.size
        npc = int(np.round(T / (tall[-1] - tall[-2])))
        ncyc = int(np.round(nsamples / npc))
        ioff = tall.size - npc * ncyc
        cyc = []
        for i in (1, 2):
            ii = ncyc - i
            if ii < 0 or ii >= ncyc:
                raise ValueError('Invalid index')
            s = ii * npc + ioff
            cyc.append(yall[s:s + npc, 1:3])
        y_last, y_prec = cyc
        with np.errstate(divide='ignore', invalid='ignore'):
            ratios = _rmse(y_last, y_prec, axis=0) / np.ptp(y_last, axis=0)
        return np.all(ratios < MAX_RMSE_PTP_RATIO)

    for i in range(nmin):
        tall, yall = integrate_cycle(tall, yall)
    while not stable(tall, yall) and i < nmax:
        tall, yall = integrate_cycle(tall, yall)
        i += 1
    ncycles = (tall.size - 2) // (NPC_DENSE - 1)
    if stats is not None:
        stats['nfe'] = nfe[0]
        stats['infos'] = infos
        stats['Z0'] = Z0
    return tall, yall, ncycles


# ----------------------------------------------------------------------------------------------
# Neuron rate functions (neurons/*.py).  Each entry: name -> (Cm0, Vm0, [(rate, func), ...])
# with funcs operating elementwise on a numpy array of membrane potentials (mV), result in s-1.
# ----------------------------------------------------------------------------------------------

def vtrap(x, y):
    return x / (np.exp(x / y) - 1)   # pneuron.py:351-354 (naive 0/0 at x = 0 kept on purpose)


def _ab_from_inf_tau(xinf, tau):
    ''' translators.py:317-324 '''
    return (lambda V: xinf(V) / tau(V)), (lambda V: (1 - xinf(V)) / tau(V))


def _pospischil_mhn(VT):
    ''' cortical.py:36-58 / thalamic.py:37-59 '''
    return [
        ('alpham', lambda V: 0.32 * vtrap(13 - (V - VT), 4) * 1e3),
        ('betam', lambda V: 0.28 * vtrap((V - VT) - 40, 5) * 1e3),
        ('alphah', lambda V: 0.128 * np.exp(-((V - VT) - 17) / 18) * 1e3),
        ('betah', lambda V: 4 / (1 + np.exp(-((V - VT) - 40) / 5)) * 1e3),
        ('alphan', lambda V: 0.032 * vtrap(15 - (V - VT), 5) * 1e3),
        ('betan', lambda V: 0.5 * np.exp(-((V - VT) - 10) / 40) * 1e3),
    ]


def _cortical_p(TauMax):
    ''' cortical.py:60-66 '''
    pinf = lambda V: 1.0 / (1 + np.exp(-(V + 35) / 10))
    taup = lambda V: TauMax / (3.3 * np.exp((V + 35) / 20) + np.exp(-(V + 35) / 20))
    a, bb = _ab_from_inf_tau(pinf, taup)
    return [('alphap', a), ('betap', bb)]


def _lts_su(Vx):
    ''' cortical.py:249-266 / thalamic.py:287-305 '''
    sinf = lambda V: 1.0 / (1.0 + np.exp(-(V + Vx + 57.0) / 6.2))

    def taus(V):
        x = np.exp(-(V + Vx + 132.0) / 16.7) + np.exp((V + Vx + 16.8) / 18.2)
        return 1.0 / 3.7 * (0.612 + 1.0 / x) * 1e-3

    uinf = lambda V: 1.0 / (1.0 + np.exp((V + Vx + 81.0) / 4.0))

    def tauu(V):
        V = np.asarray(V, dtype=float)
        lo = 1.0 / 3.7 * np.exp((V + Vx + 467.0) / 66.6) * 1e-3
        hi = 1.0 / 3.7 * (np.exp(-(V + Vx + 22) / 10.5) + 28.0) * 1e-3
        return np.where(V + Vx < -80.0, lo, hi)

    a_s, b_s = _ab_from_inf_tau(sinf, taus)
    a_u, b_u = _ab_from_inf_tau(uinf, tauu)
    return [('alphas', a_s), ('betas', b_s), ('alphau', a_u), ('betau', b_u)]


def _re_su():
    ''' thalamic.py:164-179 '''
    sinf = lambda V: 1.0 / (1.0 + np.exp(-(V + 52.0) / 7.4))
    taus = lambda V: (1 + 0.33 / (np.exp((V + 27.0) / 10.0) + np.exp(-(V + 102.0) / 15.0))) * 1e-3
    uinf = lambda V: 1.0 / (1.0 + np.exp((V + 80.0) / 5.0))
    tauu = lambda V: (28.3 + 0.33 / (
        np.exp((V + 48.0) / 4.0) + np.exp(-(V + 407.0) / 50.0))) * 1e-3
    a_s, b_s = _ab_from_inf_tau(sinf, taus)
    a_u, b_u = _ab_from_inf_tau(uinf, tauu)
    return [('alphas', a_s), ('betas', b_s), ('alphau', a_u), ('betau', b_u)]


def _tc_o():
    ''' thalamic.py:307-321 '''
    oinf = lambda V: 1.0 / (1.0 + np.exp((V + 75.0) / 5.5))
    tauo = lambda V: 1 / (np.exp(-14.59 - 0.086 * V) + np.exp(-1.87 + 0.0701 * V)) * 1e-3
    return [('alphao', lambda V: oinf(V) / tauo(V)), ('betao', lambda V: (1 - oinf(V)) / tauo(V))]


def _ib_qr():
    ''' cortical.py:349-363 '''
    return [
        ('alphaq', lambda V: 0.055 * vtrap(-(V + 27), 3.8) * 1e3),
        ('betaq', lambda V: 0.94 * np.exp(-(V + 75) / 17) * 1e3),
        ('alphar', lambda V: 0.000457 * np.exp(-(V + 13) / 50) * 1e3),
        ('betar', lambda V: 0.0065 / (np.exp(-(V + 15) / 28) + 1) * 1e3),
    ]


def _stn():
    ''' stn.py:59-152 (constants), :211-336 (kinetics), derStates order :345-359 '''
    xinf = lambda th, k: (lambda V: 1 / (1 + np.exp((V - th) / k)))
    taux1 = lambda th, sg, t0, t1: (lambda V: t0 + t1 / (1 + np.exp(-(V - th) / sg)))
    taux2 = lambda th1, th2, s1, s2, t0, t1: (
        lambda V: t0 + t1 / (np.exp(-(V - th1) / s1) + np.exp(-(V - th2) / s2)))
    gates = [
        ('a', xinf(-45, -14.7), taux1(-40, -0.5, 1e-3, 1e-3)),
        ('b', xinf(-90, 7.5), taux2(-60, -40, -30, 10, 0e-3, 200e-3)),
        ('c', xinf(-30.6, -5), taux2(-27, -50, -20, 15, 45e-3, 10e-3)),
        ('d1', xinf(-60, 7.5), taux2(-40, -20, -15, 20, 400e-3, 500e-3)),
        ('m', xinf(-40, -8), taux1(-53, -0.7, 0.2e-3, 3e-3)),
        ('h', xinf(-45.5, 6.4), taux2(-50, -50, -15, 16, 0e-3, 24.5e-3)),
        ('n', xinf(-41, -14), taux2(-40, -40, -40, 50, 0e-3, 11e-3)),
        ('p', xinf(-56, -6.7), taux2(-27, -102, -10, 15, 5e-3, 0.33e-3)),
        ('q', xinf(-85, 5.8), taux2(-50, -50, -15, 16, 0e-3, 400e-3)),
    ]
    out = []
    for k, xi_, tau in gates:
        a, bb = _ab_from_inf_tau(xi_, tau)
        out += [(f'alpha{k}', a), (f'beta{k}', bb)]
    return out


def _fh():
    ''' fh.py:61-103 '''
    q10 = 3**((36.0 - 20.0) / 10)
    Vm0 = -70.
    return [
        ('alpham', lambda V: q10 * 0.36 * vtrap(22. - (V - Vm0), 3.) * 1e3),
        ('betam', lambda V: q10 * 0.4 * vtrap(V - Vm0 - 13., 20.) * 1e3),
        ('alphah', lambda V: q10 * 0.1 * vtrap(V - Vm0 + 10.0, 6.) * 1e3),
        ('betah', lambda V: q10 * 4.5 / (np.exp((45. - (V - Vm0)) / 10.) + 1) * 1e3),
        ('alphan', lambda V: q10 * 0.02 * vtrap(35. - (V - Vm0), 10.0) * 1e3),
        ('betan', lambda V: q10 * 0.05 * vtrap(V - Vm0 - 10., 10.) * 1e3),
        ('alphap', lambda V: q10 * 0.006 * vtrap(40. - (V - Vm0), 10.0) * 1e3),
        ('betap', lambda V: q10 * 0.09 * vtrap(V - Vm0 + 25., 20.) * 1e3),
    ]


def _sw():
    ''' sweeney.py:52-66 '''
    alpham = lambda V: (126 + 0.363 * V) / (1 + np.exp(-(V + 49) / 5.3)) * 1e3
    betah = lambda V: 15.6 / (1 + np.exp(-(V + 56) / 10)) * 1e3
    return [
        ('alpham', alpham),
        ('betam', lambda V: alpham(V) / (np.exp((V + 56.2) / 4.17))),
        ('alphah', lambda V: betah(V) / np.exp((V + 74.5) / 5)),
        ('betah', betah),
    ]


def _mrg():
    ''' mrg.py:59-115 '''
    q10_mp = 2.2**((36.0 - 20.0) / 10)
    q10_h = 2.9**((36.0 - 20.0) / 10)
    q10_s = 3.0**((36.0 - 36.0) / 10)
    sh = 3.
    vtraub = -80.
    return [
        ('alpham', lambda V: q10_mp * 1.86 * vtrap(-((V + sh) + 18.4), 10.3) * 1e3),
        ('betam', lambda V: q10_mp * 0.086 * vtrap((V + sh) + 22.7, 9.16) * 1e3),
        ('alphah', lambda V: q10_h * 0.062 * vtrap((V + sh) + 111.0, 11.0) * 1e3),
        ('betah', lambda V: q10_h * 2.3 / (1 + np.exp(-((V + sh) + 28.8) / 13.4)) * 1e3),
        ('alphap', lambda V: q10_mp * 0.01 * vtrap(-(V + 27.), 10.2) * 1e3),
        ('betap', lambda V: q10_mp * 0.00025 * vtrap(V + 34., 10.) * 1e3),
        ('alphas', lambda V: q10_s * 0.3 / (1 + np.exp(-((V - vtraub) - 27.) / 5.)) * 1e3),
        ('betas', lambda V: q10_s * 0.03 / (1 + np.exp(-((V - vtraub) + 10.) / 1.)) * 1e3),
    ]


def _su():
    ''' sundt.py:60-123, Borg-Graham helpers pneuron.py:377-413 '''
    q10_Traub = 3**((36.0 - 30.0) / 10)
    q10_BG = 3**((36.0 - 30.0) / 10)
    Vrest = -65.
    msh, hsh = -6.0, 6.0
    T = 36.0 + CELSIUS_2_KELVIN
    xBG = lambda Vref, V: (V - Vref) * FARADAY / (Rg * T) * 1e-3
    aBG = lambda a0, zeta, gamma, Vref, V: a0 * np.exp(-zeta * gamma * xBG(Vref, V))
    bBG = lambda b0, zeta, gamma, Vref, V: b0 * np.exp(zeta * (1 - gamma) * xBG(Vref, V))
    return [
        ('alpham', lambda V: q10_Traub * 0.32 * vtrap((13.1 - ((V - Vrest) + msh)), 4) * 1e3),
        ('betam', lambda V: q10_Traub * 0.28 * vtrap((((V - Vrest) + msh) - 40.1), 5) * 1e3),
        ('alphah', lambda V: q10_Traub * 0.128 * np.exp((17.0 - ((V - Vrest) + hsh)) / 18) * 1e3),
        ('betah', lambda V: q10_Traub * 4 / (1 + np.exp((40.0 - ((V - Vrest) + hsh)) / 5)) * 1e3),
        ('alphan', lambda V: q10_BG * aBG(0.03, -5, 0.4, -32., V) * 1e3),
        ('betan', lambda V: q10_BG * bBG(0.03, -5, 0.4, -32., V) * 1e3),
        ('alphal', lambda V: q10_BG * aBG(0.001, 2, 1., -61., V) * 1e3),
        ('betal', lambda V: q10_BG * bBG(0.001, 2, 1., -61., V) * 1e3),
    ]


def _hh():
    ''' hh.py:31,46-75 '''
    q10 = 3**((36.0 - 6.3) / 10.)
    return [
        ('alpham', lambda V: q10 * 0.1 * vtrap(-(V + 40), 10) * 1e3),
        ('betam', lambda V: q10 * 4 * np.exp(-(V + 65) / 18) * 1e3),
        ('alphah', lambda V: q10 * 0.07 * np.exp(-(V + 65) / 20) * 1e3),
        ('betah', lambda V: q10 * 1.0 / (np.exp(-(V + 35) / 10) + 1) * 1e3),
        ('alphan', lambda V: q10 * 0.01 * vtrap(-(V + 55), 10) * 1e3),
        ('betan', lambda V: q10 * 0.125 * np.exp(-(V + 65) / 80) * 1e3),
    ]


def _leech_t():
    ''' leech.py:74-147: xinf = 1 / (1 + exp((V - half) / slope))^power; taux constant or sigmoidal '''
    xinf = lambda half, slope, power: (lambda V: 1 / (1 + np.exp((V - half) / slope))**power)
    taux = lambda half, slope, tmax, tmin: (lambda V: (tmax - tmin) / (1 + np.exp((V - half) / slope)) + tmin)
    const = lambda c: (lambda V: c + 0 * V)
    out = []
    for key, inf, tau in [('m', xinf(-35.0, -5.0, 1), const(0.1e-3)),
                          ('h', xinf(-50.0, 9.0, 2), taux(-36.0, 3.5, 14.0e-3, 0.2e-3)),
                          ('n', xinf(-22.0, -9.0, 1), taux(-10.0, 10.0, 6.0e-3, 1.0e-3)),
                          ('s', xinf(-10.0, -2.8, 1), const(0.6e-3))]:
        a, b = _ab_from_inf_tau(inf, tau)
        out += [('alpha' + key, a), ('beta' + key, b)]
    return out


def _leech_p():
    ''' leech.py:258-296 '''
    return [
        ('alpham', lambda V: -0.03 * (V + 28) / (np.exp(- (V + 28) / 15) - 1) * 1e3),
        ('betam', lambda V: 2.7 * np.exp(-(V + 53) / 18) * 1e3),
        ('alphah', lambda V: 0.045 * np.exp(-(V + 58) / 18) * 1e3),
        ('betah', lambda V: 0.72 / (np.exp(-(V + 23) / 14) + 1) * 1e3),
        ('alphan', lambda V: -0.024 * (V - 17) / (np.exp(-(V - 17) / 8) - 1) * 1e3),
        ('betan', lambda V: 0.2 * np.exp(-(V + 48) / 35) * 1e3),
        ('alphas', lambda V: -1.5 * (V - 20) / (np.exp(-(V - 20) / 5) - 1) * 1e3),
        ('betas', lambda V: 1.5 * np.exp(-(V + 25) / 10) * 1e3),
    ]


NEURONS = {
    # name: (Cm0 [F/m2], Vm0 [mV], ordered rate list)
    'RS': (1e-2, -71.9, _pospischil_mhn(-56.2) + _cortical_p(0.608)),
    'FS': (1e-2, -71.4, _pospischil_mhn(-57.9) + _cortical_p(0.502)),
    'LTS': (1e-2, -54.0, _pospischil_mhn(-50.0) + _cortical_p(4.0) + _lts_su(-7.0)),
    'IB': (1e-2, -71.4, _pospischil_mhn(-56.2) + _cortical_p(0.608) + _ib_qr()),
    'RE': (1e-2, -89.5, _pospischil_mhn(-67.0) + _re_su()),
    'TC': (1e-2, -61.93, _pospischil_mhn(-52.0) + _lts_su(0.0) + _tc_o()),
    'STN': (1e-2, -58.0, _stn()),
    'FHnode': (2e-2, -70., _fh()),
    'SWnode': (2.5e-2, -80.0, _sw()),
    'MRGnode': (2e-2, -80., _mrg()),
    'SUseg': (1e-2, -60., _su()),
    'HHseg': (1e-2, -65.0, _hh()),
    'LeechT': (1e-2, -53.58, _leech_t()),
    'LeechP': (1e-2, -48.865, _leech_p()),
    'template': (1e-2, -71.9, _pospischil_mhn(-56.2)),
    'pas': (1e-2, -70., []),          # passive membrane (pas.py:103-107): no gate
}


def neuron_Qm0(name):
    Cm0, Vm0, _ = NEURONS[name]
    return Cm0 * Vm0 * 1e-3       # pneuron.py:62-63


def neuron_Qbounds(name):
    Cm0, Vm0, _ = NEURONS[name]
    return np.array([np.round(Vm0 - 35.0), 50.0]) * Cm0 * 1e-3   # pneuron.py:423-426


def rate_names(name):
    return [k for k, _ in NEURONS[name][2]]


def eff_rates(name, Vm):
    ''' pneuron.py:268-271 '''
    with np.errstate(all='ignore'):
        return {k: np.mean(fn(Vm)) for k, fn in NEURONS[name][2]}


# ----------------------------------------------------------------------------------------------
# Effective variables (nbls.py:153-222) and grid driver (run_lookups.py:22-175)
# ----------------------------------------------------------------------------------------------

def get_bls(name, a):
    Cm0, Vm0, _ = NEURONS[name]
    return BlsConsts.from_table(a, Cm0, neuron_Qm0(name))


def compute_effvars(name, b, f, A, fs, Qm, stats=None, Qm_overtones=None, **kw):
    ''' nbls.py:153-222.  Qm_overtones: list of (amplitude, phase) pairs or None.
        :return: ([{V, [A_V1, phi_V1, ...], rates...} per fs], ncycles) '''
    fs = np.atleast_1d(np.asarray(fs, dtype=float))
    if Qm_overtones is None:
        Qm_cycle = Qm                                     # nbls.py:169-172
        novertones = 0
    else:
        A_Qm, phi_Qm = list(zip(*Qm_overtones))           # nbls.py:173-178
        Qm_fft = np.hstack(([Qm + 0j], A_Qm * (np.cos(phi_Qm) + 1j * np.sin(phi_Qm))))
        Qm_cycle = np.fft.irfft(Qm_fft, n=NPC_DENSE) * NPC_DENSE
        novertones = len(A_Qm)
    tall, yall, ncycles = sim_cycles(b, f, A, Qm_cycle, stats=stats, **kw)
    Z_cycle = yall[-NPC_DENSE:, 1]                    # nbls.py:181
    Cm_cycle = v_capacitance(b, Z_cycle)              # nbls.py:182
    out = []
    for x in fs:
        Vm_cycle = Qm_cycle / (x * Cm_cycle + (1 - x) * b.Cm0) * 1e3   # nbls.py:148-151,188
        ev = {'V': np.mean(Vm_cycle)}
        if novertones > 0:                                # nbls.py:194-201
            Vm_coeffs = np.fft.rfft(Vm_cycle)[:novertones + 1] / NPC_DENSE
            A_Vm, phi_Vm = np.abs(Vm_coeffs), np.angle(Vm_coeffs)
            for i in range(1, novertones + 1):
                ev[f'A_V{i}'] = A_Vm[i]
                ev[f'phi_V{i}'] = phi_Vm[i]
        ev.update(eff_rates(name, Vm_cycle))
        out.append(ev)
    if stats is not None:
        stats['Z_cycle'] = Z_cycle
    return out, ncycles


def rel_cm_cycle(b, f, A, Qm=0.):
    ''' Relative capacitance profile over the last cycle (bls.py:801-808: getZlast,
        getRelCmCycle), the per-point function of scripts/run_Cm_lookups.py:19-64. '''
    tall, yall, ncycles = sim_cycles(b, f, A, Qm)
    return v_capacitance(b, yall[-NPC_DENSE:, 1]) / b.Cm0


def compute_cm_lookup(b, fref, Aref):
    ''' Restatement of run_Cm_lookups.py:19-64. :return: (refs, tables) '''
    fref, Aref = np.asarray(fref, float), np.asarray(Aref, float)
    prof = np.array([rel_cm_cycle(b, f, A, 0.) for f in fref for A in Aref])
    refs = {'f': fref, 'A': Aref, 't': np.linspace(0., 1., prof.shape[1])}
    return refs, {'Cm_rel': prof.reshape(fref.size, Aref.size, -1)}


def _point(args):
    name, a, f, A, fs, Qm = args
    t0 = time.perf_counter()
    ev, ncyc = compute_effvars(name, get_bls(name, a), f, A, fs, Qm)
    return ev, ncyc, time.perf_counter() - t0


def grid_queue(aref, fref, Aref, Qref):
    ''' Queue order of run_lookups.py:98-103,140-148: a > f > A > Q. '''
    return [(a, f, A, Q) for a in aref for f in fref for A in Aref for Q in Qref]


def compute_astim_lookup(name, aref, fref, Aref, fsref, Qref, nproc=1):
    ''' Restatement of run_lookups.py:22-175 (no overtones).
        :return: (refs, tables, ncycles[na, nf, nA, nQ]) '''
    refs = {'a': np.asarray(aref, float), 'f': np.asarray(fref, float),
            'A': np.asarray(Aref, float), 'Q': np.asarray(Qref, float)}
    fsref = np.asarray(fsref, float)
    if fsref.size > 1 or fsref[0] != 1.:
        for x in ['a', 'f']:
            assert refs[x].size == 1, 'cannot span coverage fractions for more than 1 ' + x
    refs['fs'] = fsref
    dims = tuple(v.size for v in refs.values())
    jobs = [(name, a, f, A, fsref, Q) for a, f, A, Q in
            grid_queue(refs['a'], refs['f'], refs['A'], refs['Q'])]
    if nproc > 1:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(nproc) as pool:
            outs = pool.map(_point, jobs, chunksize=1)
    else:
        outs = [_point(j) for j in jobs]
    effvars = [ev for o in outs for ev in o[0]]
    keys = list(effvars[0].keys())
    tables = {k: np.array([ev[k] for ev in effvars]).reshape(dims) for k in keys}
    tcomps = np.array([o[2] for o in outs]).reshape(dims[:-1])
    tables['tcomp'] = np.moveaxis(np.array([tcomps for _ in range(dims[-1])]), 0, -1)
    ncycles = np.array([o[1] for o in outs]).reshape(dims[:-1])
    return refs, tables, ncycles
