# -*- coding: utf-8 -*-
''' Point-neuron descriptors for the lookup path.

    The lookup generation only needs, per neuron: its name, resting capacitance `Cm0`, resting
    potential `Vm0` (hence `Qm0` and the default charge range `Qbounds`) and the ordered list of
    voltage-dependent rate constants that get cycle-averaged (`rates`).  Membrane currents and
    state derivatives are NOT evaluated during table generation and are out of scope.

    Each neuron is declared as a *spec*: named constants plus rate expressions written in plain
    C arithmetic on the membrane potential `Vm` (mV), result in s^-1.  `codegen.py` turns the
    specs into CUDA device functions (csrc/generated/neuron_rates.cuh) that the fused
    cycle-averaging kernel inlines; nothing here is evaluated on the CPU.

    Reference for the kinetics: PySONIC/neurons/{cortical,thalamic,stn,fh,sweeney,mrg,sundt}.py
    (cited per neuron below); helper forms `vtrap`, Borg-Graham: PySONIC/core/pneuron.py:351-413;
    alpha/beta built from (xinf, taux) pairs: PySONIC/core/translators.py:317-324; order of the
    rates = order in which the reference's `derStates` mentions them (`pneuron.rates`,
    SURVEY.md Appendix B), which is also the table order of the pickle.
'''

import re

import numpy as np

FARADAY = 9.64853e4         # constants.py:12
Rg = 8.31342                # constants.py:13
CELSIUS_2_KELVIN = 273.15   # constants.py:17
CELSIUS = 36.0              # pneuron.py:27


class Gate:
    ''' Gating variable declared through its steady state and time constant:
        alpha = xinf / tau, beta = (1 - xinf) / tau. '''

    def __init__(self, key, xinf, tau, pre=()):
        self.key, self.xinf, self.tau, self.pre = key, xinf, tau, tuple(pre)


class Rate:
    ''' Rate constant declared directly by its expression. '''

    def __init__(self, name, expr, pre=()):
        self.name, self.expr, self.pre = name, expr, tuple(pre)


def _pospischil_mhn():
    ''' cortical.py:36-58, thalamic.py:37-59 '''
    return [
        Rate('alpham', '0.32 * vtrap(13 - (Vm - VT), 4) * 1e3'),
        Rate('betam', '0.28 * vtrap((Vm - VT) - 40, 5) * 1e3'),
        Rate('alphah', '0.128 * exp(-((Vm - VT) - 17) / 18) * 1e3'),
        Rate('betah', '4 / (1 + exp(-((Vm - VT) - 40) / 5)) * 1e3'),
        Rate('alphan', '0.032 * vtrap(15 - (Vm - VT), 5) * 1e3'),
        Rate('betan', '0.5 * exp(-((Vm - VT) - 10) / 40) * 1e3'),
    ]


def _cortical_p():
    ''' cortical.py:60-66 '''
    return [Gate('p', '1.0 / (1 + exp(-(Vm + 35) / 10))',
                 'TauMax / (3.3 * exp((Vm + 35) / 20) + exp(-(Vm + 35) / 20))')]


def _huguenard_su():
    ''' cortical.py:249-266, thalamic.py:287-305 (T-type Ca2+ gates, branch in tauu) '''
    return [
        Gate('s', '1.0 / (1.0 + exp(-(Vm + Vx + 57.0) / 6.2))',
             '1.0 / 3.7 * (0.612 + 1.0 / xs) * 1e-3',
             pre=['xs = exp(-(Vm + Vx + 132.0) / 16.7) + exp((Vm + Vx + 16.8) / 18.2)']),
        Gate('u', '1.0 / (1.0 + exp((Vm + Vx + 81.0) / 4.0))',
             '((Vm + Vx < -80.0) ? 1.0 / 3.7 * exp((Vm + Vx + 467.0) / 66.6) * 1e-3'
             ' : 1.0 / 3.7 * (exp(-(Vm + Vx + 22) / 10.5) + 28.0) * 1e-3)'),
    ]


def _stn_gate(key, thx, kx, tau0, tau1, thT, sgT):
    ''' stn.py:211-336: xinf = 1/(1+exp((V-th)/k)); taux1 (one sigmoid) or taux2 (two exps) '''
    xinf = f'1 / (1 + exp((Vm - ({thx!r})) / ({kx!r})))'
    if isinstance(thT, tuple):
        (th1, th2), (s1, s2) = thT, sgT
        tau = (f'{tau0!r} + {tau1!r} / (exp(-(Vm - ({th1!r})) / ({s1!r}))'
               f' + exp(-(Vm - ({th2!r})) / ({s2!r})))')
    else:
        tau = f'{tau0!r} + {tau1!r} / (1 + exp(-(Vm - ({thT!r})) / ({sgT!r})))'
    return Gate(key, xinf, tau)


_q10_fh = 3**((CELSIUS - 20.0) / 10)          # fh.py:61-63
_q10_mrg_mp = 2.2**((CELSIUS - 20.0) / 10)    # mrg.py:59-63
_q10_mrg_h = 2.9**((CELSIUS - 20.0) / 10)
_q10_mrg_s = 3.0**((CELSIUS - 36.0) / 10)
_q10_su = 3**((CELSIUS - 30.0) / 10)          # sundt.py:60-62
_q10_hh = 3**((CELSIUS - 6.3) / 10.)          # hh.py:31,46-48
_T = CELSIUS + CELSIUS_2_KELVIN               # pneuron.py:28

NEURON_SPECS = {
    # ---- cortical (cortical.py:122-160, 163-200, 203-300, 303-401) ----
    'RS': dict(Cm0=1e-2, Vm0=-71.9, consts=dict(VT=-56.2, TauMax=0.608),
               kin=_pospischil_mhn() + _cortical_p()),
    'FS': dict(Cm0=1e-2, Vm0=-71.4, consts=dict(VT=-57.9, TauMax=0.502),
               kin=_pospischil_mhn() + _cortical_p()),
    'LTS': dict(Cm0=1e-2, Vm0=-54.0, consts=dict(VT=-50.0, TauMax=4.0, Vx=-7.0),
                kin=_pospischil_mhn() + _cortical_p() + _huguenard_su()),
    'IB': dict(Cm0=1e-2, Vm0=-71.4, consts=dict(VT=-56.2, TauMax=0.608),
               kin=_pospischil_mhn() + _cortical_p() + [
                   Rate('alphaq', '0.055 * vtrap(-(Vm + 27), 3.8) * 1e3'),
                   Rate('betaq', '0.94 * exp(-(Vm + 75) / 17) * 1e3'),
                   Rate('alphar', '0.000457 * exp(-(Vm + 13) / 50) * 1e3'),
                   Rate('betar', '0.0065 / (exp(-(Vm + 15) / 28) + 1) * 1e3')]),
    # ---- thalamic (thalamic.py:117-179, 182-366) ----
    'RE': dict(Cm0=1e-2, Vm0=-89.5, consts=dict(VT=-67.0),
               kin=_pospischil_mhn() + [
                   Gate('s', '1.0 / (1.0 + exp(-(Vm + 52.0) / 7.4))',
                        '(1 + 0.33 / (exp((Vm + 27.0) / 10.0) + exp(-(Vm + 102.0) / 15.0))) * 1e-3'),
                   Gate('u', '1.0 / (1.0 + exp((Vm + 80.0) / 5.0))',
                        '(28.3 + 0.33 / (exp((Vm + 48.0) / 4.0) + exp(-(Vm + 407.0) / 50.0)))'
                        ' * 1e-3')]),
    'TC': dict(Cm0=1e-2, Vm0=-61.93, consts=dict(VT=-52.0, Vx=0.0),
               kin=_pospischil_mhn() + _huguenard_su() + [
                   Gate('o', '1.0 / (1.0 + exp((Vm + 75.0) / 5.5))',
                        '1 / (exp(-14.59 - 0.086 * Vm) + exp(-1.87 + 0.0701 * Vm)) * 1e-3')]),
    # ---- sub-thalamic nucleus (stn.py:59-152 constants, :345-359 state order) ----
    'STN': dict(Cm0=1e-2, Vm0=-58.0, consts={}, kin=[
        _stn_gate('a', -45, -14.7, 1e-3, 1e-3, -40, -0.5),
        _stn_gate('b', -90, 7.5, 0e-3, 200e-3, (-60, -40), (-30, 10)),
        _stn_gate('c', -30.6, -5, 45e-3, 10e-3, (-27, -50), (-20, 15)),
        _stn_gate('d1', -60, 7.5, 400e-3, 500e-3, (-40, -20), (-15, 20)),
        _stn_gate('m', -40, -8, 0.2e-3, 3e-3, -53, -0.7),
        _stn_gate('h', -45.5, 6.4, 0e-3, 24.5e-3, (-50, -50), (-15, 16)),
        _stn_gate('n', -41, -14, 0e-3, 11e-3, (-40, -40), (-40, 50)),
        _stn_gate('p', -56, -6.7, 5e-3, 0.33e-3, (-27, -102), (-10, 15)),
        _stn_gate('q', -85, 5.8, 0e-3, 400e-3, (-50, -50), (-15, 16)),
    ]),
    # ---- peripheral fibers ----
    'FHnode': dict(Cm0=2e-2, Vm0=-70., consts=dict(q10=_q10_fh, V0=-70.), kin=[   # fh.py:73-103
        Rate('alpham', 'q10 * 0.36 * vtrap(22. - (Vm - V0), 3.) * 1e3'),
        Rate('betam', 'q10 * 0.4 * vtrap(Vm - V0 - 13., 20.) * 1e3'),
        Rate('alphah', 'q10 * 0.1 * vtrap(Vm - V0 + 10.0, 6.) * 1e3'),
        Rate('betah', 'q10 * 4.5 / (exp((45. - (Vm - V0)) / 10.) + 1) * 1e3'),
        Rate('alphan', 'q10 * 0.02 * vtrap(35. - (Vm - V0), 10.0) * 1e3'),
        Rate('betan', 'q10 * 0.05 * vtrap(Vm - V0 - 10., 10.) * 1e3'),
        Rate('alphap', 'q10 * 0.006 * vtrap(40. - (Vm - V0), 10.0) * 1e3'),
        Rate('betap', 'q10 * 0.09 * vtrap(Vm - V0 + 25., 20.) * 1e3'),
    ]),
    'SWnode': dict(Cm0=2.5e-2, Vm0=-80.0, consts={}, kin=[                        # sweeney.py:52-66
        Rate('alpham', 'am', pre=['am = (126 + 0.363 * Vm) / (1 + exp(-(Vm + 49) / 5.3)) * 1e3']),
        Rate('betam', 'am / (exp((Vm + 56.2) / 4.17))'),
        Rate('alphah', 'bh / exp((Vm + 74.5) / 5)',
             pre=['bh = 15.6 / (1 + exp(-(Vm + 56) / 10)) * 1e3']),
        Rate('betah', 'bh'),
    ]),
    'MRGnode': dict(Cm0=2e-2, Vm0=-80., consts=dict(                              # mrg.py:71-115
        q10_mp=_q10_mrg_mp, q10_h=_q10_mrg_h, q10_s=_q10_mrg_s), kin=[
        Rate('alpham', 'q10_mp * 1.86 * vtrap(-(Vmh + 18.4), 10.3) * 1e3', pre=['Vmh = Vm + 3.']),
        Rate('betam', 'q10_mp * 0.086 * vtrap(Vmh + 22.7, 9.16) * 1e3'),
        Rate('alphah', 'q10_h * 0.062 * vtrap(Vmh + 111.0, 11.0) * 1e3'),
        Rate('betah', 'q10_h * 2.3 / (1 + exp(-(Vmh + 28.8) / 13.4)) * 1e3'),
        Rate('alphap', 'q10_mp * 0.01 * vtrap(-(Vm + 27.), 10.2) * 1e3'),
        Rate('betap', 'q10_mp * 0.00025 * vtrap(Vm + 34., 10.) * 1e3'),
        Rate('alphas', 'q10_s * 0.3 / (1 + exp(-(Vms - 27.) / 5.)) * 1e3', pre=['Vms = Vm - (-80.)']),
        Rate('betas', 'q10_s * 0.03 / (1 + exp(-(Vms + 10.) / 1.)) * 1e3'),
    ]),
    'SUseg': dict(Cm0=1e-2, Vm0=-60., consts=dict(                                # sundt.py:82-123
        q10T=_q10_su, q10BG=_q10_su, FARADAY=FARADAY, RgT=Rg * _T), kin=[
        Rate('alpham', 'q10T * 0.32 * vtrap((13.1 - Vmm), 4) * 1e3',
             pre=['Vmm = (Vm - (-65.)) + (-6.0)', 'Vmhh = (Vm - (-65.)) + 6.0',
                  'xn = (Vm - (-32.)) * FARADAY / RgT * 1e-3',
                  'xl = (Vm - (-61.)) * FARADAY / RgT * 1e-3']),
        Rate('betam', 'q10T * 0.28 * vtrap((Vmm - 40.1), 5) * 1e3'),
        Rate('alphah', 'q10T * 0.128 * exp((17.0 - Vmhh) / 18) * 1e3'),
        Rate('betah', 'q10T * 4 / (1 + exp((40.0 - Vmhh) / 5)) * 1e3'),
        # Borg-Graham: alpha0 exp(-zeta gamma x), beta0 exp(zeta (1 - gamma) x)  (pneuron.py:387-413)
        Rate('alphan', 'q10BG * (0.03 * exp(-(-5.) * 0.4 * xn)) * 1e3'),
        Rate('betan', 'q10BG * (0.03 * exp((-5.) * (1 - 0.4) * xn)) * 1e3'),
        Rate('alphal', 'q10BG * (0.001 * exp(-(2.) * 1. * xl)) * 1e3'),
        Rate('betal', 'q10BG * (0.001 * exp((2.) * (1 - 1.) * xl)) * 1e3'),
    ]),
    # ---- giant squid axon segment (hh.py:13-75) ----
    'HHseg': dict(Cm0=1e-2, Vm0=-65.0, consts=dict(q10=_q10_hh), kin=[
        Rate('alpham', 'q10 * 0.1 * vtrap(-(Vm + 40), 10) * 1e3'),
        Rate('betam', 'q10 * 4 * exp(-(Vm + 65) / 18) * 1e3'),
        Rate('alphah', 'q10 * 0.07 * exp(-(Vm + 65) / 20) * 1e3'),
        Rate('betah', 'q10 * 1.0 / (exp(-(Vm + 35) / 10) + 1) * 1e3'),
        Rate('alphan', 'q10 * 0.01 * vtrap(-(Vm + 55), 10) * 1e3'),
        Rate('betan', 'q10 * 0.125 * exp(-(Vm + 65) / 80) * 1e3'),
    ]),
    # ---- leech touch cell (leech.py:15-160): xinf = 1 / (1 + exp((V - half) / slope))^power,
    # ---- taux = (tauMax - tauMin) / (1 + exp((V - half) / slope)) + tauMin, or constant ----
    'LeechT': dict(Cm0=1e-2, Vm0=-53.58, consts={}, kin=[
        Gate('m', '1 / (1 + exp((Vm - (-35.0)) / (-5.0)))', '0.1e-3'),
        Gate('h', '1 / (hb * hb)', '(14.0e-3 - 0.2e-3) / (1 + exp((Vm - (-36.0)) / 3.5)) + 0.2e-3',
             pre=['hb = 1 + exp((Vm - (-50.0)) / 9.0)']),
        Gate('n', '1 / (1 + exp((Vm - (-22.0)) / (-9.0)))',
             '(6.0e-3 - 1.0e-3) / (1 + exp((Vm - (-10.0)) / 10.0)) + 1.0e-3'),
        Gate('s', '1 / (1 + exp((Vm - (-10.0)) / (-2.8)))', '0.6e-3'),
    ]),
    # ---- leech pressure cell (leech.py:236-300, 369-400) ----
    'LeechP': dict(Cm0=1e-2, Vm0=-48.865, consts={}, kin=[
        Rate('alpham', '-0.03 * (Vm + 28) / (exp(-(Vm + 28) / 15) - 1) * 1e3'),
        Rate('betam', '2.7 * exp(-(Vm + 53) / 18) * 1e3'),
        Rate('alphah', '0.045 * exp(-(Vm + 58) / 18) * 1e3'),
        Rate('betah', '0.72 / (exp(-(Vm + 23) / 14) + 1) * 1e3'),
        Rate('alphan', '-0.024 * (Vm - 17) / (exp(-(Vm - 17) / 8) - 1) * 1e3'),
        Rate('betan', '0.2 * exp(-(Vm + 48) / 35) * 1e3'),
        Rate('alphas', '-1.5 * (Vm - 20) / (exp(-(Vm - 20) / 5) - 1) * 1e3'),
        Rate('betas', '1.5 * exp(-(Vm + 25) / 10) * 1e3'),
    ]),
    # ---- template neuron (template.py:13-70): the Pospischil m, h, n gates ----
    'template': dict(Cm0=1e-2, Vm0=-71.9, consts=dict(VT=-56.2), kin=_pospischil_mhn()),
    # ---- passive membrane (pas.py:16-110): no gate, only V is tabulated; Cm0 and the resting potential
    # ---- (= ELeak) are set per instance by `passiveNeuron`, these are the defaults of pas.py:103-107 ----
    'pas': dict(Cm0=1e-2, Vm0=-70., consts={}, kin=[]),
}


# ---------------------------------------------------------------------------------------------
# Membrane currents for the SONIC simulation that consumes the tables (nbls.py:280-315,389-437):
# dQm/dt = -iNet(V_eff, x) * 1e-3, dx/dt = alpha_eff (1 - x) - beta_eff x with every effective quantity
# interpolated in the lookup.  Declared for the neurons whose states are all gates (state k <-> rates
# 2k, 2k + 1); the others (TC, STN, LeechT/P, FHnode: ion concentrations, GHK currents) are not simulated.
# iNet = sum of `currents()` in declaration order (pneuron.py:162-166); constants from the neuron files.
# dt_factor: neuron-specific output time step, in units of DT_EFFECTIVE (`chooseTimeStep`: hh.py:127-129,
# sweeney.py:105-106, mrg.py:170-172, sundt.py:176-178); spike_mph: spike amplitude threshold (sundt.py:180-182).
# ---------------------------------------------------------------------------------------------
_HH = 'gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * n * (Vm - EK)'
NEURON_SIM = {
    # cortical.py:96-119,122-200
    'RS': dict(states=['m', 'h', 'n', 'p'], consts=dict(gNabar=560.0, ENa=50.0, gKdbar=60.0, EK=-90.0, gMbar=0.75,
                                                       gLeak=0.205, ELeak=-70.3),
               inet=_HH + ' + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak)'),
    'FS': dict(states=['m', 'h', 'n', 'p'], consts=dict(gNabar=580.0, ENa=50.0, gKdbar=39.0, EK=-90.0, gMbar=0.787,
                                                       gLeak=0.38, ELeak=-70.4),
               inet=_HH + ' + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak)'),
    # cortical.py:203-300 (iCaT) and :303-401 (iCaL)
    'LTS': dict(states=['m', 'h', 'n', 'p', 's', 'u'],
                consts=dict(gNabar=500.0, ENa=50.0, gKdbar=40.0, EK=-90.0, gMbar=0.28, gLeak=0.19, ELeak=-50.0,
                            gCaTbar=4.0, ECa=120.0),
                inet=_HH + ' + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak) + gCaTbar * s * s * u * (Vm - ECa)'),
    'IB': dict(states=['m', 'h', 'n', 'p', 'q', 'r'],
               consts=dict(gNabar=500.0, ENa=50.0, gKdbar=50.0, EK=-90.0, gMbar=0.3, gLeak=0.1, ELeak=-70.0,
                           gCaLbar=1.0, ECa=120.0),
               inet=_HH + ' + gMbar * p * (Vm - EK) + gLeak * (Vm - ELeak) + gCaLbar * q * q * r * (Vm - ECa)'),
    # thalamic.py:117-179
    'RE': dict(states=['m', 'h', 'n', 's', 'u'],
               consts=dict(gNabar=2000.0, ENa=50.0, gKdbar=200.0, EK=-90.0, gCaTbar=30.0, ECa=120.0, gLeak=0.5,
                           ELeak=-90.0),
               inet=_HH + ' + gCaTbar * s * s * u * (Vm - ECa) + gLeak * (Vm - ELeak)'),
    # hh.py:22-29,100-113 and template.py:26-35,94-107
    'HHseg': dict(dt_factor=1e-1, states=['m', 'h', 'n'], consts=dict(gNabar=1200.0, ENa=50.0, gKdbar=360.0, EK=-77.0, gLeak=3.0,
                                                     ELeak=-54.3),
                  inet=_HH + ' + gLeak * (Vm - ELeak)'),
    'template': dict(states=['m', 'h', 'n'], consts=dict(gNabar=560.0, ENa=50.0, gKdbar=60.0, EK=-90.0, gLeak=0.205,
                                                        ELeak=-70.3),
                     inet=_HH + ' + gLeak * (Vm - ELeak)'),
    # sweeney.py:28-35,84-93
    'SWnode': dict(dt_factor=1e-2, states=['m', 'h'], consts=dict(gNabar=14450.0, ENa=35.64, gLeak=1280.0, ELeak=-80.01),
                   inet='gNabar * m * m * h * (Vm - ENa) + gLeak * (Vm - ELeak)'),
    # mrg.py:37-47,143-162
    'MRGnode': dict(dt_factor=1e-2, states=['m', 'h', 'p', 's'],
                    consts=dict(gNafbar=30000.0, gNapbar=100.0, ENa=50.0, gKsbar=800.0, EK=-90.0, gLeak=70.0,
                                ELeak=-90.0),
                    inet='gNafbar * m * m * m * h * (Vm - ENa) + gNapbar * p * p * p * (Vm - ENa)'
                         ' + gKsbar * s * (Vm - EK) + gLeak * (Vm - ELeak)'),
    # sundt.py:38-52,150-166 (ELeak balances the currents at rest, sundt.py:64-68)
    'SUseg': dict(dt_factor=1e-2, spike_mph=-8.0e-5, states=['m', 'h', 'n', 'l'],
                  consts=dict(gNabar=400.0, ENa=55.0, gKdbar=400.0, EK=-90.0, gLeak=1.0, ELeak=-60.069175300110516),
                  inet='gNabar * m * m * m * h * (Vm - ENa) + gKdbar * n * n * n * l * (Vm - EK)'
                       ' + gLeak * (Vm - ELeak)'),
}

NEURON_ORDER = ['RS', 'FS', 'LTS', 'IB', 'RE', 'TC', 'STN', 'FHnode', 'SWnode', 'MRGnode', 'SUseg',
                'HHseg', 'LeechT', 'LeechP', 'template', 'pas']
MAX_RATES = 18


def spec_rate_names(name):
    out = []
    for k in NEURON_SPECS[name]['kin']:
        if isinstance(k, Gate):
            out += [f'alpha{k.key}', f'beta{k.key}']
        else:
            out.append(k.name)
    return out


class PointNeuron:
    ''' Minimal mirror of the reference's `PointNeuron` interface, as far as the lookup path uses
        it (pneuron.py:22-63, 268-271, 423-426; `rates` from translators.py:411). '''

    def __init__(self, name):
        if name not in NEURON_SPECS:
            raise ValueError(f'"{name}" neuron not found. Implemented neurons are: ' +
                             ', '.join(f'"{k}"' for k in NEURON_ORDER if k != 'pas'))
        spec = NEURON_SPECS[name]
        self.name = name
        self.Cm0 = spec['Cm0']
        self.Vm0 = spec['Vm0']
        self.rates = spec_rate_names(name)
        self.neuron_id = NEURON_ORDER.index(name)
        self.states = list(NEURON_SIM[name]['states']) if name in NEURON_SIM else None
        self.dt_factor = NEURON_SIM.get(name, {}).get('dt_factor', 1.0)
        self.spike_mph = NEURON_SIM.get(name, {}).get('spike_mph')

    def __repr__(self):
        return f'PointNeuron({self.name})'

    def __eq__(self, other):
        return isinstance(other, PointNeuron) and self.name == other.name

    @property
    def is_passive(self):
        return False

    @property
    def Qm0(self):
        return self.Cm0 * self.Vm0 * 1e-3  # C/m2

    @property
    def Qbounds(self):
        ''' Bounds of the physiological charge range (pneuron.py:423-426). '''
        return np.array([np.round(self.Vm0 - 35.0), 50.0]) * self.Cm0 * 1e-3  # C/m2

    def getEffRates(self, Vm):
        ''' Cycle-averaged rate constants for a membrane potential vector (pneuron.py:268-271),
            evaluated by the generated device functions on the GPU. '''
        from ._lib import eval_mean_rates
        return eval_mean_rates(self, np.asarray(Vm, dtype=np.float64))


_PAS_PATTERN = re.compile(r'pas_Cm0_{0}uF_cm2_gLeak_{0}S_m2_ELeak_{0}mV'.format(r'([+-]?\d+\.?\d*)'))


class PassiveNeuron(PointNeuron):
    ''' Point neuron with only a passive (leakage) current (pas.py:23-100): no gating variable, so
        its lookup holds the effective potential only.  Name and lookup name follow pas.py:47-63. '''

    def __init__(self, Cm0, gLeak, ELeak):
        super().__init__('pas')
        self.Cm0, self.gLeak, self.ELeak = Cm0, gLeak, ELeak
        self.Vm0 = ELeak                                   # pas.py:74-76

    def pdict(self):
        return {'Cm0': f'{self.Cm0 * 1e2:.1f} uF/cm2', 'gLeak': f'{self.gLeak:.1f} S/m2',
                'ELeak': f'{self.ELeak:.1f} mV'}

    @staticmethod
    def code(pdict):
        pdict = {k: v.replace(' ', '').replace('/', '_') for k, v in pdict.items()}
        return 'pas_' + '_'.join(f'{k}_{v}' for k, v in pdict.items())

    @property
    def name(self):
        return self.code(self.pdict())

    @name.setter
    def name(self, value):
        pass

    @property
    def lookup_name(self):
        pdict = self.pdict()
        del pdict['gLeak']
        return self.code(pdict)

    @property
    def is_passive(self):
        return True

    def __repr__(self):
        return 'PassiveNeuron(' + ', '.join(f'{k} = {v}' for k, v in self.pdict().items()) + ')'

    def __eq__(self, other):
        return isinstance(other, PassiveNeuron) and self.name == other.name


def passiveNeuron(*args):
    ''' Passive neuron from (Cm0 F/m2, gLeak S/m2, ELeak mV) or from its name (pas.py:16-22). '''
    if len(args) == 1:
        Cm0, gLeak, ELeak = [float(x) for x in re.findall(_PAS_PATTERN, args[0])[0]]
        Cm0 *= 1e-2
    else:
        Cm0, gLeak, ELeak = args
    return PassiveNeuron(Cm0, gLeak, ELeak)


def getDefaultPassiveNeuron():
    ''' pas.py:103-107 '''
    return passiveNeuron(1e-2, 1e2, -70)


def getPointNeuron(name):
    ''' Same lookup-by-name as PySONIC/neurons/__init__.py:24-44 (passive neurons by their
        parameter-carrying name, pas.py:16-19). '''
    if isinstance(name, str) and name.startswith('pas_'):
        return passiveNeuron(name)
    return PointNeuron(name)


def check_foreign_neuron(pneuron, mine, Vm=(-120., -71.9, -40., 0., 35.)):
    ''' A point-neuron object that is not this package's own (e.g. a reference PySONIC instance) is
        mapped onto the built-in kinetics by its name: make sure that it IS that neuron -- same resting
        capacitance and potential, same ordered rate names, and, when it can evaluate its rate
        functions (`effRates()`, translators.py:396-419), the same values -- instead of silently
        ignoring modified parameters. '''
    for attr in ('Cm0', 'Vm0'):
        if hasattr(pneuron, attr) and not np.isclose(getattr(pneuron, attr), getattr(mine, attr), rtol=1e-12, atol=0):
            raise ValueError(f'{pneuron}: {attr} = {getattr(pneuron, attr)} differs from the built-in '
                             f'"{mine.name}" kinetics ({getattr(mine, attr)})')
    rates = getattr(pneuron, 'rates', None)
    if rates is not None and list(rates) != list(mine.rates):
        raise ValueError(f'{pneuron}: rate constants {list(rates)} differ from the built-in "{mine.name}" '
                         f'kinetics ({mine.rates})')
    eff = getattr(pneuron, 'effRates', None)
    if callable(eff) and mine.rates:
        from . import _lib
        if _lib.device_count() > 0:
            theirs = eff()
            ours = _lib.eval_rates(mine, np.array(Vm))
            for k in mine.rates:
                ref = np.array([float(theirs[k](np.float64(v))) for v in Vm])
                if not np.allclose(ours[k], ref, rtol=1e-9, atol=0, equal_nan=True):
                    raise ValueError(f'{pneuron}: rate "{k}" differs from the built-in "{mine.name}" kinetics')
    return mine
