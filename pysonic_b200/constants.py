# -*- coding: utf-8 -*-
''' Numerical constants of the lookup path (mirror of PySONIC/constants.py:13,27,31,34-38). '''

Rg = 8.31342               # Universal gas constant (J.mol^-1.K^-1)
DQ_LOOKUP = 1e-5           # charge density step of lookup tables (C/m2)
MAX_RMSE_PTP_RATIO = 1e-4  # periodic convergence threshold (RMSE / peak-to-peak)
NCYCLES_MAX = 10           # max number of extra cycles in periodic simulations
CHARGE_RANGE = (-300e-5, 150e-5)  # physiological charge range (C/m2)
NPC_DENSE = 1000           # samples per acoustic period
DT_EFFECTIVE = 5e-5                 # time step of effective (SONIC) integrations (s)            constants.py:42
MAX_NSAMPLES_EFFECTIVE = 1e5        # max number of samples in the output of effective simulations  constants.py:44
SPIKE_MIN_DT = 5e-4                 # spike detection on the charge signal: min interval (s)      constants.py:49
SPIKE_MIN_QAMP = 3e-5               # ... min amplitude (C/m2)                                     constants.py:50
SPIKE_MIN_QPROM = 20e-5             # ... min prominence (C/m2)                                    constants.py:51
DT_MAX_REL_TOL = 1e-5               # max relative irregularity of a time step vector              constants.py:48
