# -*- coding: utf-8 -*-
''' Numerical constants of the lookup path (mirror of PySONIC/constants.py:13,27,31,34-38). '''

Rg = 8.31342               # Universal gas constant (J.mol^-1.K^-1)
DQ_LOOKUP = 1e-5           # charge density step of lookup tables (C/m2)
MAX_RMSE_PTP_RATIO = 1e-4  # periodic convergence threshold (RMSE / peak-to-peak)
NCYCLES_MAX = 10           # max number of extra cycles in periodic simulations
CHARGE_RANGE = (-300e-5, 150e-5)  # physiological charge range (C/m2)
NPC_DENSE = 1000           # samples per acoustic period
