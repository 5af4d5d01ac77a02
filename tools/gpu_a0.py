import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
import pysonic_b200 as ps
bls = ps.BilayerSonophore(32e-9, 1e-2, 0.0)
for f in (100e3, 500e3, 4e6):
    z = bls.getZlast(ps.AcousticDrive(f, 0.0), 0.)
    cm = bls.getRelCmCycle(ps.AcousticDrive(f, 0.0), 0.)
    print(f'f={f:.0f} Z ptp/mean {np.ptp(z)/abs(z.mean()):.3e}  Cm_rel ptp {np.ptp(cm):.3e}  Z0 {z[0]:.6e} Zend {z[-1]:.6e}')
