import sys
import numpy as np
d=np.load(sys.argv[1])
tp,nr,nc,st,f,A,ia,Q=d['tpoint'],d['nrhs'].astype(float),d['ncyc'],d['status'],d['f'],d['A'],d['ia'],d['Q']
o=np.argsort(-tp)
print('top tpoint:')
for i in o[:8]: print(f'a{ia[i]} f={f[i]/1e3:.0f} A={A[i]/1e3:.1f} Q={Q[i]*1e5:.0f} tp={tp[i]:.3f} nrhs={nr[i]:.0f} us/tick(lsoda)={tp[i]/nr[i]*1e6:.2f} nc={nc[i]}')
o2=np.argsort(-nr)
print('top nrhs:')
for i in o2[:6]: print(f'a{ia[i]} f={f[i]/1e3:.0f} A={A[i]/1e3:.1f} Q={Q[i]*1e5:.0f} tp={tp[i]:.3f} nrhs={nr[i]:.0f} us/tick={tp[i]/nr[i]*1e6:.2f} nc={nc[i]}')
print('sum tp',tp.sum(),'max',tp.max(), 'sum nrhs %.3e'%nr.sum())
q=np.quantile(nr,[0,.5,.9,.99,.999,1])
for lo,hi in zip(q[:-1],q[1:]):
    m=(nr>=lo)&(nr<=hi); print(f'nrhs {lo:.0f}-{hi:.0f}: n={m.sum()} mean us/tick {np.mean(tp[m]/nr[m])*1e6:.2f} min {np.min(tp[m]/nr[m])*1e6:.2f} max tp {tp[m].max():.2f}')
