import sys, time
sys.path.insert(0,'.')
import bench, pysonic_b200 as ps
w=bench.workload('c2'); pn=ps.getPointNeuron('RS')
for i in range(3):
    t0=time.perf_counter(); lkp=ps.computeAStimLookup(pn,w['a'],w['f'],w['A'],w['fs'],w['Q'],loglevel=10); print('e2e call',i,'%.3f s'%(time.perf_counter()-t0))
