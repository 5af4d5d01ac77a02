# static SASS instruction count of the integrator kernel per source function (nvdisasm line info)
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_profile as sp
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), '..', 'pysonic_b200', 'libsonic_b200.so')
kernel = sys.argv[2] if len(sys.argv) > 2 else '_Z22sonic_integrate_kernelILb0EEv8SonicJob'
sass = sp.sass_lines(so, kernel)
here = os.path.dirname(os.path.abspath(__file__))
fmaps = {'sonic_core.h': sp.function_map(os.path.join(here, '..', 'pysonic_b200', 'csrc', 'sonic_core.h')),
         'sonic_b200.cu': sp.function_map(os.path.join(here, '..', 'pysonic_b200', 'csrc', 'sonic_b200.cu'))}
cnt = collections.Counter()
for addr, (fname, line), txt in sass:
    cnt[fmaps.get(fname, {}).get(line, fname)] += 1
print('total', len(sass))
for k, v in cnt.most_common(40):
    print(f'{k:34s} {v:6d}')
