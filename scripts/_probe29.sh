run() { env "$@" SONIC_DEBUG=1 python tools/gpu_wl.py "$*" ${WL:-c2} 2>&1 | grep -E "schedule|\"tag\"" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['tag'], {k: (v['ms'][-1], v['longest_s']) for k, v in d.items() if k != 'tag'})
    else: print('   ', l.strip()[17:])
"; }
for q in 0.7 0.85 1.0 1.15; do for sl in 1.15 1.3; do run SONIC_SCHED_STAGED_SLOWDOWN=$sl SONIC_SCHED_QUEUE_OVERHEAD=$q; done; done
