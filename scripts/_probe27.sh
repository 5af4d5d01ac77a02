run() { env "$@" SONIC_DEBUG=1 python tools/gpu_wl.py "$*" ${WL:-c2,shard2,shard8,FHnode,TC} 2>&1 | grep -E "schedule|\"tag\"" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['tag'], {k: v['ms'][-1] for k, v in d.items() if k != 'tag'})
    else: print('   ', l.strip()[17:])
"; }
for q in 1.3 1.6 1.9; do run SONIC_SCHED_STAGED_SLOWDOWN=1.3 SONIC_SCHED_QUEUE_OVERHEAD=$q; done
run SONIC_SCHED_STAGED_SLOWDOWN=1.15 SONIC_SCHED_QUEUE_OVERHEAD=1.6
