run() { env "$@" python tools/gpu_wl.py "$*" ${WL:-c2,shard2,shard8,FHnode,TC} 2>&1 | grep -E "\"tag\"" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], {k: v['ms'][-1] for k, v in d.items() if k != 'tag'})
"; }
for w in 1 0; do for sl in 1.0 1.1 1.2 1.3; do for q in 1.0 1.3 1.6; do run SONIC_WIDEN=$w SONIC_SCHED_STAGED_SLOWDOWN=$sl SONIC_SCHED_QUEUE_OVERHEAD=$q; done; done; done
