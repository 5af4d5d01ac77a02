run() { env "$@" python tools/gpu_wl.py "$*" ${WL:-c2,shard2,shard8,FHnode,TC} 2>&1 | grep -E "\"tag\"" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['tag'], {k: v['ms'][-1] for k, v in d.items() if k != 'tag'})
"; }
for m in 0.9 0.8; do for q in 1.3 1.6 1.9; do run SONIC_SCHED_TIER_MARGIN=$m SONIC_SCHED_STAGED_SLOWDOWN=1.3 SONIC_SCHED_QUEUE_OVERHEAD=$q; done; done
