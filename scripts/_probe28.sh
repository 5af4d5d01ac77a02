run() { env "$@" SONIC_DEBUG=1 python tools/gpu_wl.py "$*" ${WL:-c2} 2>&1 | grep -E "schedule|\"tag\"" | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['tag'], {k: (v['ms'][-1], v['longest_s']) for k, v in d.items() if k != 'tag'})
    else: print('   ', l.strip()[17:])
"; }
for c in "8 0" "8 16" "4 24" "8 32" "16 32" "0 32"; do set -- $c; run SONIC_SCHED_TIER1_SMS=$1 SONIC_SCHED_TIER2_SMS=$2; done
