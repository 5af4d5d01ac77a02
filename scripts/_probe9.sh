for x in 8 16 32 48; do SONIC_NESTED=2 SONIC_SCHED_EXCL_SMS=$x python tools/gpu_c2time.py excl${x}_nested 2>&1 | tail -1; done
SONIC_NESTED=2 SONIC_SCHED_EXCL_SMS=16 python tools/gpu_c2diag.py excl16n 2>&1 | tail -1
