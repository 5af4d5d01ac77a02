for x in 0 4 8 16 32; do SONIC_SCHED_EXCL_SMS=$x python tools/gpu_c2time.py excl$x 2>&1 | tail -1; done
SONIC_SCHED_EXCL_SMS=8 python tools/gpu_c2diag.py excl8 2>&1 | tail -1
