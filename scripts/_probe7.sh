export PYSONIC_B200_LIB=$PWD/pysonic_b200/variants/libsonic_sm.so
for k in 3.0 5.0 6.0 8.0; do SONIC_NESTED=0 SONIC_SCHED_GAIN=1.8 SONIC_SCHED_KDEC=$k python tools/gpu_c2time.py staged_gain1.8_kdec${k} 2>&1 | tail -1; done
for k in 4.0 6.0; do SONIC_SCHED_GAIN=1.8 SONIC_SCHED_KDEC=$k python tools/gpu_c2time.py auto_gain1.8_kdec${k} 2>&1 | tail -1; done
SONIC_NESTED=0 SONIC_SCHED_GAIN=1.8 SONIC_SCHED_KDEC=4.0 python tools/gpu_c2diag.py k4 2>&1 | tail -1
