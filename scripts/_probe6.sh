export PYSONIC_B200_LIB=$PWD/pysonic_b200/variants/libsonic_sm.so SONIC_NESTED=0
for g in 1.55 1.8 2.1 2.5; do for k in 2.5 4.0; do SONIC_SCHED_GAIN=$g SONIC_SCHED_KDEC=$k python tools/gpu_c2time.py gain${g}_kdec${k} 2>&1 | tail -1; done; done
