for k in 0 1; do SONIC_LONE_MAXK=$k PYSONIC_B200_LIB=$PWD/pysonic_b200/variants/libsonic_rt.so python tools/gpu_c2diag.py maxk$k 2>&1 | tail -1; done
