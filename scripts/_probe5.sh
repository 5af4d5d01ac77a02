export PYSONIC_B200_LIB=$PWD/pysonic_b200/variants/libsonic_sm.so
SONIC_NESTED=0 python tools/gpu_c2diag.py n0 2>&1 | tail -1
python tools/gpu_c2diag.py nauto 2>&1 | tail -1
