for v in "$@"; do PYSONIC_B200_LIB=$PWD/pysonic_b200/variants/libsonic_$v.so python tools/gpu_kprobe.py $v 2>&1 | tail -1 | tee -a gpurun_out/kprobe.jsonl; done
