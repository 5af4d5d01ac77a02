TAG=r02i
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log; tail -3 $OUT/pytest_gpu_$TAG.log
python tools/gpu_workload.py STN > $OUT/plain_STN_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sonic_average -c 1 -f -o $OUT/prof_avg_stn_$TAG python tools/gpu_workload.py STN > $OUT/ncu_avg_$TAG.log 2>&1
echo "ncu full averaging rc=$?"; tail -1 $OUT/plain_STN_$TAG.log | cut -c1-600
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 $OUT/bench_$TAG.json
