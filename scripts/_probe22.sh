CMD1="python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
$CMD1 > gpurun_out/plain_c1_fo5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sonic_integrate -s 1 -c 1 -f -o gpurun_out/prof_c1_fo5 $CMD1 > gpurun_out/ncu_c1_fo5.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/plain_c1_fo5.log | cut -c1-300
