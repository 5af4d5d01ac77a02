CMD2="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
$CMD2 > gpurun_out/plain_c2_r02k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sonic_integrate -s 1 -c 1 -f -o gpurun_out/prof_c2_r02k $CMD2 > gpurun_out/ncu_c2full_r02k.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_c2full_r02k.log
