#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), table for the spike-count check, ncu evidence.
# usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json
python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"
cat $OUT/bench_ref_$TAG.json
python tools/gpu_make_c1_table.py > $OUT/c1_table_$TAG.log 2>&1; echo "c1 table rc=$?"
if [ "${NO_NCU:-0}" = "0" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
  $CMD > $OUT/plain_launches_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
  echo "ncu launches rc=$?"
  CMD1="python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline"
  $CMD1 > $OUT/plain_c1_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sonic_integrate -s 1 -c 1 -f -o $OUT/prof_c1_$TAG $CMD1 > $OUT/ncu_c1_$TAG.log 2>&1
  echo "ncu full c1 rc=$?"
  CMD2="python bench.py --steps 1 --warmup 0 --no-cpu-baseline"
  $CMD2 > $OUT/plain_c2_$TAG.log 2>&1 && \
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum \
      --clock-control none -k regex:sonic_integrate -s 1 -c 1 --csv --log-file $OUT/c2_counters_$TAG.csv $CMD2 > $OUT/ncu_c2_$TAG.log 2>&1
  echo "ncu c2 counters rc=$?"
fi
