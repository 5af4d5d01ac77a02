#!/bin/bash
# Compact evidence set for the final build of a round: GPU tests, bench (both arms), launch list, C2 counters, workloads.
# usage (under gpurun, from the repo root): bash scripts/gpu_final.sh <tag>      (the full set is scripts/gpu_round.sh)
set -u
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log; tail -3 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; cut -c1-300 $OUT/bench_$TAG.json
python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"; cut -c1-300 $OUT/bench_ref_$TAG.json
python tools/gpu_kprobe.py $TAG > $OUT/kprobe_$TAG.json 2> $OUT/kprobe_$TAG.err; echo "kprobe rc=$?"
python tools/gpu_wl.py $TAG > $OUT/workloads_$TAG.json 2> $OUT/workloads_$TAG.err; echo "workloads rc=$?"
python tools/gpu_configs.py $TAG > $OUT/configs_$TAG.log 2>&1; echo "configs rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
$CMD > $OUT/plain_launches_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
echo "ncu launches rc=$?"
