#!/bin/bash
# Short GPU-box session: parity tests + bench (own arm) + e2e phase timing.  usage: bash scripts/gpu_quick.sh <tag>
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -15 $OUT/pytest_gpu_$TAG.log
timeout 600 python bench.py --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
SONIC_DEBUG=1 timeout 300 python tools/e2e_probe.py > $OUT/e2e_probe_$TAG.log 2>&1; tail -12 $OUT/e2e_probe_$TAG.log
