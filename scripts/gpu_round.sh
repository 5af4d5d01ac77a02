#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), parity report, ncu evidence, workloads, tables for the
# spike-count check.  usage (from the repo root, under gpurun): bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi_$TAG.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu_$TAG.log
tail -5 $OUT/pytest_gpu_$TAG.log
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
cat $OUT/bench_$TAG.json | cut -c1-600
python bench.py --impl reference > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$?"
cat $OUT/bench_ref_$TAG.json | cut -c1-400
python tools/gpu_parity_report.py $OUT/parity_$TAG.json > $OUT/parity_$TAG.log 2>&1; echo "parity report rc=$?"
python tools/gpu_configs.py $TAG > $OUT/configs_$TAG.log 2>&1; echo "configs rc=$?"
python tools/gpu_offnode.py > $OUT/offnode_$TAG.json 2> $OUT/offnode_$TAG.err; echo "offnode rc=$?"
python tools/gpu_overtones.py $TAG > $OUT/overtones_$TAG.log 2>&1; echo "overtones rc=$?"
python tools/gpu_make_sim_tables.py > $OUT/sim_tables_$TAG.log 2>&1; echo "tables rc=$?"
python tools/gpu_kprobe.py $TAG > $OUT/kprobe_$TAG.json 2> $OUT/kprobe_$TAG.err; echo "kprobe rc=$?"
# calibration of the scheduler: tick of lone lanes (register-resident run / staged tick) and of staged warps with k busy lanes
{ python tools/gpu_lone_rates.py | tail -1; SONIC_NESTED=0 python tools/gpu_lone_rates.py | tail -1; python tools/gpu_tk.py; } > $OUT/ticks_$TAG.jsonl 2> $OUT/ticks_$TAG.err; echo "ticks rc=$?"
python tools/gpu_wl.py $TAG > $OUT/workloads_$TAG.json 2> $OUT/workloads_$TAG.err; echo "workloads rc=$?"
if [ "${NO_NCU:-0}" = "0" ]; then
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras"
  $CMD > $OUT/plain_launches_$TAG.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
  echo "ncu launches rc=$?"
  CMD1="python bench.py --workload c1 --steps 1 --warmup 1 --no-cpu-baseline --no-extras"
  $CMD1 > $OUT/plain_c1_$TAG.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sonic_integrate -s 1 -c 1 -f -o $OUT/prof_c1_$TAG $CMD1 > $OUT/ncu_c1_$TAG.log 2>&1
  echo "ncu full c1 rc=$?"
  CMD2="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras"
  MET=dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__icc_request_hit_rate.pct,sass__inst_executed_local_loads,sass__inst_executed_local_stores,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum
  $CMD2 > $OUT/plain_c2_$TAG.log 2>&1 && \
  ncu --metrics $MET --clock-control none -k regex:sonic_integrate -s 1 -c 1 --csv --log-file $OUT/c2_counters_$TAG.csv $CMD2 > $OUT/ncu_c2_$TAG.log 2>&1
  echo "ncu c2 counters rc=$?"
  # averaging kernel on the C3 shape (STN, 100 coverage fractions, 19 tables), integrator on one C4 and one C5 grid
  for W in STN SWnode TC; do
    CMDW="python tools/gpu_workload.py $W"
    $CMDW > $OUT/plain_${W}_$TAG.log 2>&1 && \
    ncu --metrics $MET --clock-control none -k regex:"sonic_integrate|sonic_average" -c 4 --csv --log-file $OUT/counters_${W}_$TAG.csv $CMDW > $OUT/ncu_${W}_$TAG.log 2>&1
    echo "ncu $W rc=$?"
  done
  ncu --set full --clock-control none --import-source on -k regex:sonic_average -c 1 -f -o $OUT/prof_avg_stn_$TAG python tools/gpu_workload.py STN > $OUT/ncu_avg_$TAG.log 2>&1
  echo "ncu full averaging rc=$?"
fi
