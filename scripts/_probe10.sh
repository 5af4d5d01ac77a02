python tools/gpu_c2time.py powr_double 2>&1 | tail -1
python tools/gpu_parity_report.py gpurun_out/parity_r02_powd.json 2>&1 | grep "c1_RS\|c2_RS\|c4_SW\|c5_TC\|c3_STN_big"
