python tools/gpu_diag.py r02_c2 c2 2>&1 | tail -2
python tools/gpu_diag.py r02_fh FHnode 2>&1 | tail -1
python tools/gpu_diag.py r02_sw SWnode 2>&1 | tail -1
python tools/gpu_diag.py r02_mrg MRGnode 2>&1 | tail -1
