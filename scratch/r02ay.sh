#!/bin/bash
SOLO=1 ncu --set full --clock-control none -k "regex:wv_chol_all_kernel|wv_trtri_all_kernel" -c 2 -o /tmp/p_pers python scratch/perf_c3.py 250 > /dev/null 2>&1
ncu -i /tmp/p_pers.ncu-rep --page raw --csv > gpurun_out/r02end_persistent_raw.csv 2>/dev/null
ls -la gpurun_out/r02end_persistent_raw.csv
