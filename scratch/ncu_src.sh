#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:wv_gram_kernel" -c 1 -o /tmp/g_src python scratch/perf_c3.py 2000 > /dev/null 2>&1
ncu -i /tmp/g_src.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/gram_source.csv 2>/dev/null || ncu -i /tmp/g_src.ncu-rep --page source --csv > gpurun_out/gram_source.csv 2>/dev/null
ls -la gpurun_out/gram_source.csv; head -c 1500 gpurun_out/gram_source.csv
