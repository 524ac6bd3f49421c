#!/bin/bash
TAG=$1
mkdir -p gpurun_out
python scratch/perf_c3.py 500 > gpurun_out/plain_elem_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k "regex:wv_gram|wv_grad" -c 2 -o /tmp/p_elem python scratch/perf_c3.py 500 > /dev/null 2>&1
ncu -i /tmp/p_elem.ncu-rep --page raw --csv > gpurun_out/${TAG}_elem_raw.csv 2>/dev/null
ncu -i /tmp/p_elem.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${TAG}_gram_source.csv 2>/dev/null
ncu -i /tmp/p_elem.ncu-rep --page source --csv --kernel-id :::2 > gpurun_out/${TAG}_grad_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
