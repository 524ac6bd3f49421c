#!/bin/bash
# usage: scratch/ncu_capture.sh <tag>   (run on the GPU box under gpurun)
set -x
TAG=$1
mkdir -p gpurun_out
python scratch/perf_c3.py 500 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:wv_chol_diag -c 6 -o /tmp/p_diag python scratch/perf_c3.py 500 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:wv_gram|wv_grad|wv_kinv" -c 3 -o /tmp/p_elem python scratch/perf_c3.py 500 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wv_panel -s 4 -c 2 -o /tmp/p_panel python scratch/perf_c3.py 500 > /dev/null 2>&1
for f in diag elem panel; do
  ncu -i /tmp/p_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_${f}_raw.csv 2>/dev/null
  ncu -i /tmp/p_$f.ncu-rep --page details --csv > gpurun_out/${TAG}_${f}_details.csv 2>/dev/null
done
ncu -i /tmp/p_diag.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${TAG}_diag_source.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/
