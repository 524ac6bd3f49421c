#!/bin/bash
# usage: scratch/ncu_large.sh <tag>   (GPU box) launch list of one n=8192 evaluation + full capture of a diag / syrk / trtri launch
TAG=$1
mkdir -p gpurun_out
python scratch/perf_large.py 512 16 1 > gpurun_out/plain_large_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_large_$TAG.csv python scratch/perf_large.py 512 16 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wv_chol_diag -s 70 -c 2 -o /tmp/l_diag python scratch/perf_large.py 512 16 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:wv_syrk|wv_trtri_level" -s 6 -c 3 -o /tmp/l_syrk python scratch/perf_large.py 512 16 1 > /dev/null 2>&1
for f in diag syrk; do
  ncu -i /tmp/l_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_large_${f}_raw.csv 2>/dev/null
done
ncu -i /tmp/l_diag.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/${TAG}_large_diag_source.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
