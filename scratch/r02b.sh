#!/bin/bash
# round 2, call b: one config-3 evaluation of 2000 models, interpreter vs specialised element-wise kernels, + ncu of the latter
mkdir -p gpurun_out
WV_SPECIALIZE=0 python scratch/perf_c3.py 2000 > gpurun_out/r02b_plain_interp.log 2>&1; tail -3 gpurun_out/r02b_plain_interp.log
python scratch/perf_c3.py 2000 > gpurun_out/r02b_plain_spec.log 2>&1 || exit 1; tail -3 gpurun_out/r02b_plain_spec.log
ncu --set full --clock-control none --import-source on -k "regex:wvs_gram|wvs_grad" -c 2 -o /tmp/f_el python scratch/perf_c3.py 2000 > gpurun_out/r02b_ncu.log 2>&1
ncu -i /tmp/f_el.ncu-rep --page raw --csv > gpurun_out/r02b_spec_raw.csv 2>/dev/null
ncu -i /tmp/f_el.ncu-rep --page source --csv -k regex:wvs_gram > gpurun_out/r02b_gram_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
