#!/bin/bash
ncu --set full --clock-control none --import-source on -k "regex:wvs_gram|wvs_grad" -c 2 -o /tmp/p_spec python scratch/perf_c3.py 2000 > /dev/null 2>&1
ncu -i /tmp/p_spec.ncu-rep --page raw --csv > gpurun_out/r02zz_spec_raw.csv 2>/dev/null
ncu -i /tmp/p_spec.ncu-rep --page source --csv --kernel-name regex:wvs_grad > gpurun_out/r02zz_grad_source.csv 2>/dev/null
ls -la gpurun_out/r02zz_spec_raw.csv gpurun_out/r02zz_grad_source.csv
