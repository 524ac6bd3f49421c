#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_specialize_gpu.py -x -q > gpurun_out/r02c_spec_tests.log 2>&1
echo "spec tests rc=$?"; tail -5 gpurun_out/r02c_spec_tests.log
python scratch/perf_c3.py 2000 > gpurun_out/r02c_plain_spec.log 2>&1 || exit 1; tail -3 gpurun_out/r02c_plain_spec.log
ncu --set full --clock-control none --import-source on -k "regex:wvs_gram|wvs_grad" -c 2 -o /tmp/f_el python scratch/perf_c3.py 2000 > gpurun_out/r02c_ncu.log 2>&1
ncu -i /tmp/f_el.ncu-rep --page raw --csv > gpurun_out/r02c_spec_raw.csv 2>/dev/null
ncu -i /tmp/f_el.ncu-rep --page source --csv -k regex:wvs_gram > gpurun_out/r02c_gram_source.csv 2>/dev/null
ncu -i /tmp/f_el.ncu-rep --page source --csv -k regex:wvs_grad > gpurun_out/r02c_grad_source.csv 2>/dev/null
