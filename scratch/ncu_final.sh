#!/bin/bash
TAG=$1
mkdir -p gpurun_out
python scratch/perf_c3.py 2000 > gpurun_out/plain_c3_$TAG.log 2>&1 || exit 1
# tensor-pipe kernels of one config-3 evaluation at the bench batch size: last trtri step, a mid Cholesky step pair, kinv
ncu --set full --clock-control none --import-source on -k "regex:wv_trtri_kernel|wv_kinv_kernel" -s 8 -c 2 -o /tmp/f_tr python scratch/perf_c3.py 2000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:wv_chol_step_kernel|wv_chol_panel_kernel" -s 8 -c 2 -o /tmp/f_ch python scratch/perf_c3.py 2000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "regex:wv_gram_kernel|wv_grad_kernel" -c 2 -o /tmp/f_el python scratch/perf_c3.py 2000 > /dev/null 2>&1
for f in tr ch el; do ncu -i /tmp/f_$f.ncu-rep --page raw --csv > gpurun_out/${TAG}_c3_${f}_raw.csv 2>/dev/null; done
python scratch/perf_c5.py 200 poisson > gpurun_out/plain_c5_$TAG.log 2>&1
ncu --set full --clock-control none -k "regex:wv_site_update_kernel" -s 5 -c 1 -o /tmp/f_si python scratch/perf_c5.py 200 poisson > /dev/null 2>&1
ncu -i /tmp/f_si.ncu-rep --page raw --csv > gpurun_out/${TAG}_c5_site_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
