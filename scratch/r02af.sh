#!/bin/bash
# final-state ncu --set full captures of the factorisation kernels at the bench batch size
mkdir -p gpurun_out
python scratch/perf_c3.py 2000 > gpurun_out/r02z_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:wv_chol_panel_kernel -s 4 -c 1 -o /tmp/p_panel python scratch/perf_c3.py 2000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wv_chol_step_kernel -s 5 -c 1 -o /tmp/p_diag python scratch/perf_c3.py 2000 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wv_trtri_kernel -s 8 -c 1 -o /tmp/p_trtri python scratch/perf_c3.py 2000 > /dev/null 2>&1
for f in panel diag trtri; do
  ncu -i /tmp/p_$f.ncu-rep --page raw --csv > gpurun_out/r02z_${f}_raw.csv 2>/dev/null
done
ncu -i /tmp/p_panel.ncu-rep --page source --csv > gpurun_out/r02z_panel_source.csv 2>/dev/null
ls -la gpurun_out/r02z*
