#!/bin/bash
timeout 600 python -m pytest tests/test_lbfgs_warp_gpu.py -x -q 2>&1 | tail -15
timeout 600 python -m pytest tests/test_fit_gpu.py tests/test_c2_search_parity_gpu.py tests/test_c3_parity_gpu.py tests/test_kernel_search_gpu.py -x -q 2>&1 | tail -3
timeout 300 python scratch/search_rounds.py 2>&1 | tail -10
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:wv_lb_step -c 300 --csv --log-file gpurun_out/r02y_lbstep.csv python scratch/search_c2.py 8 3 > /dev/null 2>&1
python - <<'PY'
import csv, statistics
rows=[r for r in csv.reader(open('gpurun_out/r02y_lbstep.csv')) if len(r)>5 and r[-1].replace('.','').isdigit()]
v=[float(r[-1])/1e3 for r in rows]
print('lb_step launches',len(v),'median us',statistics.median(v), [round(x) for x in v[::15]])
PY
